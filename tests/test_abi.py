"""The C-ABI library: loads without a GPU, exports every symbol include/b200aug.h declares, the ctypes mirrors of its
structs have the C layout, and argument validation answers with error codes before anything touches the device."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200aug.h")


@pytest.fixture(scope="module")
def native():
    sys.path.insert(0, ROOT)
    import __graft_entry__  # noqa: F401  (puts the package on sys.path)

    lib = os.path.join(ROOT, "neuralnet-tracker-traincode_b200", "lib", "libb200aug.so")
    if not os.path.exists(lib):
        __graft_entry__.build()
    from trackertraincode_b200 import _native

    return _native


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200aug_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(native):
    names = declared_functions()
    assert "b200aug_fused_forward" in names and "b200aug_photometric_f32" in names
    for n in names:
        assert hasattr(native.lib, n), f"{n} declared in include/b200aug.h but not exported"
    assert sorted(native.EXPORTS) == names, "the ctypes binding must cover exactly the header's entry points"


def test_abi_version_and_strerror(native):
    m = re.search(r"#define B200AUG_ABI_VERSION (\d+)", open(HEADER).read())
    assert native.lib.b200aug_abi_version() == int(m.group(1)) == native.ABI_VERSION
    assert native.lib.b200aug_strerror(0) == b"ok"
    assert b"invalid" in native.lib.b200aug_strerror(1)
    assert native.lib.b200aug_strerror(12345) == b"unknown error"


def test_struct_layouts_match_c(native):
    """sizeof/offsetof as the C compiler sees them vs the ctypes mirrors."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200aug.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(B200AugSrc), sizeof(B200AugField), sizeof(B200AugPhotoParams), sizeof(B200AugFusedArgs));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(B200AugFusedArgs, src_table), offsetof(B200AugFusedArgs, fields),
         offsetof(B200AugFusedArgs, view_roi_out), offsetof(B200AugFusedArgs, order), offsetof(B200AugFusedArgs, workspace_stride),
         offsetof(B200AugFusedArgs, photo));
  printf("%zu %zu %zu\n", offsetof(B200AugPhotoParams, apply), offsetof(B200AugPhotoParams, noise_std), offsetof(B200AugPhotoParams, seed));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    v = [int(x) for x in out]
    N = native
    assert v[:4] == [C.sizeof(N.Src), C.sizeof(N.Field), C.sizeof(N.PhotoParams), C.sizeof(N.FusedArgs)]
    F = N.FusedArgs
    assert v[4:10] == [F.src_table.offset, F.fields.offset, F.view_roi_out.offset, F.order.offset, F.workspace_stride.offset,
                       F.photo.offset]
    P = N.PhotoParams
    assert v[10:] == [P.apply.offset, P.noise_std.offset, P.seed.offset]


def test_argument_validation_needs_no_gpu(native):
    N = native
    lib = N.lib
    assert lib.b200aug_fused_forward(None, None) == 1
    a = N.FusedArgs()
    assert lib.b200aug_fused_forward(C.byref(a), None) == 1  # struct_size not set
    a.struct_size = C.sizeof(N.FusedArgs)
    a.batch, a.out_w, a.out_h = 0, 129, 129
    assert lib.b200aug_fused_forward(C.byref(a), None) == 0  # empty batch: nothing to do, no launch
    a.batch = -1
    assert lib.b200aug_fused_forward(C.byref(a), None) == 1
    a.batch, a.flags = 4, N.F_FOCUS  # focus without parameters
    assert lib.b200aug_fused_forward(C.byref(a), None) == 1
    a.flags = N.F_WHITEN  # whiten without normalize
    assert lib.b200aug_fused_forward(C.byref(a), None) == 1
    a.flags, a.n_fields = 0, N.MAX_FIELDS + 1
    assert lib.b200aug_fused_forward(C.byref(a), None) == 1
    assert lib.b200aug_apply_affine2d(None, 0, 1, 1, None, None) == 1
    p = N.PhotoParams()
    assert lib.b200aug_photometric_f32(None, None, None, 1, 8, 8, C.byref(p), 0.0, None) == 1
    # shared-memory budget: the pose-net geometry fits, an absurd one does not
    assert 0 < lib.b200aug_fused_smem_bytes(129, 129, 0) <= 227 * 1024
    assert lib.b200aug_fused_smem_bytes(2048, 2048, 0) == 0
    assert lib.b200aug_workspace_stride(512) >= 512 * 512 and lib.b200aug_workspace_stride(0) == 0


def test_product_path_fails_loudly_without_gpu_or_library(native):
    """No CPU fallback: CPU tensors are refused, and a missing library is an import error (not a silent detour)."""
    import torch

    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation, batch as dtb

    b = Batch(Metadata((64, 64), 2, None, None, {"image": FieldCategory.image, "roi": FieldCategory.roi}),
              {"image": torch.zeros(2, 64, 64, dtype=torch.uint8), "roi": torch.tensor([[8.0, 8, 40, 40]] * 2)})
    if not torch.cuda.is_available():
        with pytest.raises((native.NativeError, RuntimeError, AssertionError)):
            FusedPoseAugmentation(32, device="cpu")(b)
        with pytest.raises(native.NativeError):
            dtb.normalize_batch(b)
        with pytest.raises(native.NativeError):
            dtb.photometric_f32(torch.zeros(1, 1, 8, 8), None)
    code = ("import sys; sys.path.insert(0, %r); import trackertraincode_b200._native as n" % os.path.join(ROOT, "neuralnet-tracker-traincode_b200"))
    env = dict(os.environ)
    r = subprocess.run([sys.executable, "-c", "import os\n" + code.replace("import trackertraincode_b200._native as n",
                        "import importlib, types\nimport trackertraincode_b200.datasets\n"
                        "import trackertraincode_b200._native as n\nn.LIB_PATH='/nonexistent/lib.so'\n"
                        "try:\n    n._load()\n    print('LOADED')\nexcept n.NativeError as e:\n    print('RAISED')")],
                       capture_output=True, text=True, env=env)
    assert "RAISED" in r.stdout, r.stdout + r.stderr
