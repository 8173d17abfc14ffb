"""Build libb200aug.so (the C-ABI CUDA library) in-tree for sm_100a.  `python build.py` or `build()`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib", "libb200aug.so")
SOURCES = [os.path.join(HERE, "csrc", "b200aug_fused.cu"), os.path.join(HERE, "csrc", "b200aug_jpeg.cu")]
HEADERS = [os.path.join(HERE, "csrc", "b200aug_math.cuh"), os.path.join(ROOT, "include", "b200aug.h")]
# -fmad=false: the resamplers and the 2x3 algebra reproduce host arithmetic bit for bit; FMAs are spelled explicitly.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    # nvJPEG (the JPEG -> grayscale frame entry) is linked statically: the library must not depend on a CUDA toolkit install
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + ["-lnvjpeg_static", "-lculibos"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
