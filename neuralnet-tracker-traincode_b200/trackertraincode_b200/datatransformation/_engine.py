"""Marshals a `Batch` into one call of the fused CUDA kernel (include/b200aug.h: b200aug_fused_forward).

Everything here is host-side plumbing: pointer/shape bookkeeping, output allocation, parameter upload.  The
arithmetic of the path lives in csrc/.  No CPU fallback exists: tensors must be on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _native as N
from ..datasets.batch import Batch, FieldCategory, as_category, imagelike_categories

_CAT_CODE = {
    FieldCategory.quat: N.CAT_QUAT,
    FieldCategory.xys: N.CAT_XYS,
    FieldCategory.roi: N.CAT_ROI,
    FieldCategory.points: N.CAT_POINTS,
}
_CAT_DIMS = {N.CAT_QUAT: (4,), N.CAT_XYS: (3,), N.CAT_ROI: (4,), N.CAT_POINTS: (2, 3)}

STATUS_TEXT = {N.S_EMPTY_BOX: "empty view box",
               N.S_UNSUPPORTED: f"anti-alias prefilter wider than {N.PREFILTER_MAX_TAPS} taps, or its canvas exceeds the workspace",
               N.S_ROWBUF: "source segment exceeds rowbuf_capacity"}


@dataclass
class GeoParams:
    """RoiFocusRandomizationParameters (+ optional host-evaluated cos/sin) as device tensors."""

    scales: torch.Tensor
    angles: torch.Tensor
    translations: torch.Tensor
    cos_sin: Optional[torch.Tensor] = None


@dataclass
class PhotoParams:
    """One call's draws of the two KorniaImageDistortions stages (see include/b200aug.h: B200AugPhotoParams)."""

    order: Sequence[int]
    apply: torch.Tensor  # bool/uint8 [B, 6]
    bits: torch.Tensor  # int32 [B]
    gamma: torch.Tensor  # float32 [B]
    contrast: torch.Tensor
    brightness: torch.Tensor
    noise_apply: torch.Tensor  # bool/uint8 [B, 4]
    noise_std: Sequence[float] = (4.0 / 255.0, 16.0 / 255.0, 32.0 / 255.0, 64.0 / 255.0)
    seed: int = 0
    sample_offset: int = 0
    clip: bool = True
    noise_clip: Sequence[bool] = (False, False, False, False)  # RandomGaussianNoiseWithClipping per stage


@dataclass
class FusedResult:
    batch: Batch
    view_roi: Optional[torch.Tensor] = None
    tr: Optional[torch.Tensor] = None
    status: Optional[torch.Tensor] = None
    trace: Optional[torch.Tensor] = None


class PreparedCall:
    """A fully marshalled b200aug_fused_forward call: argument block, output tensors and the inputs it points to.
    `launch()` enqueues it on the current stream and can be repeated (same inputs -> same outputs), which is what a
    steady-state loop or a CUDA-graph capture wants."""

    def __init__(self, args, device, result: FusedResult, keep, bound_stream: Optional[int] = None):
        self.args, self.device, self.result, self._keep = args, device, result, keep
        # a call that uses the per-(device, stream) scratch caches may only run on the stream it was prepared on: another
        # stream's launches would share its canvases and plan records (private_scratch=True lifts the restriction)
        self.bound_stream = bound_stream

    def launch(self, stream: Optional[int] = None, phase: int = N.PHASE_ALL) -> FusedResult:
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.bound_stream is not None and stream != self.bound_stream:
            raise N.NativeError("this call shares scratch buffers with other calls prepared on its stream; prepare it with "
                                "private_scratch=True to launch it on a different stream")
        self.args.phase = phase
        N.check(N.lib.b200aug_fused_forward(C.byref(self.args), C.c_void_p(stream)), "b200aug_fused_forward")
        return self.result

    def launch_plan(self, stream: Optional[int] = None) -> None:
        """Phase 1 alone (plan_kernel: plans, resize tables, labels, side outputs).  A loop that pipelines steps runs this
        for step s + 1 on a second stream while `launch_main` of step s is still at work; the call must have been prepared
        with `private_scratch=True` so that its plan records are not another call's."""
        self.launch(stream, N.PHASE_PLAN)

    def launch_main(self, stream: Optional[int] = None) -> FusedResult:
        """Phase 2 alone (canvas workers + resampling + photometric chain), after `launch_plan` has completed (event order is
        the caller's business)."""
        return self.launch(stream, N.PHASE_MAIN)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise N.NativeError(f"{what} must live on a CUDA device (got {t.device}); this path has no CPU implementation")


def _require_device_visible(t: torch.Tensor, what: str):
    """Source frames may also stay in PINNED host memory: under unified addressing the kernel reads page-locked host
    memory through the same pointer (zero copy), so only the bytes inside the view boxes ever cross PCIe."""
    if not (t.is_cuda or t.is_pinned()):
        raise N.NativeError(f"{what} must live on a CUDA device or in pinned host memory (got pageable {t.device} memory); "
                            "this path has no CPU implementation")


def _dev(t, device, dtype) -> torch.Tensor:
    t = torch.as_tensor(t)
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


class _HostPack:
    """All the small host-side parameter arrays of one call travel as ONE pinned staging buffer and ONE host->device copy
    (a dozen pageable `.to(device)` calls cost more host time than the kernel takes).  `add` returns a token; after
    `upload()` `ptr(token)` is the device address of that array."""

    _pools: Dict[Any, List[Any]] = {}  # device -> [(pinned buffer, event of its last upload)]

    def __init__(self, device):
        self.device = device
        self.items: List[Tuple[int, torch.Tensor]] = []
        self.total = 0
        self.dev: Optional[torch.Tensor] = None

    def add(self, t, dtype, shape):
        t = torch.as_tensor(t)
        if t.is_cuda:  # already on the device: used in place
            t = t.to(dtype).reshape(shape).contiguous()
            return ("dev", t)
        t = t.to(dtype).reshape(shape).contiguous()
        off = (self.total + 15) & ~15
        self.total = off + t.numel() * t.element_size()
        self.items.append((off, t))
        return ("pack", off)

    def upload(self, stream) -> List[Any]:
        if not self.items:
            return []
        pool = self._pools.setdefault((self.device.type, self.device.index), [])
        slot = None
        for i, (buf, ev) in enumerate(pool):
            if buf.numel() >= self.total and ev.query():
                slot = i
                break
        if slot is None:
            pool.append((torch.empty(max(self.total, 1 << 16), dtype=torch.uint8, pin_memory=True), torch.cuda.Event()))
            slot = len(pool) - 1
        buf, ev = pool[slot]
        for off, t in self.items:
            n = t.numel() * t.element_size()
            buf[off:off + n].copy_(t.reshape(-1).view(torch.uint8))
        self.dev = torch.empty(self.total, dtype=torch.uint8, device=self.device)
        self.dev.copy_(buf[:self.total], non_blocking=True)
        ev.record(stream)
        return [self.dev]

    def ptr(self, token) -> int:
        kind, v = token
        return v.data_ptr() if kind == "dev" else self.dev.data_ptr() + v

    @staticmethod
    def keep(token) -> List[Any]:
        return [token[1]] if token[0] == "dev" else []


def host_cos_sin(angles: torch.Tensor) -> Optional[torch.Tensor]:
    """cos/sin evaluated exactly like the reference does on the host (Affine2d.trs, affine2d.py:46-47); None when the
    angles already live on the device (the kernel then uses the correctly rounded values)."""
    if angles.is_cuda:
        return None
    a = angles.to(torch.float32)
    return torch.stack([torch.cos(a), torch.sin(a)], dim=-1)


class _ImageSource:
    """Single-channel uint8 sources of one image field: stacked tensor or ragged list -> kernel descriptors."""

    def __init__(self, value, batched: bool, device):
        self.keep: List[Any] = []
        self.table = None
        self.uniform = N.Src()
        self.stride = 0
        if isinstance(value, (list, tuple)):
            imgs = [self._hw(v) for v in value]
            tab = np.zeros((len(imgs), 6), dtype=np.int32)
            ptrs = tab[:, :2].view(np.int64)
            for i, im in enumerate(imgs):
                _require_device_visible(im, "image")
                ptrs[i, 0] = im.data_ptr()
                tab[i, 2], tab[i, 3], tab[i, 4] = im.shape[1], im.shape[0], im.stride(0)
            self.keep += imgs
            self.table = torch.from_numpy(tab).to(device, non_blocking=True)
            self.n = len(imgs)
            self.wh = None
            return
        v = value if batched else value[None]
        _require_device_visible(v, "image")
        if v.dtype != torch.uint8:
            raise TypeError(f"image must be uint8, got {v.dtype}")
        if v.dim() == 4:
            if v.shape[-1] == 1:
                v = v[..., 0]
            elif v.shape[1] == 1:
                v = v[:, 0]
            else:
                raise N.NativeError("only single-channel images are supported on this path")
        assert v.dim() == 3, f"bad image shape {tuple(value.shape)}"
        if v.stride(2) != 1:
            v = v.contiguous()
        self.keep.append(v)
        self.uniform = N.Src(v.data_ptr(), v.shape[2], v.shape[1], v.stride(1), 0)
        self.stride = v.stride(0)
        self.n = v.shape[0]
        self.wh = (v.shape[2], v.shape[1])

    @staticmethod
    def _hw(t: torch.Tensor) -> torch.Tensor:
        if t.dtype != torch.uint8:
            raise TypeError(f"image must be uint8, got {t.dtype}")
        if t.dim() == 3:
            if t.shape[-1] == 1:
                t = t[..., 0]
            elif t.shape[0] == 1:
                t = t[0]
            else:
                raise N.NativeError("only single-channel images are supported on this path")
        if t.stride(1) != 1:
            t = t.contiguous()
        return t


WORKSPACE_SIDE = 512  # widest rotated canvas kept in scratch (larger ones are produced row by row)
_WORKSPACES: Dict[Any, torch.Tensor] = {}


def _workspace(device, B: int):
    """Scratch for the rotated canvases of one launch (B200AugFusedArgs.workspace): one region per sample, cached per
    device and grown on demand.  Launches on one stream reuse it safely; concurrent streams must not share it."""
    stride = int(N.lib.b200aug_workspace_stride(WORKSPACE_SIDE))
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < B * stride:
        ws = torch.empty(B * stride, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws, stride


_PLAN_BUFFERS: Dict[Any, torch.Tensor] = {}

DOWNFILTER_CODES = {None: N.DOWN_AREA, "area": N.DOWN_AREA, "gaussian": N.DOWN_GAUSSIAN, "hamming": N.DOWN_HAMMING}
_HAMMING_TABLES: Dict[Any, Tuple[torch.Tensor, int]] = {}


def hamming_windows() -> np.ndarray:
    """float64 [32, 64]: row r holds the normalised Hamming window of 2 r + 1 taps, evaluated the way the reference's
    scipy.signal.windows.hamming does (image_geometric_cv2.py:51-57: general cosine sum over linspace(-pi, pi, n) with the
    coefficients [0.54, 1 - 0.54], then divided by its sum) -- the last bit matters: cv2 picks its column filter by whether
    the float64 window is exactly mirror symmetric."""
    out = np.zeros((32, 64), np.float64)
    coeff = (0.54, 1.0 - 0.54)
    for r in range(1, 32):
        n = 2 * r + 1
        fac = np.linspace(-np.pi, np.pi, n)
        w = np.zeros(n)
        for k, a in enumerate(coeff):
            w += a * np.cos(k * fac)
        out[r, :n] = w / np.sum(w)
    return out


def hamming_table(device) -> Tuple[torch.Tensor, int]:
    """Device copy of the window table + cv2's symmetry bits (B200AugFusedArgs.hamming_taps / hamming_sym_mask), cached."""
    key = (device.type, device.index)
    hit = _HAMMING_TABLES.get(key)
    if hit is None:
        win = np.ascontiguousarray(hamming_windows())
        taps = np.zeros(32 * 64, np.float32)
        mask = C.c_uint64(0)
        N.check(N.lib.b200aug_hamming_table(win.ctypes.data, taps.ctypes.data, C.byref(mask)), "b200aug_hamming_table")
        hit = (torch.from_numpy(taps).to(device), int(mask.value))
        _HAMMING_TABLES[key] = hit
    return hit


UPFILTER_CODES = {None: N.UP_LINEAR, "linear": N.UP_LINEAR, "cubic": N.UP_CUBIC, "lanczos": N.UP_LANCZOS}
_REMAP_TABLES: Dict[Any, torch.Tensor] = {}


def remap_tables(device) -> torch.Tensor:
    """cv2's fixed-point warpAffine tables for INTER_CUBIC and INTER_LANCZOS4 (B200AugFusedArgs.remap_tabs), cached."""
    key = (device.type, device.index)
    hit = _REMAP_TABLES.get(key)
    if hit is None:
        host = np.zeros(1024 * 16 + 1024 * 64, np.int16)
        N.check(N.lib.b200aug_remap_table(N.UP_CUBIC, host.ctypes.data), "b200aug_remap_table")
        N.check(N.lib.b200aug_remap_table(N.UP_LANCZOS, host[1024 * 16:].ctypes.data), "b200aug_remap_table")
        hit = _REMAP_TABLES[key] = torch.from_numpy(host).to(device)
    return hit


def set_downfilter(args, downfilter: Optional[str], device, upfilter: Optional[str] = None) -> List[Any]:
    """Fill B200AugFusedArgs.downfilter / upfilter (+ the Hamming / interpolation tables); returns what must outlive the launch."""
    if downfilter not in DOWNFILTER_CODES:
        raise ValueError(f"downfilter {downfilter!r}: one of 'area', 'gaussian', 'hamming'")
    if upfilter not in UPFILTER_CODES:
        raise ValueError(f"upfilter {upfilter!r}: one of 'linear', 'cubic', 'lanczos'")
    keep: List[Any] = []
    args.downfilter = DOWNFILTER_CODES[downfilter]
    if args.downfilter == N.DOWN_HAMMING:
        taps, mask = hamming_table(device)
        args.hamming_taps, args.hamming_sym_mask = taps.data_ptr(), mask
        keep.append(taps)
    args.upfilter = UPFILTER_CODES[upfilter]
    if args.upfilter != N.UP_LINEAR:
        tabs = remap_tables(device)
        args.remap_tabs = tabs.data_ptr()
        keep.append(tabs)
    return keep


def _plan_buffer(device, B: int, ow: int, oh: int):
    """Scratch for the plans + resize tables of one launch (B200AugFusedArgs.plans), cached like the canvas workspace."""
    stride = int(N.lib.b200aug_plan_stride(ow, oh))
    nbytes = int(N.lib.b200aug_plan_buffer_bytes(B, ow, oh))  # B records + the tail (work counters, per-sample flags)
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _PLAN_BUFFERS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _PLAN_BUFFERS[key] = buf
    return buf, stride


FIRST_WAVE_SAMPLES = int(os.environ.get('B200AUG_FIRST_WAVE', '148'))  # sample clusters (2 CTAs) next to the canvas workers in the first wave of a 148-SM part


ROTATED_FIRST_WAVE_COST = float(os.environ.get('B200AUG_ROT_COST', '1e9'))  # rotated samples at least this dear may start in the first wave


def launch_order(B: int, geo: Optional[GeoParams], photo: Optional[PhotoParams]) -> Optional[torch.Tensor]:
    """Launch order of the fused kernel's clusters (B200AugFusedArgs.order): most expensive photometric chains first, but
    no rotated sample in the first wave -- their canvases come from the canvas workers, which start with the first wave
    and walk the rotated samples in this same order.  Only parameters that are still on the host are looked at -- this never
    synchronises with the device; returns None when nothing distinguishes the samples.  (numpy: a few small arrays.)"""
    cost = np.zeros(B, np.float32)
    known = False
    if photo is not None and not torch.as_tensor(photo.apply).is_cuda:
        # relative to a plain crop (15 us per CTA in the r02 trace): blur +20 us, equalize +3, a noise stage +5
        ap = torch.as_tensor(photo.apply).reshape(B, N.NUM_OPS).numpy().astype(bool)
        chosen = np.zeros(N.NUM_OPS, bool)
        chosen[list(photo.order)] = True
        cost += (ap[:, 5] & chosen[5]) * np.float32(1.3) + (ap[:, 0] & chosen[0]) * np.float32(0.2)
        cost += torch.as_tensor(photo.noise_apply).reshape(B, N.NUM_NOISE).numpy().astype(np.float32).sum(1) * np.float32(0.35)
        known = True
    rot = None
    if geo is not None and not geo.angles.is_cuda:
        rot = geo.angles.reshape(B).numpy() != 0
        known = known or bool(rot.any())
    if not known or (float(cost.max()) == float(cost.min()) and (rot is None or not rot.any())):
        return None
    order = np.argsort(-cost, kind="stable")
    if rot is not None and rot.any():
        # the first wave: the dearest unrotated samples; everything else by cost behind them
        r = rot[order] & (cost[order] < ROTATED_FIRST_WAVE_COST)
        head = order[~r][:FIRST_WAVE_SAMPLES]
        taken = np.zeros(B, bool)
        taken[head] = True
        order = np.concatenate([head, order[~taken[order]]])
    return torch.from_numpy(order.astype(np.int32))


def marshal_photo(photo: PhotoParams, B: int, device, pack: Optional[_HostPack] = None):
    """PhotoParams -> the C struct B200AugPhotoParams (device arrays uploaded); returns (struct, tensors to keep alive).
    With `pack` the arrays join the caller's single parameter upload and (struct, finish) is returned instead: call
    finish(struct_in_use) after pack.upload() to fill in the pointers."""
    f32, u8 = torch.float32, torch.uint8
    p = N.PhotoParams()
    p.n_order = len(photo.order)
    for i, op in enumerate(photo.order):
        p.order[i] = int(op)
    p.clip = int(photo.clip)
    specs = ((photo.apply, u8, (B, N.NUM_OPS)), (photo.bits, torch.int32, (B,)), (photo.gamma, f32, (B,)),
             (photo.contrast, f32, (B,)), (photo.brightness, f32, (B,)), (photo.noise_apply, u8, (B, N.NUM_NOISE)))
    if pack is not None:
        toks = [pack.add(t, dt, sh) for t, dt, sh in specs]
    else:
        ap, bi, ga, co, br, na = (_dev(t, device, dt).reshape(sh) for t, dt, sh in specs)
        p.apply, p.bits, p.gamma, p.contrast, p.brightness, p.noise_apply = (t.data_ptr() for t in (ap, bi, ga, co, br, na))
    for i, s in enumerate(photo.noise_std):
        p.noise_std[i] = float(s)
    for i, c in enumerate(photo.noise_clip):
        p.noise_clip[i] = int(bool(c))
    p.seed, p.sample_offset = int(photo.seed) & (2**64 - 1), int(photo.sample_offset)
    if pack is not None:
        def finish(target):  # (a ctypes struct is copied on assignment: fill in the copy that is launched)
            target.apply, target.bits, target.gamma, target.contrast, target.brightness, target.noise_apply = (pack.ptr(t) for t in toks)
            return [x for t in toks for x in pack.keep(t)]

        return p, finish
    return p, [ap, bi, ga, co, br, na]


def fused_forward(batch: Batch, **kw) -> FusedResult:
    """Run the stages selected by `flags` on every field of `batch` in one kernel launch; returns a new Batch."""
    call = prepare_fused(batch, **kw)
    with torch.cuda.device(call.device):
        return call.launch()


def prepare_fused(batch: Batch, *, flags: int, out_size, geo: Optional[GeoParams] = None,
                  do_flip: Optional[torch.Tensor] = None, rot_dir: Optional[torch.Tensor] = None,
                  photo: Optional[PhotoParams] = None, roi_variable: str = "roi", landmark_variable: str = "pt3d_68",
                  beyond_border_shift: float = 0.3, insert_backtransform: bool = False, rowbuf_capacity: int = 0,
                  want_view_roi: bool = False, want_status: bool = False, image_key: Optional[str] = None,
                  want_trace: bool = False, use_workspace: bool = True, cluster_size: int = 0,
                  schedule: bool = True, preplan: bool = True, private_scratch: bool = False,
                  downfilter: Optional[str] = None, upfilter: Optional[str] = None) -> PreparedCall:
    """Marshal one fused call (allocate outputs, upload parameters) without launching it."""
    meta = batch.meta
    batched = meta.prefixshape != ()
    (B,) = meta.prefixshape if batched else (1,)
    ow, oh = (out_size, out_size) if isinstance(out_size, int) else tuple(out_size)
    # the device of the label fields decides (the source frames may be pinned host memory, read by the kernel in place)
    device = None
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and (v.is_cuda or as_category(meta.categories.get(k)) not in imagelike_categories):
            device = v.device
            break
    if device is None:
        device = batch.device
    if device.type != "cuda":
        raise N.NativeError(f"batch lives on {device}; the B200 path needs CUDA tensors (there is no CPU fallback)")

    args = N.FusedArgs()
    args.struct_size = C.sizeof(N.FusedArgs)
    args.batch, args.out_w, args.out_h, args.flags = B, ow, oh, flags
    args.rowbuf_capacity = rowbuf_capacity
    args.beyond_border_shift = beyond_border_shift
    args.roi_field = args.landmark_field = -1
    args.cluster_size = cluster_size
    args.warp_ctas = int(os.environ.get("B200AUG_WARP_CTAS", "0"))  # experiment knob, 0 = library default
    keep: List[Any] = []
    out_data: Dict[str, Any] = {}
    keep += set_downfilter(args, downfilter, device, upfilter)

    # ---- fields
    image_keys = [k for k in batch.keys() if as_category(meta.categories.get(k)) == FieldCategory.image]
    if image_key is not None:
        image_keys = [image_key]
    if len(image_keys) > 1:
        raise N.NativeError("one image field per call")  # TODO(next): loop over image-like fields
    nf = 0
    pending: List[Tuple[str, torch.Tensor, tuple]] = []
    for k, v in batch.items():
        cat = as_category(meta.categories.get(k))
        if cat in imagelike_categories:
            if cat == FieldCategory.semseg:
                raise N.NativeError("semseg fields are not on the B200 path")
            continue
        code = _CAT_CODE.get(cat)
        is_bt = k == "image_backtransform" and isinstance(v, torch.Tensor) and v.is_floating_point()
        if is_bt:
            # affinetrafo.py:137-147: every transform rewrites the back-transform as BT @ tr^-1, whatever its category;
            # GeneralFocusRoi(insert_backtransform=True) starts it afresh instead (geometric.py:226-227)
            if (flags & N.F_FOCUS) and insert_backtransform:
                continue
            if not (flags & (N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE)):
                out_data[k] = v
                continue
            code = N.CAT_BACKTRANSFORM
        if code is None or not isinstance(v, torch.Tensor) or not v.is_floating_point():
            out_data[k] = v
            continue
        _require_cuda(v, k)
        t = v if batched else v[None]
        t = t.to(torch.float32).contiguous()
        if is_bt:
            if tuple(t.shape[1:]) != (2, 3):
                raise ValueError(f"image_backtransform must be [..., 2, 3], got {tuple(v.shape)}")
            dim, count = 6, 1
        else:
            if t.shape[-1] not in _CAT_DIMS[code]:
                raise ValueError(f"field {k!r} of category {cat.value!r} has last dim {t.shape[-1]}")
            dim = t.shape[-1]
            count = int(np.prod(t.shape[1:-1])) if t.dim() > 2 else 1
        if nf >= N.MAX_FIELDS:
            raise N.NativeError(f"more than {N.MAX_FIELDS} label fields")
        o = torch.empty_like(t)
        f = args.fields[nf]
        f.category, f.count, f.dim, f.inp, f.out = code, count, dim, t.data_ptr(), o.data_ptr()
        if k == roi_variable:
            args.roi_field = nf
        if k == landmark_variable:
            args.landmark_field = nf
        keep += [t, o]
        pending.append((k, o, tuple(v.shape)))
        nf += 1
    args.n_fields = nf

    # ---- parameters
    f32, u8 = torch.float32, torch.uint8
    pack = _HostPack(device)
    late: List[Any] = []  # (setter, token): pointers filled in once the pack is on the device
    if flags & N.F_FOCUS:
        assert geo is not None
        late += [("scales", pack.add(geo.scales, f32, (B,))), ("angles", pack.add(geo.angles, f32, (B,))),
                 ("translations", pack.add(geo.translations, f32, (B, 2)))]
        if geo.cos_sin is not None:
            late.append(("cos_sin", pack.add(geo.cos_sin, f32, (B, 2))))
    if flags & N.F_FLIPROT:
        if do_flip is not None:
            late.append(("do_flip", pack.add(do_flip, u8, (B,))))
        if rot_dir is not None:
            late.append(("rot_dir", pack.add(rot_dir, torch.int8, (B,))))
    photo_finish = None
    if flags & N.F_PHOTOMETRIC:
        assert photo is not None
        args.photo, photo_finish = marshal_photo(photo, B, device, pack)

    # ---- image
    img_out = None
    if image_keys:
        src = _ImageSource(batch[image_keys[0]], batched, device)
        assert src.n == B, f"image count {src.n} != batch {B}"
        keep.append(src)
        if src.table is not None:
            args.src_table = src.table.data_ptr()
        else:
            args.src_uniform, args.src_stride = src.uniform, src.stride
        if not (flags & N.F_FOCUS) and src.wh is not None and src.wh != (ow, oh):
            raise ValueError(f"without the focus stage images must already be {ow}x{oh}, got {src.wh}")
        if flags & N.F_NORMALIZE:
            img_out = torch.empty((B, 1, oh, ow), dtype=f32, device=device)
            args.image_f32_out = img_out.data_ptr()
        else:
            img_out = torch.empty((B, 1, oh, ow), dtype=u8, device=device)
            args.image_u8_out = img_out.data_ptr()

    else:
        # label-only call: the label frame (normalize / flip need W and H) comes from the metadata
        w_, h_ = meta.image_wh
        args.src_uniform = N.Src(None, int(w_), int(h_), int(w_), 0)

    view_roi = tr = status = None
    if schedule and image_keys and B > 1:
        order = launch_order(B, geo if (flags & N.F_FOCUS) else None, photo if (flags & N.F_PHOTOMETRIC) else None)
        if order is not None:
            late.append(("order", pack.add(order, torch.int32, (B,))))
    # ---- one upload for all the parameter arrays
    keep += pack.upload(torch.cuda.current_stream(device))
    for name, tok in late:
        setattr(args, name, pack.ptr(tok))
        keep += pack.keep(tok)
    if photo_finish is not None:
        keep += photo_finish(args.photo)
    if (flags & N.F_FOCUS) and image_keys and use_workspace:
        if private_scratch:
            stride = int(N.lib.b200aug_workspace_stride(WORKSPACE_SIDE))
            ws = torch.empty(B * stride, dtype=torch.uint8, device=device)
        else:
            ws, stride = _workspace(device, B)
        args.workspace, args.workspace_stride = ws.data_ptr(), stride
        keep.append(ws)
    if image_keys and preplan:
        if private_scratch:  # this call's own plan records (its plan phase may overlap another call's main phase)
            pstride = int(N.lib.b200aug_plan_stride(ow, oh))
            pb = torch.empty(int(N.lib.b200aug_plan_buffer_bytes(B, ow, oh)), dtype=torch.uint8, device=device)
        else:
            pb, pstride = _plan_buffer(device, B, ow, oh)
        args.plans, args.plan_stride = pb.data_ptr(), pstride
        keep.append(pb)
    if flags & N.F_FOCUS:
        if want_view_roi:
            view_roi = torch.empty((B, 4), dtype=torch.int32, device=device)
            args.view_roi_out = view_roi.data_ptr()
        tr = torch.empty((B, 2, 3), dtype=f32, device=device)
        args.tr_out = tr.data_ptr()
        if insert_backtransform:
            bt = torch.empty((B, 2, 3), dtype=f32, device=device)
            args.backtransform_out = bt.data_ptr()
            args.flags = flags | N.F_INSERT_BACKTRANSFORM
    if want_status:
        status = torch.empty((B,), dtype=torch.int32, device=device)
        args.status_out = status.data_ptr()
    trace = None
    if want_trace:  # per-CTA timeline (profiling aid, see include/b200aug.h: trace_out)
        # B * cluster_size CTAs of the fused kernel, then up to 1024 CTAs of warp_kernel
        trace = torch.zeros((B * (cluster_size or 2) + 1024, 16), dtype=torch.int64, device=device)
        args.trace_out = trace.data_ptr()

    # ---- assemble the result (tensors are written when the call is launched)
    new_meta = meta
    for k, o, shape in pending:
        out_data[k] = o.reshape(shape)
    if image_keys:
        out_data[image_keys[0]] = img_out if batched else img_out[0]
    if (flags & N.F_FOCUS) and insert_backtransform:
        out_data["image_backtransform"] = bt if batched else bt[0]
    # keep the reference's key order
    ordered = {k: out_data[k] for k in batch.keys() if k in out_data}
    ordered.update((k, v) for k, v in out_data.items() if k not in ordered)
    res = FusedResult(Batch(new_meta, ordered), view_roi, tr, status, trace)
    res._keep = keep  # inputs must outlive the asynchronous launch
    shared = image_keys and not private_scratch and (preplan or ((flags & N.F_FOCUS) and use_workspace))
    return PreparedCall(args, device, res, keep, torch.cuda.current_stream(device).cuda_stream if shared else None)


class StatusWatch:
    """Deferred check of the per-sample status arrays (B200AugFusedArgs.status_out): a degenerate roi (empty view box)
    yields a zero-filled crop where the reference's cv2.resize raises.  The transforms copy the status of every call to
    pinned host memory behind the launch and look at the copies whose event has passed at the next call -- no
    synchronisation, the error surfaces one or two steps late; `flush()` (end of an epoch, tests) waits for all of them."""

    def __init__(self):
        self._pending = []

    def watch(self, status: torch.Tensor, what: str):
        host = torch.empty(status.shape, dtype=status.dtype, pin_memory=True)
        host.copy_(status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(status.device))
        self._pending.append((ev, host, what))
        self.poll()

    def poll(self, wait: bool = False):
        keep = []
        for ev, host, what in self._pending:
            if wait:
                ev.synchronize()
            if not ev.query():
                keep.append((ev, host, what))
                continue
            s = host.numpy()
            bad = np.nonzero(s)[0]
            if bad.size:
                self._pending = []
                raise N.NativeError(f"{what}: " + "; ".join(f"sample {i}: {STATUS_TEXT.get(int(s[i]), s[i])}" for i in bad[:8]))
        self._pending = keep

    def flush(self):
        self.poll(wait=True)


def raise_on_status(status: torch.Tensor):
    """Synchronising check of the per-sample status array (debug / tests)."""
    s = status.cpu().numpy()
    bad = np.nonzero(s)[0]
    if bad.size:
        raise N.NativeError("; ".join(f"sample {i}: {STATUS_TEXT.get(int(s[i]), s[i])}" for i in bad[:8]))


def apply_affine2d_fields(tr: torch.Tensor, fields: List[Tuple[FieldCategory, torch.Tensor]]) -> List[torch.Tensor]:
    """b200aug_apply_affine2d on [B, ...] label tensors; `tr` is [B,2,3] or [2,3] (broadcast)."""
    dev = tr.device
    _require_cuda(tr, "transform")
    tr = tr.to(torch.float32).contiguous()
    B = fields[0][1].shape[0]
    arr = (N.Field * len(fields))()
    outs, keep = [], []
    for i, (cat, t) in enumerate(fields):
        t = t.to(torch.float32).contiguous()
        o = torch.empty_like(t)
        if cat == "image_backtransform":  # (a key, not a category: BT @ tr^-1, affinetrafo.py:137-147)
            arr[i].category, arr[i].dim, arr[i].count = N.CAT_BACKTRANSFORM, 6, 1
        else:
            arr[i].category, arr[i].dim = _CAT_CODE.get(as_category(cat), N.CAT_GENERAL), t.shape[-1]
            arr[i].count = int(np.prod(t.shape[1:-1])) if t.dim() > 2 else 1
        arr[i].inp, arr[i].out = t.data_ptr(), o.data_ptr()
        outs.append(o)
        keep.append(t)
    stride = 6 if tr.dim() == 3 else 0
    with torch.cuda.device(dev):
        N.check(N.lib.b200aug_apply_affine2d(C.c_void_p(tr.data_ptr()), stride, B, len(fields), arr,
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "b200aug_apply_affine2d")
    return outs
