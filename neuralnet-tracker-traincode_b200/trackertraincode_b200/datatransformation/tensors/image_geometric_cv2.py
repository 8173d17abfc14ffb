"""croprescale_image_cv2 / affine_transform_image_cv2 on device images, tensors/image_geometric_cv2.py:85-155.

Same signatures; the arithmetic of cv2.warpAffine / cv2.resize is the fused kernel's (explicit_view_roi / explicit_tr of
B200AugFusedArgs), bit-exact against OpenCV for the filters it implements.  Single-channel uint8 images, [C,H,W] or
[H,W,C], one image or a stack [B,...]."""
from __future__ import annotations

import ctypes as C
from typing import Literal, Optional, Tuple, Union

import torch

from ... import _native as N
from ...neuralnets.affine2d import Affine2d
from .. import _engine as E
from .representation import ensure_image_nchw

DownFilters = Literal["gaussian", "hamming", "area"]
UpFilters = Literal["linear", "cubic", "lanczos"]


def _extract_size_tuple(new_size: Union[int, Tuple[int, int]]):
    try:
        new_w, new_h = new_size
    except TypeError:
        new_w = new_h = int(new_size)
    return int(new_w), int(new_h)


def _check_filters(downfilter, upfilter):
    up = "linear" if upfilter is None else upfilter
    down = "area" if downfilter is None else downfilter
    if up not in ("linear", "cubic", "lanczos"):
        raise KeyError(up)
    if down not in ("area", "gaussian", "hamming"):
        raise NotImplementedError(f"Filter: {down}")
    return down, up


def _resample(img: torch.Tensor, new_size, view_roi: Optional[torch.Tensor], tr: Optional[torch.Tensor], filters) -> torch.Tensor:
    downfilter, upfilter = filters
    if not img.is_cuda:
        raise N.NativeError(f"image lives on {img.device}; the B200 path needs CUDA tensors (there is no CPU fallback)")
    new_w, new_h = _extract_size_tuple(new_size)
    x = ensure_image_nchw(img)
    single = x.dim() == 3
    x = x[None] if single else x
    src = E._ImageSource(x, True, x.device)
    B = src.n
    args = N.FusedArgs()
    args.struct_size = C.sizeof(N.FusedArgs)
    args.batch, args.out_w, args.out_h, args.flags = B, new_w, new_h, N.F_FOCUS
    args.roi_field = args.landmark_field = -1
    args.src_uniform, args.src_stride = src.uniform, src.stride
    keep = [src] + E.set_downfilter(args, downfilter, x.device, upfilter)
    if view_roi is not None:
        v = view_roi.to(x.device, torch.int32).reshape(-1, 4).expand(B, 4).contiguous()
        args.explicit_view_roi = v.data_ptr()
    else:
        v = tr.to(x.device, torch.float32).reshape(-1, 2, 3).expand(B, 2, 3).contiguous()
        args.explicit_tr = v.data_ptr()
    status = None
    if tr is not None or args.downfilter != N.DOWN_AREA:  # rotated canvases / smoothed canvases live in the workspace
        # the canvas is the box itself, or the output blown up by the transform's scale (image_geometric_cv2.py:121-122)
        if view_roi is not None:
            side = int((v[:, 2:] - v[:, :2]).max().item())
        else:
            sc = float(tr.reshape(-1, 2, 3)[:, :, :2].det().abs().sqrt().min().item())
            side = int(max(new_w, new_h) / max(sc, 1e-6)) + 2 if sc < 1.0 else max(new_w, new_h)
        if side <= E.WORKSPACE_SIDE:
            ws, stride = E._workspace(x.device, B)
        else:  # larger than the cached scratch of the training path: a buffer of this call's own
            stride = int(N.lib.b200aug_workspace_stride(min(side, 8192)))
            ws = torch.empty(B * stride, dtype=torch.uint8, device=x.device)
        args.workspace, args.workspace_stride = ws.data_ptr(), stride
        keep.append(ws)
    if args.downfilter != N.DOWN_AREA:
        status = torch.empty((B,), dtype=torch.int32, device=x.device)
        args.status_out = status.data_ptr()
    pb, pstride = E._plan_buffer(x.device, B, new_w, new_h)
    args.plans, args.plan_stride = pb.data_ptr(), pstride
    out = torch.empty((B, 1, new_h, new_w), dtype=torch.uint8, device=x.device)
    args.image_u8_out = out.data_ptr()
    with torch.cuda.device(x.device):
        N.check(N.lib.b200aug_fused_forward(C.byref(args), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                "b200aug_fused_forward")
    out._b200aug_keep = (keep, v, pb)  # inputs must outlive the asynchronous launch
    if status is not None:  # (a prefilter the kernels cannot serve leaves zeros: say so instead)
        E.raise_on_status(status)
    return out[0] if single else out


def croprescale_image_cv2(img: torch.Tensor, roi: torch.Tensor, new_size, downfilter: Optional[DownFilters] = None,
                          upfilter: Optional[UpFilters] = None) -> torch.Tensor:
    """Zero-padded integer crop + cv2.resize (image_geometric_cv2.py:138-155)."""
    return _resample(img, new_size, torch.as_tensor(roi), None, _check_filters(downfilter, upfilter))


def affine_transform_image_cv2(img: torch.Tensor, tr: Affine2d, new_size, downfilter: Optional[DownFilters] = None,
                               upfilter: Optional[UpFilters] = None) -> torch.Tensor:
    """Anti-aliased warpAffine (image_geometric_cv2.py:85-135)."""
    return _resample(img, new_size, None, tr.tensor(), _check_filters(downfilter, upfilter))
