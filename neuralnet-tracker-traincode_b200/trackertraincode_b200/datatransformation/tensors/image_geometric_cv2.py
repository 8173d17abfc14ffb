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
    if (up, down) != ("linear", "area"):
        raise N.NativeError(f"filters ({up!r}, {down!r}) are not implemented on the B200 path (linear / area are)")


def _resample(img: torch.Tensor, new_size, view_roi: Optional[torch.Tensor], tr: Optional[torch.Tensor]) -> torch.Tensor:
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
    keep = [src]
    if view_roi is not None:
        v = view_roi.to(x.device, torch.int32).reshape(-1, 4).expand(B, 4).contiguous()
        args.explicit_view_roi = v.data_ptr()
    else:
        v = tr.to(x.device, torch.float32).reshape(-1, 2, 3).expand(B, 2, 3).contiguous()
        args.explicit_tr = v.data_ptr()
        ws, stride = E._workspace(x.device, B)
        args.workspace, args.workspace_stride = ws.data_ptr(), stride
        keep.append(ws)
    pb, pstride = E._plan_buffer(x.device, B, new_w, new_h)
    args.plans, args.plan_stride = pb.data_ptr(), pstride
    out = torch.empty((B, 1, new_h, new_w), dtype=torch.uint8, device=x.device)
    args.image_u8_out = out.data_ptr()
    with torch.cuda.device(x.device):
        N.check(N.lib.b200aug_fused_forward(C.byref(args), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                "b200aug_fused_forward")
    out._b200aug_keep = (keep, v, pb)  # inputs must outlive the asynchronous launch
    return out[0] if single else out


def croprescale_image_cv2(img: torch.Tensor, roi: torch.Tensor, new_size, downfilter: Optional[DownFilters] = None,
                          upfilter: Optional[UpFilters] = None) -> torch.Tensor:
    """Zero-padded integer crop + cv2.resize (image_geometric_cv2.py:138-155)."""
    _check_filters(downfilter, upfilter)
    return _resample(img, new_size, torch.as_tensor(roi), None)


def affine_transform_image_cv2(img: torch.Tensor, tr: Affine2d, new_size, downfilter: Optional[DownFilters] = None,
                               upfilter: Optional[UpFilters] = None) -> torch.Tensor:
    """Anti-aliased warpAffine (image_geometric_cv2.py:85-135)."""
    _check_filters(downfilter, upfilter)
    return _resample(img, new_size, None, tr.tensor())
