"""Device mirror of the image decoding that feeds the augmentation path (reference: trackertraincode/datasets/preprocessing.py).

`imdecode(blob, color=False)` of the reference is `cv2.imdecode(blob, 0)` on a worker's CPU core (preprocessing.py:42-54;
0.46 ms per 450 x 450 frame); here a whole batch of JPEG blobs is decoded to grayscale frames on the GPU (nvJPEG behind
b200aug_decode_jpeg_gray), ready to be handed to `FusedPoseAugmentation` as a ragged list or a stacked tensor.
Only the grayscale path exists (the pose pipeline is monochrome, dshdf5pose.py:201); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Union

import numpy as np
import torch

from .. import _native as N

Blob = Union[bytes, bytearray, memoryview, np.ndarray]


def _as_array(blob: Blob) -> np.ndarray:
    a = np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else np.ascontiguousarray(blob, dtype=np.uint8).reshape(-1)
    if a.size == 0:
        raise ValueError("empty JPEG blob")
    return a


def jpeg_size(blob: Blob):
    """(width, height, components) from the JPEG header."""
    a = _as_array(blob)
    w, h, c = C.c_int32(), C.c_int32(), C.c_int32()
    N.check(N.lib.b200aug_jpeg_info(a.ctypes.data, a.size, C.byref(w), C.byref(h), C.byref(c)), "b200aug_jpeg_info")
    return w.value, h.value, c.value


def imdecode_batch(blobs: Sequence[Blob], device="cuda", stack: bool = False) -> Union[List[torch.Tensor], torch.Tensor]:
    """Decode JPEG blobs to grayscale uint8 frames [H, W] on `device` (stream-ordered on the current stream).
    stack=True returns one [B, H, W] tensor and requires equal sizes."""
    device = torch.device(device)
    if device.type != "cuda":
        raise N.NativeError("JPEG decoding runs on a CUDA device; there is no CPU fallback on this path")
    arrs = [_as_array(b) for b in blobs]
    n = len(arrs)
    if n == 0:
        return torch.empty((0, 0, 0), dtype=torch.uint8, device=device) if stack else []
    with torch.cuda.device(device):
        sizes = [jpeg_size(a)[:2] for a in arrs]
        if stack:
            if len(set(sizes)) != 1:
                raise ValueError(f"stack=True needs equal frame sizes, got {sorted(set(sizes))}")
            w, h = sizes[0]
            whole = torch.empty((n, h, w), dtype=torch.uint8, device=device)
            frames = list(whole.unbind(0))
        else:
            frames = [torch.empty((h, w), dtype=torch.uint8, device=device) for w, h in sizes]
        data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        lengths = (C.c_size_t * n)(*[a.size for a in arrs])
        dst = (C.c_void_p * n)(*[f.data_ptr() for f in frames])
        pitch = (C.c_int32 * n)(*[f.stride(0) for f in frames])
        rc = N.lib.b200aug_decode_jpeg_gray(data, lengths, n, dst, pitch, torch.cuda.current_stream(device).cuda_stream)
        if rc != 0:
            raise N.NativeError(f"b200aug_decode_jpeg_gray: {N.lib.b200aug_strerror(rc).decode()} (nvjpeg status {N.lib.b200aug_jpeg_last_status()})")
        # the host blobs are read while the call runs (the Huffman stage is synchronous in nvJPEG's batched decode); keep
        # them alive until here
        del arrs
    return whole if stack else frames


def imdecode(blob: Blob, color: bool = False, device="cuda") -> torch.Tensor:
    """preprocessing.py:42-54 for color=False (grayscale, [H, W] uint8)."""
    if color:
        raise N.NativeError("only the grayscale decode is on the B200 path (the pose pipeline is monochrome)")
    return imdecode_batch([blob], device=device)[0]
