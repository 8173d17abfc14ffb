"""Bit-level numpy models of the three OpenCV resamplers the reference's pixel path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  OpenCV is a third-party dependency of the reference
(`requirements.txt:6`, unpinned; opencv-python-headless 4.13.0 in this image) whose source is not under
/root/reference, so its published fixed-point / float32 algorithms are restated here and pinned bit-exact
against the live binary in tests/test_oracle_cv2_model.py.  These models are the *specification* the CUDA
kernels in neuralnet-tracker-traincode_b200/csrc implement.

Call sites in the reference:
  cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT)  trackertraincode/datatransformation/tensors/image_geometric_cv2.py:112-119,124-131
  cv2.resize(INTER_AREA | INTER_LINEAR)           trackertraincode/datatransformation/tensors/image_geometric_cv2.py:65-82
"""
from __future__ import annotations

import numpy as np

AB_BITS = 10  # coordinates carried in 1/1024 px ...
INTER_BITS = 5  # ... and rounded to 1/32 px
ROUND_DELTA = (1 << AB_BITS) // (1 << INTER_BITS) // 2  # 16
RESIZE_COEF_BITS = 11  # cv2.resize(INTER_LINEAR) on u8: 11-bit fixed-point taps


def invert_affine_f64(M) -> np.ndarray:
    """The dst->src map cv2.warpAffine derives from a forward 2x3 matrix (double precision, no FMA)."""
    M = np.asarray(M, dtype=np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2] = b1
    M[1, 2] = b2
    return M


def warp_affine_fixed_coords(M, dw: int, dh: int):
    """Integer source coordinates (in 1/32 px) for every destination pixel: returns X, Y of shape [dh, dw]."""
    Mi = invert_affine_f64(M)
    x = np.arange(dw, dtype=np.float64)
    y = np.arange(dh, dtype=np.float64)
    adelta = np.rint(Mi[0, 0] * x * (1 << AB_BITS)).astype(np.int64)
    bdelta = np.rint(Mi[1, 0] * x * (1 << AB_BITS)).astype(np.int64)
    X0 = np.rint((Mi[0, 1] * y + Mi[0, 2]) * (1 << AB_BITS)).astype(np.int64) + ROUND_DELTA
    Y0 = np.rint((Mi[1, 1] * y + Mi[1, 2]) * (1 << AB_BITS)).astype(np.int64) + ROUND_DELTA
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    return X, Y


def warp_affine_linear_u8(src: np.ndarray, M, dw: int, dh: int) -> np.ndarray:
    """cv2.warpAffine(src, M, (dw, dh), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, borderValue=0), u8 1-channel.

    Weights: the 32x32 bilinear table entries are rint((1-b)(1-a)*32768) etc. with a=fx/32, b=fy/32, which
    are the exact integers 32*(32-fx)*(32-fy) ...; they always sum to 32768, so
    (sum w_i p_i + 16384) >> 15  ==  (sum w'_i p_i + 512) >> 10  with w' = w/32.
    """
    assert src.dtype == np.uint8 and src.ndim == 2
    h, w = src.shape
    X, Y = warp_affine_fixed_coords(M, dw, dh)
    ix = X >> INTER_BITS
    iy = Y >> INTER_BITS
    fx = X & 31
    fy = Y & 31

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64)
        return np.where(ok, v, 0)

    top = (32 - fx) * tap(iy, ix) + fx * tap(iy, ix + 1)
    bot = (32 - fx) * tap(iy + 1, ix) + fx * tap(iy + 1, ix + 1)
    acc = (32 - fy) * top + fy * bot
    return ((acc + 512) >> 10).astype(np.uint8)


def area_tab(ssize: int, dsize: int):
    """OpenCV's DecimateAlpha table for one axis: list of (di, si, alpha:float32), in emission order."""
    scale = 1.0 / (float(dsize) / float(ssize))  # cv::resize derives scale as 1/inv_scale (two roundings)
    tab = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cw = min(scale, ssize - f1)
        s1 = int(np.ceil(f1))
        s2 = min(int(np.floor(f2)), ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - f1) / cw)))
        for s in range(s1, s2):
            tab.append((d, s, np.float32(1.0 / cw)))
        if f2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(f2 - s2, 1.0), cw) / cw)))
    return tab, scale


def _is_int_scale(scale: float) -> bool:
    return abs(scale - int(scale)) < np.finfo(np.float64).eps


def resize_area_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_AREA) for down-scaling (both axis scales >= 1), u8 1-channel.

    General path: float32 horizontal pass per source row (taps in table order), float32 vertical accumulation
    in source-row order, rint at the end -- all unfused mul/add like the SSE3-baseline binary.
    Integer-factor path (both scales integral): integer box sums; 2x2 is (a+b+c+d+2)>>2, other factors
    rint(float(sum) * float(1/(sx*sy))).
    """
    assert src.dtype == np.uint8 and src.ndim == 2
    sh, sw = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()
    xtab, scale_x = area_tab(sw, dw)
    ytab, scale_y = area_tab(sh, dh)
    assert scale_x >= 1 and scale_y >= 1, "INTER_AREA up-scaling falls back to the linear kernel; not modelled"
    if _is_int_scale(scale_x) and _is_int_scale(scale_y):
        ix, iy = int(scale_x), int(scale_y)
        box = src[: dh * iy, : dw * ix].astype(np.int64).reshape(dh, iy, dw, ix).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((box + 2) >> 2).astype(np.uint8)
        inv = np.float32(1.0 / (ix * iy))
        return np.clip(np.rint(box.astype(np.float32) * inv), 0, 255).astype(np.uint8)
    f32 = np.float32
    S = src.astype(f32)
    # horizontal pass for every source row: buf[sy, dx]
    buf = np.zeros((sh, dw), dtype=f32)
    for d, s, a in xtab:  # table order == accumulation order per dx
        buf[:, d] = buf[:, d] + S[:, s] * a
    out = np.zeros((dh, dw), dtype=f32)
    started = np.zeros(dh, dtype=bool)
    for d, s, b in ytab:
        if not started[d]:
            out[d] = b * buf[s]
            started[d] = True
        else:
            out[d] = out[d] + b * buf[s]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def linear_tab(ssize: int, dsize: int):
    """cv2.resize(INTER_LINEAR) u8 per-axis taps: (ofs:int, w0:int, w1:int) with 11-bit weights (x-axis flavour)."""
    scale = 1.0 / (float(dsize) / float(ssize))
    ofs = np.zeros(dsize, dtype=np.int64)
    w0 = np.zeros(dsize, dtype=np.int64)
    w1 = np.zeros(dsize, dtype=np.int64)
    frac = np.zeros(dsize, dtype=np.float32)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        ofs[d], frac[d] = s, f
    return ofs, frac, scale


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_LINEAR), u8 1-channel (the non-"EXACT" fixed-point kernel).

    x: taps reset to a single source pixel at the borders; y: rows clamped with the weights kept.
    H pass: int32 = S[sx]*a0 + S[sx+1]*a1 (a = rint(w*2048));  V pass: ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.
    """
    assert src.dtype == np.uint8 and src.ndim == 2
    sh, sw = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()
    one = 1 << RESIZE_COEF_BITS
    xo, xf, _ = linear_tab(sw, dw)
    yo, yf, _ = linear_tab(sh, dh)
    # x border rule (hal::resize): sx < 0 -> (sx, fx) = (0, 0); sx >= sw-1 -> (sw-1, 0)
    xf = xf.copy()
    lo = xo < 0
    hi = xo >= sw - 1
    xo = np.where(lo, 0, np.where(hi, sw - 1, xo))
    xf[lo | hi] = 0
    a0 = np.rint((np.float32(1.0) - xf) * np.float32(one)).astype(np.int64)
    a1 = np.rint(xf * np.float32(one)).astype(np.int64)
    b0 = np.rint((np.float32(1.0) - yf) * np.float32(one)).astype(np.int64)
    b1 = np.rint(yf * np.float32(one)).astype(np.int64)
    S = src.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    H = S[:, xo] * a0[None, :] + S[:, x1] * a1[None, :]  # [sh, dw] int
    r0 = np.clip(yo, 0, sh - 1)
    r1 = np.clip(yo + 1, 0, sh - 1)
    out = (((b0[:, None] * (H[r0] >> 4)) >> 16) + ((b1[:, None] * (H[r1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
