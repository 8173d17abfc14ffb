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


# ----------------------------------------------------------------------------- anti-alias prefilters
# trackertraincode/datatransformation/tensors/image_geometric_cv2.py:47-62 (_apply_antialias_filter): the image is smoothed
# at its own resolution and then resized with INTER_LINEAR (:76-81).

GAUSS_FRACTION_BITS = 8  # cv2.GaussianBlur on u8: taps as 8.8 fixed point (ufixedpoint16), exact integer arithmetic


def gaussian_ksize_u8(sigma: float) -> int:
    """Kernel size cv2.GaussianBlur derives from sigma for 8-bit images when ksize = (0, 0): cvRound(sigma*3*2 + 1) | 1."""
    return int(np.rint(sigma * 3 * 2 + 1)) | 1


def gaussian_kernel_fixed(sigma: float) -> np.ndarray:
    """The 8-bit fixed-point Gaussian taps of cv2.GaussianBlur(u8): exp(-x^2 / (2 sigma^2)) normalised in double, scaled by
    256 and rounded with the rounding error carried to the next tap ("error diffusion", both halves mirrored); the centre
    tap takes whatever makes the sum exactly 256.  int64[n]."""
    n = gaussian_ksize_u8(sigma)
    n2 = (n - 1) // 2
    x = np.arange(n2, dtype=np.float64) - n2
    vals = np.exp((-0.5 / (sigma * sigma)) * x * x)
    norm = 1.0 / (2.0 * vals.sum() + 1.0)
    out = np.zeros(n, np.int64)
    err, tot = 0.0, 0
    for i in range(n2):
        adj = vals[i] * norm * float(1 << GAUSS_FRACTION_BITS) + err
        v0 = int(np.rint(adj))
        err = adj - v0
        out[i] = out[n - 1 - i] = v0
        tot += v0
    out[n2] = (1 << GAUSS_FRACTION_BITS) - 2 * tot
    return out


def reflect101(idx, size: int) -> np.ndarray:
    """cv::borderInterpolate(BORDER_REFLECT_101) on an index array (folds repeatedly when the overhang exceeds the size)."""
    idx = np.asarray(idx).copy()
    if size == 1:
        return np.zeros_like(idx)
    while True:
        bad = (idx < 0) | (idx >= size)
        if not bad.any():
            return idx
        idx = np.where(idx < 0, -idx, idx)
        idx = np.where(idx >= size, 2 * size - idx - 2, idx)


def _border_index(idx, size: int, border: str) -> np.ndarray:
    return np.clip(idx, 0, size - 1) if border == "replicate" else reflect101(idx, size)


def gaussian_blur_u8(src: np.ndarray, sigma_x: float, sigma_y: float = None, border: str = "reflect101") -> np.ndarray:
    """cv2.GaussianBlur(src, (0, 0), sigmaX=sigma_x, sigmaY=sigma_y, borderType=REFLECT_101 | REPLICATE), u8 1-channel:
    rows (sigma_x taps) then columns (sigma_y taps) in integers, (sum + 2^15) >> 16 at the end."""
    assert src.dtype == np.uint8 and src.ndim == 2
    kx = gaussian_kernel_fixed(sigma_x)
    ky = gaussian_kernel_fixed(sigma_x if sigma_y is None else sigma_y)
    h, w = src.shape
    S = src.astype(np.int64)
    hp = sum(kx[i] * S[:, _border_index(np.arange(w) + i - len(kx) // 2, w, border)] for i in range(len(kx)))
    v = sum(ky[i] * hp[_border_index(np.arange(h) + i - len(ky) // 2, h, border)] for i in range(len(ky)))
    return np.clip((v + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


# What the reference's gaussian down-filter really computes.  image_geometric_cv2.py:50 calls
#   cv2.GaussianBlur(img, (0, 0), ks, ks, cv2.BORDER_REPLICATE)
# positionally, and the Python signature is GaussianBlur(src, ksize, sigmaX[, dst[, sigmaY[, borderType]]]): the second `ks`
# lands in `dst` (ignored) and BORDER_REPLICATE (= 1) in sigmaY.  So: sigmaX = 0.5 / scale, sigmaY = 1.0 (7 taps), default
# border (REFLECT_101).  Results "identical to the reference's" means reproducing exactly that.
REFERENCE_GAUSSIAN_SIGMA_Y = 1.0
REFERENCE_GAUSSIAN_BORDER = "reflect101"


def hamming_kernel(scale_factor: float) -> np.ndarray:
    """The normalised Hamming window of image_geometric_cv2.py:51-57 (float64[n], n odd >= 3 for scale_factor < 1):
    scipy.signal.windows.hamming(n) = 0.54 - 0.46 cos(2 pi i / (n - 1)), evaluated the way scipy does
    (0.54 + 0.46 cos(linspace(-pi, pi, n))) because the last bit decides whether cv2 sees a symmetric kernel."""
    ks = 1.0 / scale_factor
    n = max(1, round(ks * 2 + 1))
    n = n if (n & 1) else n + 1
    if n == 1:
        return np.ones(1)
    fac = np.linspace(-np.pi, np.pi, n)
    w = np.zeros(n)
    for k, a in enumerate((0.54, 1.0 - 0.54)):  # (general_hamming passes [alpha, 1 - alpha])
        w += a * np.cos(k * fac)
    return w / np.sum(w)


def _fma32(a, b, c):
    """float32 fused multiply-add of arrays, exactly rounded: the product of two float32 is exact in float64, the sum is
    rounded once to 53 bits and once more to 24 -- the rare double-rounding cases (the float64 sum sits exactly on a float32
    rounding boundary) are redone in rational arithmetic."""
    from fractions import Fraction
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    c = np.asarray(c, np.float32)
    a, b, c = np.broadcast_arrays(a, b, c)
    s = a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)
    out = s.astype(np.float32)
    risky = (s.view(np.uint64) & np.uint64((1 << 29) - 1)) == np.uint64(1 << 28)
    if risky.any():
        out = out.copy()
        for idx in zip(*np.nonzero(risky)):
            exact = Fraction(float(a[idx])) * Fraction(float(b[idx])) + Fraction(float(c[idx]))
            lo = np.float32(s[idx])
            cands = [lo, np.nextafter(lo, np.float32(np.inf)), np.nextafter(lo, np.float32(-np.inf))]
            # nearest float32; ties to even mantissa
            best = min(cands, key=lambda f: (abs(Fraction(float(f)) - exact), int(np.float32(f).view(np.uint32)) & 1))
            out[idx] = best
    return out


def sep_filter_symmetric_flag(k: np.ndarray) -> bool:
    """cv::getKernelType's KERNEL_SYMMETRICAL test: exact mirror equality of the double coefficients."""
    return bool(np.array_equal(k, k[::-1]))


def sep_filter_u8(src: np.ndarray, k: np.ndarray) -> np.ndarray:
    """cv2.sepFilter2D(src, -1, k, k) for a u8 1-channel image and a float64 kernel that cv2 does not classify as a
    symmetric smoothing kernel in integers -- i.e. its float32 path (filter.simd.hpp), border BORDER_REFLECT_101:
      rows:    s = k[0] p[0];  s = fma(k[i], p[i], s)                                   i = 1 .. n-1
      columns: symmetric k:   s = k[c] h[c];  s = fma(k[c+i], h[c-i] + h[c+i], s)       i = 1 .. c
               otherwise:     s = k[0] h[0];  s = fma(k[i], h[i], s)
      out = saturate_u8(rint(s)).
    This is the arithmetic of cv2's vector loops (AVX2 / AVX-512 builds); the last `width mod V` columns of an image go
    through cv2's scalar tail loops, which round differently (V depends on the CPU), so against the live binary the model is
    bit-exact on the vector body and within 1 LSB on a handful of tail pixels (tests/test_oracle_cv2_model.py)."""
    assert src.dtype == np.uint8 and src.ndim == 2
    kf = np.asarray(k, np.float64).astype(np.float32)
    n, c = len(kf), len(kf) // 2
    h, w = src.shape

    S = src.astype(np.float32)
    cols = [reflect101(np.arange(w) + i - c, w) for i in range(n)]
    hp = kf[0] * S[:, cols[0]]
    for i in range(1, n):
        hp = _fma32(kf[i], S[:, cols[i]], hp)
    rows = [reflect101(np.arange(h) + i - c, h) for i in range(n)]
    if sep_filter_symmetric_flag(np.asarray(k, np.float64)):
        v = kf[c] * hp[rows[c]]
        for i in range(1, c + 1):
            v = _fma32(kf[c + i], hp[rows[c - i]] + hp[rows[c + i]], v)
    else:
        v = kf[0] * hp[rows[0]]
        for i in range(1, n):
            v = _fma32(kf[i], hp[rows[i]], v)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def antialias_prefilter_u8(img: np.ndarray, scale_factor: float, kind: str) -> np.ndarray:
    """_apply_antialias_filter (image_geometric_cv2.py:47-62) for u8 1-channel images -- as the reference's calls evaluate
    (see REFERENCE_GAUSSIAN_SIGMA_Y above)."""
    if kind == "gaussian":
        return gaussian_blur_u8(img, 0.5 / scale_factor, REFERENCE_GAUSSIAN_SIGMA_Y, REFERENCE_GAUSSIAN_BORDER)
    if kind == "hamming":
        return sep_filter_u8(img, hamming_kernel(scale_factor))
    raise NotImplementedError(kind)


# ----------------------------------------------------------------------------- the cubic / Lanczos up-filters
# trackertraincode/datatransformation/tensors/image_geometric_cv2.py:65-82 (cv2.resize with INTER_CUBIC / INTER_LANCZOS4 when a
# crop grows) and :105-119 (cv2.warpAffine with those flags when the transform up-scales).  OpenCV's own 8-bit kernels are
# fixed point and restated here bit for bit.  One caveat, checked in tests/test_oracle_upfilters.py: the x86 wheels route
# cv2.resize(INTER_CUBIC) through Intel IPP, whose proprietary cubic differs from OpenCV's by at most 1 grey level on ~4 % of
# the pixels (cv2.ipp.setUseIPP(False) gives the OpenCV kernel, and the model is bit-exact against it); INTER_LANCZOS4 and both
# warpAffine modes run OpenCV's kernels with or without IPP.

REMAP_COEF_BITS = 15
INTER_TAB_SIZE = 32


def cubic_coeffs(x) -> np.ndarray:
    """cv::interpolateCubic (A = -0.75), float32, operation for operation."""
    F = np.float32
    x, A = F(x), F(-0.75)
    c0 = ((A * (x + F(1)) - F(5) * A) * (x + F(1)) + F(8) * A) * (x + F(1)) - F(4) * A
    c1 = ((A + F(2)) * x - (A + F(3))) * x * x + F(1)
    c2 = ((A + F(2)) * (F(1) - x) - (A + F(3))) * (F(1) - x) * (F(1) - x) + F(1)
    c3 = F(1) - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], F)


_LANCZOS_CS = ((1.0, 0.0), (-0.70710678118654752440084436210485, -0.70710678118654752440084436210485), (0.0, 1.0),
               (0.70710678118654752440084436210485, -0.70710678118654752440084436210485), (-1.0, 0.0),
               (0.70710678118654752440084436210485, 0.70710678118654752440084436210485), (0.0, -1.0),
               (-0.70710678118654752440084436210485, 0.70710678118654752440084436210485))


def lanczos4_coeffs(x) -> np.ndarray:
    """cv::interpolateLanczos4: sin / cos of the first tap's phase in double, the other taps by the angle-sum table, each
    divided by its squared phase, the float32 results normalised by their float32 sum."""
    F = np.float32
    x = F(x)
    if x < np.finfo(F).eps:
        c = np.zeros(8, F)
        c[3] = 1
        return c
    y0 = -(float(x) + 3) * np.pi * 0.25
    s0, c0 = np.sin(y0), np.cos(y0)
    c, s = np.zeros(8, F), F(0)
    for i in range(8):
        y = -(float(x) + 3 - i) * np.pi * 0.25
        c[i] = F((_LANCZOS_CS[i][0] * s0 + _LANCZOS_CS[i][1] * c0) / (y * y))
        s = F(s + c[i])
    return (c * (F(1) / s)).astype(F)


_KERNELS = {"cubic": (cubic_coeffs, 4), "lanczos": (lanczos4_coeffs, 8)}


def resize_taps(ssize: int, dsize: int, kind: str):
    """cv2.resize per-axis tables for the 4- / 8-tap kernels: first-tap offsets (before the border clamp) and 11-bit taps."""
    coef, k = _KERNELS[kind]
    scale = 1.0 / (float(dsize) / float(ssize))
    ofs = np.zeros(dsize, np.int64)
    taps = np.zeros((dsize, k), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        taps[d] = np.rint(coef(f) * np.float32(1 << RESIZE_COEF_BITS)).astype(np.int64)
        ofs[d] = s - (k // 2 - 1)
    return ofs, taps


def resize_cubic_or_lanczos_u8(src: np.ndarray, dw: int, dh: int, kind: str) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_CUBIC | INTER_LANCZOS4), u8 1-channel, OpenCV's own kernel: rows into
    int32 with 11-bit taps (source indices clamped to the image); columns: Lanczos in integers, (sum + 2^21) >> 22; cubic in
    float32 the way OpenCV's vector loop does it (VResizeCubicVec_32s8u: the row sums converted to float, taps scaled by
    2^-22, s = r3 b3, s = fma(r2, b2, s), s = fma(r1, b1, s), s = fma(r0, b0, s), rint) -- its scalar loop for the last
    `width mod 16` columns is the integer formula, which differs on the odd near-tie (about one pixel in 10^5)."""
    assert src.dtype == np.uint8 and src.ndim == 2
    sh, sw = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()
    k = _KERNELS[kind][1]
    xo, xc = resize_taps(sw, dw, kind)
    yo, yc = resize_taps(sh, dh, kind)
    S = src.astype(np.int64)
    H = np.zeros((sh, dw), np.int64)
    for t in range(k):
        H += S[:, np.clip(xo + t, 0, sw - 1)] * xc[:, t][None, :]
    if kind == "cubic":
        F = np.float32
        b = yc.astype(F) * (F(1.0) / F((1 << RESIZE_COEF_BITS) * (1 << RESIZE_COEF_BITS)))
        rows = [H[np.clip(yo + t, 0, sh - 1)].astype(F) for t in range(4)]
        acc = rows[3] * b[:, 3][:, None]
        for t in (2, 1, 0):
            acc = _fma32(rows[t], np.broadcast_to(b[:, t][:, None], acc.shape), acc)
        return np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    out = np.zeros((dh, dw), np.int64)
    for t in range(k):
        out += H[np.clip(yo + t, 0, sh - 1)] * yc[:, t][:, None]
    return np.clip((out + (1 << (2 * RESIZE_COEF_BITS - 1))) >> (2 * RESIZE_COEF_BITS), 0, 255).astype(np.uint8)


_REMAP_TABS = {}


def remap_table(kind: str) -> np.ndarray:
    """cv::initInterTab2D(fixpt): int64 [32, 32, k, k] -- for every 1/32-pixel phase (fy, fx) the outer product of the 1-D
    float32 taps, scaled by 2^15 and rounded to short; if the k*k entries do not sum to 2^15 the difference goes to the
    largest (sum too small) or smallest (too large) of the four entries [k/2 .. k/2+1]^2, exactly as OpenCV picks them."""
    if kind in _REMAP_TABS:
        return _REMAP_TABS[kind]
    coef, k = _KERNELS[kind]
    one = 1 << REMAP_COEF_BITS
    t1 = np.stack([coef(np.float32(i) * np.float32(1.0 / INTER_TAB_SIZE)) for i in range(INTER_TAB_SIZE)])
    tab = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, k, k), np.int64)
    k2 = k // 2
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = (t1[i][:, None] * t1[j][None, :]).astype(np.float32)
            it = np.clip(np.rint(v * np.float32(one)).astype(np.int64), -32768, 32767)
            diff = int(it.sum()) - one
            if diff != 0:
                big, small = (k2, k2), (k2, k2)
                for a in range(k2, k2 + 2):
                    for b in range(k2, k2 + 2):
                        if it[a, b] < it[small]:
                            small = (a, b)
                        elif it[a, b] > it[big]:
                            big = (a, b)
                if diff < 0:
                    it[big] -= diff
                else:
                    it[small] -= diff
            tab[i, j] = it
    _REMAP_TABS[kind] = tab
    return tab


def warp_affine_cubic_or_lanczos_u8(src: np.ndarray, M, dw: int, dh: int, kind: str) -> np.ndarray:
    """cv2.warpAffine(src, M, (dw, dh), flags=INTER_CUBIC | INTER_LANCZOS4, borderMode=BORDER_CONSTANT, borderValue=0): the
    same 1/32-pixel fixed-point coordinates as the bilinear mode, k x k taps from remap_table, taps outside the image count
    as 0, (sum + 2^14) >> 15."""
    assert src.dtype == np.uint8 and src.ndim == 2
    k = _KERNELS[kind][1]
    tab = remap_table(kind)
    X, Y = warp_affine_fixed_coords(M, dw, dh)
    ix, iy, fx, fy = X >> INTER_BITS, Y >> INTER_BITS, X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)
    sh, sw = src.shape
    S = src.astype(np.int64)
    o = k // 2 - 1
    W = tab[fy, fx]
    out = np.zeros((dh, dw), np.int64)
    for a in range(k):
        yy = iy - o + a
        oky = (yy >= 0) & (yy < sh)
        for b in range(k):
            xx = ix - o + b
            ok = oky & (xx >= 0) & (xx < sw)
            out += np.where(ok, S[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)], 0) * W[:, :, a, b]
    return np.clip((out + (1 << (REMAP_COEF_BITS - 1))) >> REMAP_COEF_BITS, 0, 255).astype(np.uint8)
