// Device-side scalar math shared by the augmentation kernels: 2x3 similarity algebra, label transforms,
// Philox noise.  Every float op that must reproduce the reference's host arithmetic bit for bit is spelled with
// an explicit round-to-nearest intrinsic so that no mul/add pair is ever contracted into an FMA (the TU is also
// compiled with -fmad=false).  The operation order mirrors oracle/affine.py and oracle/labels.py, which are pinned
// against the reference (trackertraincode/neuralnets/affine2d.py, tensors/affinetrafo.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200aug {

struct Aff {  // [[a00 a01 a02], [a10 a11 a12]]
  float a00, a01, a02, a10, a11, a12;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ Aff aff_identity() { return Aff{1.f, 0.f, 0.f, 0.f, 1.f, 0.f}; }

// Affine2d.range_remap_2d, affine2d.py:118-133
__device__ __forceinline__ Aff aff_range_remap(float inx0, float iny0, float inx1, float iny1, float ox0, float oy0,
                                               float ox1, float oy1) {
  float sx = fdiv(sub(ox1, ox0), sub(inx1, inx0));
  float sy = fdiv(sub(oy1, oy0), sub(iny1, iny0));
  return Aff{sx, 0.f, sub(ox0, mul(inx0, sx)), 0.f, sy, sub(oy0, mul(iny0, sy))};
}

// Affine2d.__matmul__, affine2d.py:173-180.  R: torch.matmul(2x2,2x2) evaluates fma(a1, b1, a0*b0);
// T: matvecmul is unfused (a0*b0)+(a1*b1), then += T  (pinned in tests/golden).
__device__ __forceinline__ Aff aff_compose(const Aff& a, const Aff& b) {
  Aff m;
  m.a00 = __fmaf_rn(a.a01, b.a10, mul(a.a00, b.a00));
  m.a01 = __fmaf_rn(a.a01, b.a11, mul(a.a00, b.a01));
  m.a10 = __fmaf_rn(a.a11, b.a10, mul(a.a10, b.a00));
  m.a11 = __fmaf_rn(a.a11, b.a11, mul(a.a10, b.a01));
  m.a02 = add(add(mul(a.a00, b.a02), mul(a.a01, b.a12)), a.a02);
  m.a12 = add(add(mul(a.a10, b.a02), mul(a.a11, b.a12)), a.a12);
  return m;
}

__device__ __forceinline__ float aff_det(const Aff& m) { return sub(mul(m.a00, m.a11), mul(m.a01, m.a10)); }

// Affine2d.scales, affine2d.py:189-192: sqrt(sequential float32 sum of squares) / float32(sqrt(2))
__device__ __forceinline__ float aff_scales(const Aff& m) {
  float ss = add(add(add(mul(m.a00, m.a00), mul(m.a01, m.a01)), mul(m.a10, m.a10)), mul(m.a11, m.a11));
  return fdiv(__fsqrt_rn(ss), 1.41421353816986083984375f);
}

// Affine2d.inv, affine2d.py:182-186 (closed form; the reference uses LAPACK, labels only need 1e-4)
__device__ __forceinline__ Aff aff_inv(const Aff& m) {
  float d = aff_det(m);
  Aff r;
  r.a00 = fdiv(m.a11, d);
  r.a01 = fdiv(-m.a01, d);
  r.a10 = fdiv(-m.a10, d);
  r.a11 = fdiv(m.a00, d);
  r.a02 = -add(mul(r.a00, m.a02), mul(r.a01, m.a12));
  r.a12 = -add(mul(r.a10, m.a02), mul(r.a11, m.a12));
  return r;
}

// affinevecmul, math.py:17-20
__device__ __forceinline__ void aff_apply(const Aff& m, float x, float y, float& ox, float& oy) {
  ox = add(add(mul(m.a00, x), mul(m.a01, y)), m.a02);
  oy = add(add(mul(m.a10, x), mul(m.a11, y)), m.a12);
}

// Correctly rounded float cos/sin via double (equals torch's CPU value for the production angles; see oracle/affine.py)
__device__ __forceinline__ void cos_sin_rn(float angle, float& cs, float& sn) {
  double s, c;
  sincos((double)angle, &s, &c);
  cs = (float)c;
  sn = (float)s;
}

// ---------------------------------------------------------------- label transforms (tensors/affinetrafo.py)

struct AffDerived {  // per-transform scalars shared by all items of a sample
  Aff m;
  float det, sqrt_abs_det, scales, detsign, qk, qw;
};

__device__ __noinline__ AffDerived aff_derive(const Aff& m) {
  AffDerived d;
  d.m = m;
  d.det = aff_det(m);
  d.sqrt_abs_det = __fsqrt_rn(fabsf(d.det));
  d.scales = aff_scales(m);
  d.detsign = (d.det > 0.f) ? 1.f : ((d.det < 0.f) ? -1.f : 0.f);
  // transform_rot, affinetrafo.py:98-113: in-plane angle from the "y" column
  float alpha = atan2f(-m.a01, m.a11);
  float half = mul(alpha, 0.5f);
  d.qw = cosf(half);
  d.qk = mul(sinf(half), d.detsign);
  return d;
}

// transform_points, affinetrafo.py:37-58
__device__ __forceinline__ void tf_point(const AffDerived& d, float* v, int dim) {
  float x, y;
  aff_apply(d.m, v[0], v[1], x, y);
  v[0] = x;
  v[1] = y;
  if (dim == 3) v[2] = mul(d.sqrt_abs_det, v[2]);
}

// transform_coord, affinetrafo.py:89-95
__device__ __forceinline__ void tf_coord(const AffDerived& d, float* v) {
  float x, y;
  aff_apply(d.m, v[0], v[1], x, y);
  v[0] = x;
  v[1] = y;
  v[2] = mul(d.scales, v[2]);
}

// transform_roi, affinetrafo.py:75-86
__device__ __forceinline__ void tf_roi(const AffDerived& d, float* v) {
  float xs[4], ys[4];
  aff_apply(d.m, v[0], v[1], xs[0], ys[0]);
  aff_apply(d.m, v[0], v[3], xs[1], ys[1]);
  aff_apply(d.m, v[2], v[1], xs[2], ys[2]);
  aff_apply(d.m, v[2], v[3], xs[3], ys[3]);
  v[0] = fminf(fminf(xs[0], xs[1]), fminf(xs[2], xs[3]));
  v[1] = fminf(fminf(ys[0], ys[1]), fminf(ys[2], ys[3]));
  v[2] = fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3]));
  v[3] = fmaxf(fmaxf(ys[0], ys[1]), fmaxf(ys[2], ys[3]));
}

// transform_rot, affinetrafo.py:98-127 with torchquaternion.mult (torchquaternion.py:40-48), zrot = (0,0,qk,qw)
__device__ __forceinline__ void tf_quat(const AffDerived& d, float* q) {
  float vi = q[0], vj = q[1], vk = q[2], vw = q[3];
  float w = sub(mul(d.qw, vw), mul(d.qk, vk));
  float i = sub(mul(d.qw, vi), mul(d.qk, vj));
  float j = add(mul(d.qk, vi), mul(d.qw, vj));
  float k = add(mul(d.qk, vw), mul(d.qw, vk));
  q[0] = i;
  q[1] = mul(d.detsign, j);
  q[2] = mul(d.detsign, k);
  q[3] = w;
}

// Left/right permutation of the 68-landmark scheme (facemodel/keypoints68.py:7-76), generated from its structure.
__device__ __forceinline__ int flip_map68(int i) {
  if (i < 17) return 16 - i;                 // jaw
  if (i < 27) return 43 - i;                 // brows 17..26
  if (i < 31) return i;                      // nose bridge
  if (i < 36) return 66 - i;                 // nostrils 31..35
  if (i < 40) return 81 - i;                 // right eye upper lid 36..39 -> 45..42
  if (i < 42) return 87 - i;                 // 40,41 -> 47,46
  if (i < 46) return 81 - i;                 // left eye 42..45 -> 39..36
  if (i < 48) return 87 - i;                 // 46,47 -> 41,40
  if (i < 55) return 102 - i;                // outer lip top 48..54
  if (i < 60) return 114 - i;                // outer lip bottom 55..59
  if (i < 65) return 124 - i;                // inner lip top 60..64
  return 132 - i;                            // inner lip bottom 65..67
}

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller (oracle/photometric.py)

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  float u1 = (float)((xa >> 8) + 1u) * 5.9604644775390625e-08f;  // (0, 1]
  float u2 = (float)(xb >> 8) * 5.9604644775390625e-08f;         // [0, 1)
  // hardware approximations (lg2 / sqrt / sin / cos units): the noise only has to match the oracle's float32 Box-Muller
  // to well within the 1/255 pixel tolerance (worst case ~1e-4 of a grey level at the largest noise std)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

}  // namespace b200aug
