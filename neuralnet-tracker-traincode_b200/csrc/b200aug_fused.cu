// Fused augmentation kernel for sm_100a: one CTA per sample does
//   label transforms -> crop / warp resample (OpenCV-exact) -> flip/rot90 -> normalise -> photometric chain -> whiten
// reading each source byte once from HBM and writing the float32 crop once.
//
// Pixel path (the specification is oracle/cv2_model.py, pinned bit-exact against cv2):
//   producer   one canvas row at a time into a per-warp shared-memory row buffer
//                CROP : zero-padded integer crop row (image_geometric_cv2.py:28-44), 16-byte vector loads
//                WARP : cv2.warpAffine INTER_LINEAR fixed point (1/32 px coordinates, 15-bit weights)
//   consumer   resizes canvas -> out_w x out_h
//                AREA      cv2.resize INTER_AREA, float32 horizontal then vertical accumulation in table order
//                AREA_INT  integer-factor INTER_AREA (box sums)
//                LINEAR    cv2.resize INTER_LINEAR, 11-bit fixed point
//                COPY      canvas already has the output size
//   each warp owns a band of output rows and walks the canvas rows that feed it, so the canvas never exists in memory.
// The uint8 crop lands in a shared-memory tile (flip/rot90 applied by the store address).  The stage-1 photometric
// ops are point functions of the uint8 value, so they collapse into a 256-entry LUT per sample (equalize's histogram
// is the uint8 histogram pushed through the LUT prefix); only the 5x5 blur needs neighbours.  The output pass streams
// the tile through LUT [+blur] [+noise, Philox] [+clip] [-0.5] into coalesced float32 stores.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "b200aug.h"
#include "b200aug_math.cuh"

namespace b200aug {

constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
static_assert(NTHREADS == 256, "thread v owns photometric LUT entry v");
constexpr int RMAX = 5;            // column rounds (of 32) per group: 160 output columns share one staged segment
constexpr int ROW_SLOTS = 2;       // staged canvas rows per warp (LINEAR needs two)
constexpr int DEFAULT_ROWBUF = 1024;

enum SrcMode { SRC_CROP = 0, SRC_WARP = 1, SRC_PLAIN = 2 };
enum RsMode { RS_COPY = 0, RS_AREA = 1, RS_AREA_INT = 2, RS_LINEAR = 3 };

struct Plan {
  // source
  const uint8_t* src;
  int sw, sh, pitch;
  int src_mode, rs_mode;
  int cw, ch;          // canvas size
  int x0, y0;          // crop origin
  double mi[6];        // warp: dst->src map (cv2 inverts the forward matrix in double)
  double scale_x, scale_y;
  int iscale_x, iscale_y;
  float inv_area;
  int do_flip, rot_dir;
  int status;
  int has_t2;
  AffDerived t_half, t1, t2, t3;   // label transforms in pipeline order
  int flip_parity;                 // number of mirroring transforms is odd -> landmark permutation
  // photometric
  int n_ops;
  int ops[B200AUG_NUM_OPS];
  int blur_pos;                    // index into ops of the blur, -1 if none
  int eq_pos;                      // index into ops of equalize, -1 if none
  int eq_step0;                    // equalize degenerated (step == 0)
  int bits;
  float gamma, contrast, brightness_shift;
  int noise_on[B200AUG_NUM_NOISE];
  int any_noise;
};

struct SmemLayout {
  size_t off_tabs, off_tile, off_rowbuf, total;
  int ntab;
};

__host__ __device__ inline SmemLayout smem_layout(int ow, int oh, int cap) {
  SmemLayout L;
  size_t o = 0;
  o += (sizeof(Plan) + 15) & ~size_t(15);
  o += 256 * 4 * 4;  // lut, eq_lut (float) ; hist8, binhist (u32)
  L.off_tabs = o;
  L.ntab = ow + oh;
  o += (size_t)L.ntab * 5 * 4;
  o = (o + 15) & ~size_t(15);
  L.off_tile = o;
  o += ((size_t)ow * oh + 15) & ~size_t(15);
  L.off_rowbuf = o;
  o += (size_t)NWARPS * ROW_SLOTS * cap;
  L.total = o;
  return L;
}

struct KArgs {
  B200AugFusedArgs a;
};

__device__ __forceinline__ int rint_d2i(double v) { return __double2int_rn(v); }

// ------------------------------------------------------------------------------------------------ plan

__device__ void build_plan(const B200AugFusedArgs& a, int b, Plan& P, float box[4], bool have_box) {
  const int ow = a.out_w, oh = a.out_h;
  P.status = B200AUG_S_OK;
  B200AugSrc s;
  if (a.src_table) {
    s = a.src_table[b];
  } else {
    s = a.src_uniform;
    s.ptr += (int64_t)b * a.src_stride;
  }
  P.src = s.ptr;
  P.sw = s.width;
  P.sh = s.height;
  P.pitch = s.pitch;
  P.do_flip = (a.flags & B200AUG_F_FLIPROT) && a.do_flip ? (a.do_flip[b] != 0) : 0;
  P.rot_dir = (a.flags & B200AUG_F_FLIPROT) && a.rot_dir ? (int)a.rot_dir[b] : 0;

  P.t_half = aff_derive(Aff{1.f, 0.f, 0.5f, 0.f, 1.f, 0.5f});
  Aff t1 = aff_identity();
  int W = P.sw, H = P.sh;  // label coordinate frame before normalisation

  if (a.flags & B200AUG_F_FOCUS) {
    // GeneralFocusRoi._compute_view_roi, geometric.py:135-156 (float32 elementwise, op for op)
    float bx0 = box[0], by0 = box[1], bx1 = box[2], by1 = box[3];
    float f = a.scales[b], rx = a.translations[2 * b], ry = a.translations[2 * b + 1];
    float bw = sub(bx1, bx0), bh = sub(by1, by0);
    float cx = mul(0.5f, add(bx1, bx0)), cy = mul(0.5f, add(by1, by0));
    float size = mul(fmaxf(bw, bh), f);
    float bbs = a.beyond_border_shift;
    float wx = add(mul(0.5f, fabsf(sub(size, bw))), mul(bbs, fminf(size, bw)));
    float wy = add(mul(0.5f, fabsf(sub(size, bh))), mul(bbs, fminf(size, bh)));
    float tx = mul(wx, rx), ty = mul(wy, ry);
    float hs = mul(size, 0.5f);
    // torch.round (half to even) -> int32, geometric.py:205
    int vx0 = (int)rintf(add(sub(cx, hs), tx)), vy0 = (int)rintf(add(sub(cy, hs), ty));
    int vx1 = (int)rintf(add(add(cx, hs), tx)), vy1 = (int)rintf(add(add(cy, hs), ty));
    if (a.view_roi_out) {
      int32_t* o = a.view_roi_out + 4 * (size_t)b;
      o[0] = vx0; o[1] = vy0; o[2] = vx1; o[3] = vy1;
    }
    // geometric.py:159-178: tr = (denorm @ rot @ norm) @ range_remap(view -> [0,out])
    Aff tr_roi = aff_range_remap((float)vx0, (float)vy0, (float)vx1, (float)vy1, 0.f, 0.f, (float)ow, (float)oh);
    Aff nrm = aff_range_remap(0.f, 0.f, (float)ow, (float)oh, -1.f, -1.f, 1.f, 1.f);
    Aff den = aff_range_remap(-1.f, -1.f, 1.f, 1.f, 0.f, 0.f, (float)ow, (float)oh);
    float angle = a.angles ? a.angles[b] : 0.f;
    float cs, sn;
    if (a.cos_sin) {
      cs = a.cos_sin[2 * b];
      sn = a.cos_sin[2 * b + 1];
    } else {
      cos_sin_rn(angle, cs, sn);
    }
    Aff rot = Aff{cs, -sn, 0.f, sn, cs, 0.f};
    t1 = aff_compose(aff_compose(aff_compose(den, rot), nrm), tr_roi);

    P.x0 = vx0;
    P.y0 = vy0;
    if (angle != 0.f) {
      // affine_transform_image_cv2, image_geometric_cv2.py:85-135
      P.src_mode = SRC_WARP;
      double sf = (double)aff_scales(t1);
      Aff M;
      if (sf > 1.0) {
        M = aff_compose(t1, Aff{1.f, 0.f, 0.5f, 0.f, 1.f, 0.5f});
        P.cw = ow;
        P.ch = oh;
      } else {
        P.cw = rint_d2i((double)ow / sf);  // Python round(): half to even on doubles
        P.ch = rint_d2i((double)oh / sf);
        float sc = (float)((double)P.ch / (double)oh);
        M = aff_compose(Aff{sc, 0.f, 0.f, 0.f, sc, 0.f}, t1);
      }
      // cv2.warpAffine: invert in double (no FMA), oracle/cv2_model.py:invert_affine_f64
      double m0 = M.a00, m1 = M.a01, m2 = M.a02, m3 = M.a10, m4 = M.a11, m5 = M.a12;
      double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
      D = (D != 0.0) ? __ddiv_rn(1.0, D) : 0.0;
      double A11 = __dmul_rn(m4, D), A22 = __dmul_rn(m0, D);
      m0 = A11;
      m1 = __dmul_rn(m1, -D);
      m3 = __dmul_rn(m3, -D);
      m4 = A22;
      double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
      double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
      P.mi[0] = m0; P.mi[1] = m1; P.mi[2] = b1; P.mi[3] = m3; P.mi[4] = m4; P.mi[5] = b2;
    } else {
      P.src_mode = SRC_CROP;
      P.cw = vx1 - vx0;
      P.ch = vy1 - vy0;
    }
    W = ow;
    H = oh;
  } else {
    // no geometric stage: the source already is the crop
    P.src_mode = SRC_PLAIN;
    P.x0 = 0;
    P.y0 = 0;
    P.cw = P.sw;
    P.ch = P.sh;
  }
  P.t1 = aff_derive(t1);
  if (a.tr_out && (a.flags & B200AUG_F_FOCUS)) {
    float* o = a.tr_out + 6 * (size_t)b;
    o[0] = t1.a00; o[1] = t1.a01; o[2] = t1.a02; o[3] = t1.a10; o[4] = t1.a11; o[5] = t1.a12;
  }
  if (a.backtransform_out && (a.flags & B200AUG_F_FOCUS)) {
    Aff iv = aff_inv(t1);
    float* o = a.backtransform_out + 6 * (size_t)b;
    o[0] = iv.a00; o[1] = iv.a01; o[2] = iv.a02; o[3] = iv.a10; o[4] = iv.a11; o[5] = iv.a12;
  }

  // resize decision, image_geometric_cv2.py:65-82 + cv::resize dispatch
  if (P.cw <= 0 || P.ch <= 0) {
    P.status = B200AUG_S_EMPTY_BOX;
    P.rs_mode = RS_COPY;
  } else if (P.cw == ow && P.ch == oh) {
    P.rs_mode = RS_COPY;
  } else {
    double scale_factor = 0.5 * ((double)ow / (double)P.cw + (double)oh / (double)P.ch);
    P.scale_x = 1.0 / ((double)ow / (double)P.cw);
    P.scale_y = 1.0 / ((double)oh / (double)P.ch);
    if (scale_factor < 1.0) {
      if (P.scale_x >= 1.0 && P.scale_y >= 1.0) {
        P.iscale_x = rint_d2i(P.scale_x);
        P.iscale_y = rint_d2i(P.scale_y);
        bool fast = fabs(P.scale_x - P.iscale_x) < 2.220446049250313e-16 && fabs(P.scale_y - P.iscale_y) < 2.220446049250313e-16;
        P.rs_mode = fast ? RS_AREA_INT : RS_AREA;
        P.inv_area = (float)(1.0 / (double)(P.iscale_x * P.iscale_y));
      } else {
        P.status = B200AUG_S_UNSUPPORTED;
        P.rs_mode = RS_COPY;
      }
    } else {
      P.rs_mode = RS_LINEAR;
    }
  }

  // horizontal_flip_and_rot_90 label transform, geometric.py:242-252
  P.has_t2 = (P.do_flip || P.rot_dir != 0);
  Aff t2 = aff_identity();
  if (P.has_t2) {
    float w = (float)W, h = (float)H;
    if (P.rot_dir != 0) {
      t2 = aff_compose(t2, aff_range_remap(-1.f, -1.f, 1.f, 1.f, 0.f, 0.f, w, h));
      float cs, sn;
      cos_sin_rn((float)((double)P.rot_dir * 3.141592653589793 * 0.5), cs, sn);
      t2 = aff_compose(t2, Aff{cs, -sn, 0.f, sn, cs, 0.f});
      t2 = aff_compose(t2, aff_range_remap(0.f, 0.f, w, h, -1.f, -1.f, 1.f, 1.f));
    }
    if (P.do_flip) t2 = aff_compose(t2, aff_range_remap(0.f, 0.f, w, h, w, 0.f, 0.f, h));
  }
  P.t2 = aff_derive(t2);
  // normalize_batch label transform, normalization.py:36-40
  P.t3 = aff_derive(aff_range_remap(0.f, 0.f, (float)W, (float)H, -1.f, -1.f, 1.f, 1.f));
  int nflip = (P.t1.det < 0.f) + (P.has_t2 && P.t2.det < 0.f);
  P.flip_parity = nflip & 1;

  // photometric parameters of this sample
  P.n_ops = 0;
  P.blur_pos = -1;
  P.eq_pos = -1;
  P.eq_step0 = 0;
  P.any_noise = 0;
  if (a.flags & B200AUG_F_PHOTOMETRIC) {
    const B200AugPhotoParams& pp = a.photo;
    for (int k = 0; k < pp.n_order; ++k) {
      int op = pp.order[k];
      if (pp.apply[(size_t)b * B200AUG_NUM_OPS + op]) {
        if (op == B200AUG_OP_BLUR) P.blur_pos = P.n_ops;
        if (op == B200AUG_OP_EQUALIZE) P.eq_pos = P.n_ops;
        P.ops[P.n_ops++] = op;
      }
    }
    P.bits = pp.bits ? pp.bits[b] : 8;
    P.gamma = pp.gamma ? pp.gamma[b] : 1.f;
    P.contrast = pp.contrast ? pp.contrast[b] : 1.f;
    P.brightness_shift = pp.brightness ? sub(pp.brightness[b], 1.f) : 0.f;
    for (int s = 0; s < B200AUG_NUM_NOISE; ++s) {
      P.noise_on[s] = pp.noise_apply ? pp.noise_apply[(size_t)b * B200AUG_NUM_NOISE + s] : 0;
      P.any_noise |= P.noise_on[s];
    }
  }
}

// ------------------------------------------------------------------------------------------------ labels

__device__ __forceinline__ void transform_item(const AffDerived& d, int category, float* v, int dim) {
  switch (category) {
    case B200AUG_CAT_POINTS: tf_point(d, v, dim); break;
    case B200AUG_CAT_XYS: tf_coord(d, v); break;
    case B200AUG_CAT_ROI: tf_roi(d, v); break;
    case B200AUG_CAT_QUAT: tf_quat(d, v); break;
    default: break;
  }
}

__device__ void run_item_chain(const Plan& P, uint32_t flags, int category, float* v, int dim) {
  if ((flags & B200AUG_F_HALF_PIXEL) && (category == B200AUG_CAT_POINTS || category == B200AUG_CAT_XYS))
    transform_item(P.t_half, category, v, dim);
  if (flags & B200AUG_F_FOCUS) transform_item(P.t1, category, v, dim);
  if ((flags & B200AUG_F_FLIPROT) && P.has_t2) transform_item(P.t2, category, v, dim);
  if (flags & B200AUG_F_NORMALIZE) transform_item(P.t3, category, v, dim);
}

__device__ void transform_labels(const B200AugFusedArgs& a, const Plan& P, int b) {
  for (int f = 0; f < a.n_fields; ++f) {
    const B200AugField& F = a.fields[f];
    if (!F.out || !F.in) continue;
    const int dim = F.dim, cnt = F.count;
    const float* in = F.in + (size_t)b * cnt * dim;
    float* out = F.out + (size_t)b * cnt * dim;
    if (F.category == B200AUG_CAT_GENERAL || dim > 4) {
      if (in != out)
        for (int i = threadIdx.x; i < cnt * dim; i += NTHREADS) out[i] = in[i];
      continue;
    }
    const bool is_roi_from_lm = (a.flags & B200AUG_F_ROI_FROM_LANDMARKS) && f == a.roi_field;
    if (is_roi_from_lm) continue;  // written by warp 0 in the prologue
    for (int i = threadIdx.x; i < cnt; i += NTHREADS) {
      int si = i;
      if (F.category == B200AUG_CAT_POINTS && cnt == 68 && P.flip_parity) si = flip_map68(i);
      float v[4];
      for (int k = 0; k < dim; ++k) v[k] = in[si * dim + k];
      run_item_chain(P, a.flags, F.category, v, dim);
      for (int k = 0; k < dim; ++k) out[i * dim + k] = v[k];
    }
  }
}

// PutRoiFromLandmarks (batch/misc.py:22-25): [min_xy, max_xy] over the landmarks, by one warp.
__device__ void landmark_box(const float* pts, int cnt, int dim, const AffDerived* pre, const AffDerived* tr, float box[4]) {
  const int lane = threadIdx.x & 31;
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = lane; i < cnt; i += 32) {
    float v[3] = {pts[i * dim], pts[i * dim + 1], 0.f};
    if (pre) tf_point(*pre, v, 2);
    if (tr) tf_point(*tr, v, 2);
    mnx = fminf(mnx, v[0]); mny = fminf(mny, v[1]);
    mxx = fmaxf(mxx, v[0]); mxy = fmaxf(mxy, v[1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  box[0] = mnx; box[1] = mny; box[2] = mxx; box[3] = mxy;
}

// ------------------------------------------------------------------------------------------------ resize tables

struct Tabs {
  int* start;   // first source index
  int* n;       // AREA: tap count | bit30 has_first | bit31 has_last ; LINEAR: second source index
  float* a;     // AREA: alpha first ; LINEAR: weight0 (as int bits)
  float* b;     // AREA: alpha mid   ; LINEAR: weight1 (as int bits)
  float* c;     // AREA: alpha last
};

// cv::computeResizeAreaTab for one destination index (oracle/cv2_model.py:area_tab)
__device__ void area_tab_entry(int d, double scale, int ssize, int& start, int& nflags, float& af, float& am, float& al) {
  double f1 = d * scale;
  double f2 = f1 + scale;
  double cw = fmin(scale, (double)ssize - f1);
  int s1 = (int)ceil(f1);
  int s2 = min((int)floor(f2), ssize - 1);
  s1 = min(s1, s2);
  int n = 0;
  bool hf = (s1 - f1) > 1e-3;
  start = hf ? s1 - 1 : s1;
  if (hf) { af = (float)((s1 - f1) / cw); ++n; } else af = 0.f;
  n += max(s2 - s1, 0);
  am = (float)(1.0 / cw);
  bool hl = (f2 - s2) > 1e-3;
  if (hl) { al = (float)(fmin(fmin(f2 - s2, 1.0), cw) / cw); ++n; } else al = 0.f;
  nflags = n | (hf ? (1 << 30) : 0) | (hl ? (1u << 31) : 0);
}

// cv::resize INTER_LINEAR taps (oracle/cv2_model.py:linear_tab + border rules)
__device__ void linear_tab_entry(int d, double scale, int ssize, bool is_x, int& i0, int& i1, int& w0, int& w1) {
  float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (is_x) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    i0 = s;
    i1 = min(s + 1, ssize - 1);
  } else {
    i0 = min(max(s, 0), ssize - 1);
    i1 = min(max(s + 1, 0), ssize - 1);
  }
  w0 = (int)rintf(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  w1 = (int)rintf(__fmul_rn(f, 2048.f));
}

// ------------------------------------------------------------------------------------------------ row producers

// Stage canvas row `y`, columns [lo, hi), into `buf`; returns the byte offset of column `lo` inside buf.
__device__ __forceinline__ int produce_row(const Plan& P, int y, int lo, int hi, uint8_t* buf, int lane) {
  if (P.src_mode != SRC_WARP) {
    const int sy = P.y0 + y;
    const int sx_lo = P.x0 + lo, sx_hi = P.x0 + hi;
    const bool row_in = (sy >= 0) && (sy < P.sh);
    if (row_in && sx_lo >= 0 && sx_hi <= P.sw) {
      const uint8_t* g = P.src + (size_t)sy * P.pitch + sx_lo;
      const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
      const uint8_t* g0 = reinterpret_cast<const uint8_t*>(ga & ~uintptr_t(15));
      const int shift = (int)(ga & 15);
      const int nvec = (shift + (hi - lo) + 15) >> 4;
      // stay inside the image allocation (the last row's tail is the only place a 16-byte load could leave it)
      const uint8_t* img_end = P.src + (size_t)(P.sh - 1) * P.pitch + P.sw;
      if (g0 >= P.src && g0 + (size_t)nvec * 16 <= img_end) {
        for (int v = lane; v < nvec; v += 32) {
          uint4 q = __ldg(reinterpret_cast<const uint4*>(g0) + v);
          reinterpret_cast<uint4*>(buf)[v] = q;
        }
        return shift;
      }
      for (int i = lane; i < hi - lo; i += 32) buf[i] = __ldg(g + i);
      return 0;
    }
    for (int i = lane; i < hi - lo; i += 32) {
      int sx = sx_lo + i;
      buf[i] = (row_in && sx >= 0 && sx < P.sw) ? __ldg(P.src + (size_t)sy * P.pitch + sx) : (uint8_t)0;
    }
    return 0;
  }
  // cv2.warpAffine INTER_LINEAR / BORDER_CONSTANT(0), oracle/cv2_model.py:warp_affine_linear_u8
  const int X0 = rint_d2i(__dmul_rn(__dadd_rn(__dmul_rn(P.mi[1], (double)y), P.mi[2]), 1024.0)) + 16;
  const int Y0 = rint_d2i(__dmul_rn(__dadd_rn(__dmul_rn(P.mi[4], (double)y), P.mi[5]), 1024.0)) + 16;
  for (int i = lane; i < hi - lo; i += 32) {
    const double xd = (double)(lo + i);
    const int ad = rint_d2i(__dmul_rn(__dmul_rn(P.mi[0], xd), 1024.0));
    const int bd = rint_d2i(__dmul_rn(__dmul_rn(P.mi[3], xd), 1024.0));
    const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
    const int ix = X >> 5, iy = Y >> 5, fx = X & 31, fy = Y & 31;
    const bool r0 = (iy >= 0) && (iy < P.sh), r1 = (iy + 1 >= 0) && (iy + 1 < P.sh);
    const bool c0 = (ix >= 0) && (ix < P.sw), c1 = (ix + 1 >= 0) && (ix + 1 < P.sw);
    const uint8_t* p = P.src + (ptrdiff_t)iy * P.pitch + ix;
    const int p00 = (r0 && c0) ? __ldg(p) : 0;
    const int p01 = (r0 && c1) ? __ldg(p + 1) : 0;
    const int p10 = (r1 && c0) ? __ldg(p + P.pitch) : 0;
    const int p11 = (r1 && c1) ? __ldg(p + P.pitch + 1) : 0;
    const int top = (32 - fx) * p00 + fx * p01;
    const int bot = (32 - fx) * p10 + fx * p11;
    buf[i] = (uint8_t)(((32 - fy) * top + fy * bot + 512) >> 10);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ the kernel

__device__ __forceinline__ int tile_index(const Plan& P, int ow, int oh, int dy, int dx) {
  // geometric.py:256-264: flip(-1), then swapaxes + flip for the 90-degree rotations
  int x1 = P.do_flip ? (ow - 1 - dx) : dx;
  int y1 = dy;
  int x, y;
  if (P.rot_dir == 0) { x = x1; y = y1; }
  else if (P.rot_dir == 1) { x = ow - 1 - y1; y = x1; }
  else { x = y1; y = oh - 1 - x1; }
  return y * ow + x;
}

__device__ __forceinline__ uint8_t sat_u8_rint(float v) {
  int r = __float2int_rn(v);
  return (uint8_t)min(max(r, 0), 255);
}

// pointwise stage-1 ops [from, to) of this sample's op list applied to x (oracle/photometric.py)
__device__ float apply_point_ops(const Plan& P, float x, int from, int to, const float* eq_lut) {
  for (int k = from; k < to; ++k) {
    switch (P.ops[k]) {
      case B200AUG_OP_EQUALIZE: {
        float im = __fmul_rn(x, 255.f);
        if (P.eq_step0) x = __fdiv_rn(im, 255.f);
        else x = __fdiv_rn(eq_lut[min(max((int)im, 0), 255)], 255.f);
      } break;
      case B200AUG_OP_POSTERIZE: {
        int q = (int)__fmul_rn(x, 255.f) & 255;
        int sh = 8 - P.bits;
        x = __fdiv_rn((float)((q >> sh) << sh), 255.f);
      } break;
      case B200AUG_OP_GAMMA: x = fminf(fmaxf(powf(x, P.gamma), 0.f), 1.f); break;
      case B200AUG_OP_CONTRAST: x = fminf(fmaxf(__fmul_rn(x, P.contrast), 0.f), 1.f); break;
      case B200AUG_OP_BRIGHTNESS: x = fminf(fmaxf(__fadd_rn(x, P.brightness_shift), 0.f), 1.f); break;
      default: break;
    }
  }
  return x;
}

__device__ __forceinline__ int eq_bin(float x) {
  float im = __fmul_rn(x, 255.f);
  int i = (int)__fmul_rn(__fdiv_rn(im, 255.f), 256.f);
  return min(max(i, 0), 255);
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// value of pixel p after the LUT prefix and (if present) the blur: input to the post-blur ops
__device__ float base_value(const Plan& P, const uint8_t* tile, const float* lut, int ow, int oh, int p) {
  if (P.blur_pos < 0) return lut[tile[p]];
  const int y = p / ow, x = p - y * ow;
  // gaussian_blur2d(5, sigma 1.5), reflect border, horizontal pass then vertical pass (oracle/photometric.py)
  // float32(exp(-t^2/(2 sigma^2))) / float32 sum, identical to oracle/photometric.py:gaussian_kernel1d
  const float g[5] = {0x1.ebd752p-4f, 0x1.defcdep-3f, 0x1.2b1778p-2f, 0x1.defcdep-3f, 0x1.ebd752p-4f};
  int xs[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) xs[i] = reflect_idx(x + i - 2, ow);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint8_t* row = tile + reflect_idx(y + j - 2, oh) * ow;
    float t = __fmul_rn(g[0], lut[row[xs[0]]]);
#pragma unroll
    for (int i = 1; i < 5; ++i) t = __fadd_rn(t, __fmul_rn(g[i], lut[row[xs[i]]]));
    acc = (j == 0) ? __fmul_rn(g[0], t) : __fadd_rn(acc, __fmul_rn(g[j], t));
  }
  return acc;
}

__global__ void __launch_bounds__(NTHREADS) fused_augment_kernel(const __grid_constant__ KArgs K, int cap) {
  const B200AugFusedArgs& a = K.a;
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x;
  const int ow = a.out_w, oh = a.out_h, npix = ow * oh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SmemLayout L = smem_layout(ow, oh, cap);
  Plan& P = *reinterpret_cast<Plan*>(smem);
  float* lut = reinterpret_cast<float*>(smem + ((sizeof(Plan) + 15) & ~size_t(15)));
  float* eq_lut = lut + 256;
  unsigned* hist8 = reinterpret_cast<unsigned*>(eq_lut + 256);
  unsigned* binhist = hist8 + 256;
  Tabs T;
  T.start = reinterpret_cast<int*>(smem + L.off_tabs);
  T.n = T.start + L.ntab;
  T.a = reinterpret_cast<float*>(T.n + L.ntab);
  T.b = T.a + L.ntab;
  T.c = T.b + L.ntab;
  uint8_t* tile = smem + L.off_tile;
  uint8_t* rowbuf = smem + L.off_rowbuf + (size_t)warp * ROW_SLOTS * cap;

  // ---- prologue: warp 0 builds the plan ------------------------------------------------------------------
  if (warp == 0) {
    float box[4] = {0.f, 0.f, 0.f, 0.f};
    const bool lm = (a.flags & B200AUG_F_ROI_FROM_LANDMARKS) && a.landmark_field >= 0;
    const AffDerived half = aff_derive(Aff{1.f, 0.f, 0.5f, 0.f, 1.f, 0.5f});
    if (lm) {
      const B200AugField& F = a.fields[a.landmark_field];
      landmark_box(F.in + (size_t)b * F.count * F.dim, F.count, F.dim, (a.flags & B200AUG_F_HALF_PIXEL) ? &half : nullptr, nullptr, box);
    } else if (a.roi_field >= 0) {
      const float* r = a.fields[a.roi_field].in + 4 * (size_t)b;
      box[0] = r[0]; box[1] = r[1]; box[2] = r[2]; box[3] = r[3];
    }
    if (lane == 0) build_plan(a, b, P, box, true);
    __syncwarp();
    if (lm && a.roi_field >= 0 && a.fields[a.roi_field].out) {
      // landmarks mode: the roi label is regenerated from the transformed landmarks after the crop (pipelines.py:347-351)
      const B200AugField& F = a.fields[a.landmark_field];
      float nb[4];
      landmark_box(F.in + (size_t)b * F.count * F.dim, F.count, F.dim, (a.flags & B200AUG_F_HALF_PIXEL) ? &half : nullptr,
                   (a.flags & B200AUG_F_FOCUS) ? &P.t1 : nullptr, nb);
      if (lane == 0) {
        if ((a.flags & B200AUG_F_FLIPROT) && P.has_t2) tf_roi(P.t2, nb);
        if (a.flags & B200AUG_F_NORMALIZE) tf_roi(P.t3, nb);
        float* o = a.fields[a.roi_field].out + 4 * (size_t)b;
        o[0] = nb[0]; o[1] = nb[1]; o[2] = nb[2]; o[3] = nb[3];
      }
    }
  }
  __syncthreads();

  if (tid == 0 && a.status_out) a.status_out[b] = P.status;
  transform_labels(a, P, b);

  const bool want_image = (a.flags & B200AUG_F_NORMALIZE) ? (a.image_f32_out != nullptr) : (a.image_u8_out != nullptr);
  if (!want_image) return;

  // ---- resize tables ---------------------------------------------------------------------------------------
  const int rs = P.rs_mode;
  if (rs == RS_AREA || rs == RS_LINEAR) {
    for (int i = tid; i < ow + oh; i += NTHREADS) {
      const bool is_x = i < ow;
      const int d = is_x ? i : i - ow;
      const double sc = is_x ? P.scale_x : P.scale_y;
      const int ss = is_x ? P.cw : P.ch;
      if (rs == RS_AREA) {
        int st, nf; float af, am, al;
        area_tab_entry(d, sc, ss, st, nf, af, am, al);
        T.start[i] = st; T.n[i] = nf; T.a[i] = af; T.b[i] = am; T.c[i] = al;
      } else {
        int i0, i1, w0, w1;
        linear_tab_entry(d, sc, ss, is_x, i0, i1, w0, w1);
        T.start[i] = i0; T.n[i] = i1; T.a[i] = __int_as_float(w0); T.b[i] = __int_as_float(w1);
      }
    }
  }
  if (a.flags & B200AUG_F_PHOTOMETRIC) {
    hist8[tid] = 0;
    binhist[tid] = 0;
  }
  __syncthreads();

  // ---- resample into the uint8 tile: each warp owns a band of output rows --------------------------------
  const int dy_begin = (warp * oh) / NWARPS, dy_end = ((warp + 1) * oh) / NWARPS;
  const bool ok = (P.status == B200AUG_S_OK);
  bool overflow = false;
  for (int g0 = 0; g0 < ow; g0 += 32 * RMAX) {
    const int gcols = min(32 * RMAX, ow - g0);
    const int glast = g0 + gcols - 1;
    // canvas columns this group needs (identical for every row)
    int seg_lo, seg_hi;
    if (rs == RS_COPY) { seg_lo = g0; seg_hi = g0 + gcols; }
    else if (rs == RS_AREA) { seg_lo = T.start[g0]; seg_hi = T.start[glast] + (T.n[glast] & 0xffff); }
    else if (rs == RS_AREA_INT) { seg_lo = g0 * P.iscale_x; seg_hi = (g0 + gcols) * P.iscale_x; }
    else { seg_lo = T.start[g0]; seg_hi = T.n[glast] + 1; }
    if (!ok || seg_hi - seg_lo + 32 > cap) {
      if (ok) overflow = true;
      for (int dy = dy_begin; dy < dy_end; ++dy)
        for (int dx = g0 + lane; dx < g0 + gcols; dx += 32) tile[tile_index(P, ow, oh, dy, dx)] = 0;
      continue;
    }
    int slot_row[ROW_SLOTS] = {INT_MIN, INT_MIN};
    int slot_shift[ROW_SLOTS] = {0, 0};
    float hbuf[RMAX];  // AREA: horizontal pass of the row staged in slot 0
#pragma unroll
    for (int j = 0; j < RMAX; ++j) hbuf[j] = 0.f;

    for (int dy = dy_begin; dy < dy_end; ++dy) {
      if (rs == RS_AREA) {
        const int ys = T.start[ow + dy], ynf = T.n[ow + dy];
        const int yn = ynf & 0xffff;
        const bool yhf = ynf & (1 << 30), yhl = ynf & (1u << 31);
        const float yaf = T.a[ow + dy], yam = T.b[ow + dy], yal = T.c[ow + dy];
        float acc[RMAX];
        for (int k = 0; k < yn; ++k) {
          const int sy = ys + k;
          const float beta = (k == 0 && yhf) ? yaf : ((k == yn - 1 && yhl) ? yal : yam);
          if (slot_row[0] != sy) {
            __syncwarp();
            slot_shift[0] = produce_row(P, sy, seg_lo, seg_hi, rowbuf, lane);
            slot_row[0] = sy;
            __syncwarp();
            const uint8_t* S = rowbuf + slot_shift[0] - seg_lo;
#pragma unroll
            for (int j = 0; j < RMAX; ++j) {
              const int dx = g0 + 32 * j + lane;
              float h = 0.f;
              if (dx <= glast) {
                const int xs = T.start[dx], xnf = T.n[dx];
                const int xn = xnf & 0xffff;
                const bool xhf = xnf & (1 << 30), xhl = xnf & (1u << 31);
                const float xaf = T.a[dx], xam = T.b[dx], xal = T.c[dx];
                for (int t = 0; t < xn; ++t) {
                  const float al = (t == 0 && xhf) ? xaf : ((t == xn - 1 && xhl) ? xal : xam);
                  h = __fadd_rn(h, __fmul_rn((float)S[xs + t], al));
                }
              }
              hbuf[j] = h;
            }
          }
#pragma unroll
          for (int j = 0; j < RMAX; ++j)
            acc[j] = (k == 0) ? __fmul_rn(beta, hbuf[j]) : __fadd_rn(acc[j], __fmul_rn(beta, hbuf[j]));
        }
#pragma unroll
        for (int j = 0; j < RMAX; ++j) {
          const int dx = g0 + 32 * j + lane;
          if (dx <= glast) tile[tile_index(P, ow, oh, dy, dx)] = sat_u8_rint(acc[j]);
        }
      } else if (rs == RS_COPY) {
        __syncwarp();
        const int sh = produce_row(P, dy, seg_lo, seg_hi, rowbuf, lane);
        __syncwarp();
        for (int dx = g0 + lane; dx <= glast; dx += 32) tile[tile_index(P, ow, oh, dy, dx)] = rowbuf[sh + dx - seg_lo];
      } else if (rs == RS_AREA_INT) {
        int acc[RMAX];
#pragma unroll
        for (int j = 0; j < RMAX; ++j) acc[j] = 0;
        for (int k = 0; k < P.iscale_y; ++k) {
          __syncwarp();
          const int sh = produce_row(P, dy * P.iscale_y + k, seg_lo, seg_hi, rowbuf, lane);
          __syncwarp();
          const uint8_t* S = rowbuf + sh - seg_lo;
#pragma unroll
          for (int j = 0; j < RMAX; ++j) {
            const int dx = g0 + 32 * j + lane;
            if (dx <= glast)
              for (int t = 0; t < P.iscale_x; ++t) acc[j] += S[dx * P.iscale_x + t];
          }
        }
#pragma unroll
        for (int j = 0; j < RMAX; ++j) {
          const int dx = g0 + 32 * j + lane;
          if (dx <= glast) {
            uint8_t v = (P.iscale_x == 2 && P.iscale_y == 2) ? (uint8_t)((acc[j] + 2) >> 2)
                                                             : sat_u8_rint(__fmul_rn((float)acc[j], P.inv_area));
            tile[tile_index(P, ow, oh, dy, dx)] = v;
          }
        }
      } else {  // RS_LINEAR
        const int r0 = T.start[ow + dy], r1 = T.n[ow + dy];
        const int b0 = __float_as_int(T.a[ow + dy]), b1 = __float_as_int(T.b[ow + dy]);
        // keep the two rows in the two slots; rows advance monotonically so reuse whatever is already staged
        int s0 = (slot_row[0] == r0) ? 0 : ((slot_row[1] == r0) ? 1 : -1);
        if (s0 < 0) {
          s0 = (slot_row[0] == r1) ? 1 : 0;
          __syncwarp();
          slot_shift[s0] = produce_row(P, r0, seg_lo, seg_hi, rowbuf + s0 * cap, lane);
          slot_row[s0] = r0;
        }
        int s1 = (slot_row[0] == r1) ? 0 : ((slot_row[1] == r1) ? 1 : -1);
        if (s1 < 0) {
          s1 = 1 - s0;
          __syncwarp();
          slot_shift[s1] = produce_row(P, r1, seg_lo, seg_hi, rowbuf + s1 * cap, lane);
          slot_row[s1] = r1;
        }
        __syncwarp();
        const uint8_t* S0 = rowbuf + s0 * cap + slot_shift[s0] - seg_lo;
        const uint8_t* S1 = rowbuf + s1 * cap + slot_shift[s1] - seg_lo;
        for (int dx = g0 + lane; dx <= glast; dx += 32) {
          const int x0 = T.start[dx], x1 = T.n[dx];
          const int a0 = __float_as_int(T.a[dx]), a1 = __float_as_int(T.b[dx]);
          const int h0 = S0[x0] * a0 + S0[x1] * a1;
          const int h1 = S1[x0] * a0 + S1[x1] * a1;
          const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
          tile[tile_index(P, ow, oh, dy, dx)] = (uint8_t)min(max(v, 0), 255);
        }
      }
    }
    __syncwarp();
  }
  if (overflow && lane == 0 && a.status_out) a.status_out[b] = B200AUG_S_ROWBUF;
  __syncthreads();

  // ---- uint8 output (geometric stages only) --------------------------------------------------------------
  if (!(a.flags & B200AUG_F_NORMALIZE)) {
    uint8_t* out = a.image_u8_out + (size_t)b * npix;
    for (int p = tid; p < npix; p += NTHREADS) out[p] = tile[p];
    return;
  }

  // ---- photometric LUT ------------------------------------------------------------------------------------
  const int n_pre = (P.blur_pos >= 0) ? P.blur_pos : P.n_ops;   // ops folded into the LUT
  const bool eq_in_lut = (P.eq_pos >= 0 && P.eq_pos < n_pre);
  const bool eq_after_blur = (P.eq_pos >= 0 && P.eq_pos >= n_pre);
  float xv = __fmul_rn((float)tid, 0.00390625f);  // normalize_batch: u8 * (1/256), thread v owns LUT entry v
  if (eq_in_lut) {
    for (int p = tid; p < npix; p += NTHREADS) atomicAdd(&hist8[tile[p]], 1u);
    xv = apply_point_ops(P, xv, 0, P.eq_pos, eq_lut);
    __syncthreads();
    if (hist8[tid]) atomicAdd(&binhist[eq_bin(xv)], hist8[tid]);
    __syncthreads();
  }
  // equalize LUT from binhist (kornia _scale_channel / _build_lut, see oracle/photometric.py:equalize)
  auto build_eq_lut = [&]() {
    if (warp == 0) {
      unsigned loc[8], run = 0, last_nz = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { loc[i] = binhist[lane * 8 + i]; run += loc[i]; }
      unsigned incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      unsigned excl = incl - run;
      // last non-zero bin
      int my_last = -1;
#pragma unroll
      for (int i = 0; i < 8; ++i) if (loc[i]) my_last = lane * 8 + i;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) my_last = max(my_last, __shfl_xor_sync(0xffffffffu, my_last, o));
      last_nz = (my_last >= 0) ? binhist[my_last] : 0u;
      const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
      const unsigned step = (total - last_nz) / 255u;
      if (lane == 0) P.eq_step0 = (step == 0);
      if (step) {
        unsigned c = excl;  // cumulative count of bins < i
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int bin = lane * 8 + i;
          // lut[bin] = clamp((cumsum[bin-1] + step/2) / step), lut[0] = 0
          eq_lut[bin] = (bin == 0) ? 0.f : (float)min((c + step / 2u) / step, 255u);
          c += loc[i];
        }
      }
    }
  };
  if (eq_in_lut) {
    build_eq_lut();
    __syncthreads();
    xv = apply_point_ops(P, xv, P.eq_pos, n_pre, eq_lut);
  } else {
    xv = apply_point_ops(P, xv, 0, n_pre, eq_lut);
  }
  lut[tid] = xv;
  __syncthreads();

  if (eq_after_blur) {
    // rare: equalize sits after the blur -> histogram of the blurred (+ intermediate ops) values
    for (int p = tid; p < npix; p += NTHREADS) {
      float x = base_value(P, tile, lut, ow, oh, p);
      x = apply_point_ops(P, x, n_pre + 1, P.eq_pos, eq_lut);
      atomicAdd(&binhist[eq_bin(x)], 1u);
    }
    __syncthreads();
    build_eq_lut();
    __syncthreads();
  }

  // ---- output pass ----------------------------------------------------------------------------------------
  float* out = a.image_f32_out + (size_t)b * npix;
  const int Q = (npix + 3) >> 2;
  const bool whiten = a.flags & B200AUG_F_WHITEN;
  const bool photo = a.flags & B200AUG_F_PHOTOMETRIC;
  const bool clip = photo && a.photo.clip;
  const uint64_t sid = a.photo.sample_offset + (uint64_t)b;
  const uint2 key = make_uint2((uint32_t)a.photo.seed, (uint32_t)(a.photo.seed >> 32));
  for (int g = tid; g < Q; g += NTHREADS) {
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = g + i * Q;
      float v = 0.f;
      if (p < npix) {
        v = base_value(P, tile, lut, ow, oh, p);
        if (P.blur_pos >= 0) v = apply_point_ops(P, v, P.blur_pos + 1, P.n_ops, eq_lut);
      }
      x[i] = v;
    }
    if (P.any_noise) {
#pragma unroll
      for (int s = 0; s < B200AUG_NUM_NOISE; ++s) {
        if (!P.noise_on[s]) continue;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)s, (uint32_t)sid, 0x6E6F6973u), key);
        float z[4];
        box_muller(r.x, r.y, z[0], z[1]);
        box_muller(r.z, r.w, z[2], z[3]);
        const float sd = a.photo.noise_std[s];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = __fadd_rn(x[i], __fmul_rn(sd, z[i]));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = g + i * Q;
      if (p < npix) {
        float v = x[i];
        if (clip) v = fminf(fmaxf(v, 0.f), 1.f);
        if (whiten) v = __fsub_rn(v, 0.5f);
        out[p] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ apply_affine2d

__global__ void apply_affine2d_kernel(const float* __restrict__ tr, int64_t tr_stride, int n_fields,
                                      const B200AugField f0, const B200AugField f1, const B200AugField f2,
                                      const B200AugField f3, const B200AugField f4, const B200AugField f5,
                                      const B200AugField f6, const B200AugField f7) {
  const B200AugField fs[B200AUG_MAX_FIELDS] = {f0, f1, f2, f3, f4, f5, f6, f7};
  const int b = blockIdx.x;
  __shared__ AffDerived D;
  if (threadIdx.x == 0) {
    const float* m = tr + (size_t)b * tr_stride;
    D = aff_derive(Aff{m[0], m[1], m[2], m[3], m[4], m[5]});
  }
  __syncthreads();
  for (int f = 0; f < n_fields; ++f) {
    const B200AugField& F = fs[f];
    const int dim = F.dim, cnt = F.count;
    const float* in = F.in + (size_t)b * cnt * dim;
    float* out = F.out + (size_t)b * cnt * dim;
    if (F.category == B200AUG_CAT_GENERAL || dim > 4) {
      if (in != out)
        for (int i = threadIdx.x; i < cnt * dim; i += blockDim.x) out[i] = in[i];
      continue;
    }
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      int si = (F.category == B200AUG_CAT_POINTS && cnt == 68 && D.det < 0.f) ? flip_map68(i) : i;
      float v[4];
      for (int k = 0; k < dim; ++k) v[k] = in[si * dim + k];
      transform_item(D, F.category, v, dim);
      for (int k = 0; k < dim; ++k) out[i * dim + k] = v[k];
    }
  }
}

}  // namespace b200aug

// ================================================================================================== C ABI

using namespace b200aug;

static thread_local int g_last_cuda_error = 0;

extern "C" int b200aug_abi_version(void) { return B200AUG_ABI_VERSION; }

extern "C" const char* b200aug_strerror(int code) {
  switch (code) {
    case B200AUG_OK: return "ok";
    case B200AUG_E_INVALID_ARG: return "invalid argument";
    case B200AUG_E_UNSUPPORTED: return "unsupported request";
    case B200AUG_E_SMEM: return "row buffer capacity does not fit in shared memory";
    case B200AUG_E_CUDA: return "CUDA runtime error (see b200aug_last_cuda_error)";
    default: return "unknown error";
  }
}

extern "C" int b200aug_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" size_t b200aug_fused_smem_bytes(int out_w, int out_h, int rowbuf_capacity) {
  if (out_w <= 0 || out_h <= 0) return 0;
  int cap = rowbuf_capacity > 0 ? rowbuf_capacity : DEFAULT_ROWBUF;
  cap = (cap + 15) & ~15;
  size_t t = smem_layout(out_w, out_h, cap).total;
  return t <= 227 * 1024 ? t : 0;
}

static int check_fields(int n, const B200AugField* f) {
  if (n < 0 || n > B200AUG_MAX_FIELDS) return B200AUG_E_INVALID_ARG;
  for (int i = 0; i < n; ++i) {
    if (f[i].count <= 0 || f[i].dim <= 0 || !f[i].in || !f[i].out) return B200AUG_E_INVALID_ARG;
    const int c = f[i].category, d = f[i].dim;
    if (c == B200AUG_CAT_QUAT && d != 4) return B200AUG_E_INVALID_ARG;
    if (c == B200AUG_CAT_XYS && d != 3) return B200AUG_E_INVALID_ARG;
    if (c == B200AUG_CAT_ROI && d != 4) return B200AUG_E_INVALID_ARG;
    if (c == B200AUG_CAT_POINTS && d != 2 && d != 3) return B200AUG_E_INVALID_ARG;
    if (c == B200AUG_CAT_POINTS && f[i].count == 68 && f[i].in == f[i].out) return B200AUG_E_INVALID_ARG;
  }
  return B200AUG_OK;
}

extern "C" int b200aug_fused_forward(const B200AugFusedArgs* args, void* stream) {
  if (!args || args->struct_size != (int32_t)sizeof(B200AugFusedArgs)) return B200AUG_E_INVALID_ARG;
  const B200AugFusedArgs& a = *args;
  if (a.batch < 0 || a.out_w <= 0 || a.out_h <= 0) return B200AUG_E_INVALID_ARG;
  if (a.batch == 0) return B200AUG_OK;
  if (!a.src_table && !a.src_uniform.ptr && (a.image_u8_out || a.image_f32_out)) return B200AUG_E_INVALID_ARG;
  int rc = check_fields(a.n_fields, a.fields);
  if (rc) return rc;
  if (a.flags & B200AUG_F_FOCUS) {
    if (!a.scales || !a.translations) return B200AUG_E_INVALID_ARG;
    const bool lm = (a.flags & B200AUG_F_ROI_FROM_LANDMARKS) != 0;
    if (lm && (a.landmark_field < 0 || a.landmark_field >= a.n_fields)) return B200AUG_E_INVALID_ARG;
    if (!lm && (a.roi_field < 0 || a.roi_field >= a.n_fields)) return B200AUG_E_INVALID_ARG;
    if (a.roi_field >= a.n_fields) return B200AUG_E_INVALID_ARG;
  }
  if ((a.flags & B200AUG_F_PHOTOMETRIC)) {
    if (!(a.flags & B200AUG_F_NORMALIZE)) return B200AUG_E_INVALID_ARG;
    if (a.photo.n_order < 0 || a.photo.n_order > B200AUG_NUM_OPS || (a.photo.n_order > 0 && !a.photo.apply)) return B200AUG_E_INVALID_ARG;
    for (int k = 0; k < a.photo.n_order; ++k)
      if (a.photo.order[k] < 0 || a.photo.order[k] >= B200AUG_NUM_OPS) return B200AUG_E_INVALID_ARG;
  }
  if ((a.flags & B200AUG_F_WHITEN) && !(a.flags & B200AUG_F_NORMALIZE)) return B200AUG_E_INVALID_ARG;
  if ((a.flags & B200AUG_F_FLIPROT) && a.rot_dir && a.out_w != a.out_h) return B200AUG_E_UNSUPPORTED;
  if ((a.flags & B200AUG_F_NORMALIZE) && a.image_u8_out && !a.image_f32_out) return B200AUG_E_INVALID_ARG;

  int cap = a.rowbuf_capacity > 0 ? a.rowbuf_capacity : DEFAULT_ROWBUF;
  cap = (cap + 15) & ~15;
  const size_t smem = smem_layout(a.out_w, a.out_h, cap).total;
  if (smem > 227 * 1024) return B200AUG_E_SMEM;
  cudaError_t e = cudaFuncSetAttribute(fused_augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  KArgs K;
  K.a = a;
  fused_augment_kernel<<<a.batch, NTHREADS, smem, (cudaStream_t)stream>>>(K, cap);
  e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}

extern "C" int b200aug_apply_affine2d(const float* tr, int64_t tr_stride, int batch, int n_fields,
                                      const B200AugField* fields, void* stream) {
  if (!tr || batch < 0 || !fields || n_fields <= 0) return B200AUG_E_INVALID_ARG;
  int rc = check_fields(n_fields, fields);
  if (rc) return rc;
  if (batch == 0) return B200AUG_OK;
  B200AugField f[B200AUG_MAX_FIELDS] = {};
  for (int i = 0; i < n_fields; ++i) f[i] = fields[i];
  apply_affine2d_kernel<<<batch, 128, 0, (cudaStream_t)stream>>>(tr, tr_stride, n_fields, f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}
