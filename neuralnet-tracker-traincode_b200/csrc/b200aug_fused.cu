// The augmentation kernels for sm_100a (include/b200aug.h: b200aug_fused_forward):
//   plan_kernel            one small CTA per sample: view box / transforms / cv2 resize tables / ALL label transforms
//   fused_augment_kernel   canvas workers (cv2.warpAffine of the rotated samples into an L2-resident canvas, a work-stealing
//                          role taken by some CTAs of the grid) + one 2-CTA cluster per sample:
//                          crop / canvas -> cv2.resize (OpenCV-exact) -> flip/rot90 -> normalise -> photometric chain -> whiten
// reading each source byte once from HBM and writing the float32 crop once.
//
// Pixel path (the specification is oracle/cv2_model.py, pinned bit-exact against cv2).  The "canvas" is the image
// cv2.resize sees: the zero-padded integer crop (image_geometric_cv2.py:28-44) or the output of cv2.warpAffine
// (INTER_LINEAR fixed point: 1/32 px coordinates, 15-bit weights).  A crop's canvas never exists in memory:
//   fast path  cv2.resize INTER_AREA with a non-integer factor (what training and evaluation produce), and cv2's 2-tap
//              kernels: each warp owns a band of output rows and streams the canvas rows that feed it once, in order,
//              through a per-warp cp.async ring in shared memory.  The lane keeps its columns' tap weights in registers
//              (dense, zero padded), does the float32 horizontal pass and accumulates the vertical pass in table order.
//   per-pixel  everything else (integer-factor INTER_AREA, plain copy, extreme sizes) goes through scalar_out_px(), the
//              straight restatement of the model.
// The uint8 crop lands in a shared-memory tile (flip/rot90 applied by the store address).  The stage-1 photometric
// ops are point functions of the uint8 value, so they collapse into a 256-entry LUT per sample (equalize's histogram
// is the uint8 histogram pushed through the LUT prefix); only the 5x5 blur needs neighbours.  The output pass streams
// the tile through LUT [+blur] [+noise, Philox] [+clip] [-0.5] into coalesced float32 stores; samples whose chain is a
// single point function skip tile and output pass (direct mode).
#include <cuda_runtime.h>
#include <vector>
#include <cstdlib>
#include <math.h>
#include <stdint.h>

#include "b200aug.h"
#include "b200aug_math.cuh"

namespace b200aug {

constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;
static_assert(NTHREADS == 256, "thread v owns photometric LUT entry v");
constexpr int RMAX = 4;            // column rounds (of 32) per group: 128 output columns share one staged segment
constexpr int TAIL_COLS = 8;       // a last group of at most this many columns (129 = 128 + 1) goes through the per-pixel path
                                   // instead of costing every canvas row a nearly empty round
constexpr int KMAX = 6;            // widest INTER_AREA tap count the register path unrolls (scale factors up to ~5)
constexpr int ROWBUF_SLACK = 16;   // the word-wise tap fetch may touch up to 11 bytes past the last tap
constexpr int DT_CAP = 512;        // widest warp canvas with per-column delta tables in shared memory
constexpr int DEFAULT_ROWBUF = 2800;  // per-warp staging bytes: a ring of RING_D crop-row segments in flight (up to 690 B each:
                                      // 128 output columns at scale factors up to ~5), or one staged warpAffine tile
                                      // footprint (a 32 x 16 tile rotated by up to 45 degrees: 36 rows x 64 B + shifts).
                                      // Smaller than the 3.5 KB it could use: the shared memory not taken is L1 (2 % faster)
constexpr int RING_D = 4;          // canvas rows in flight per warp (cp.async commit groups)
constexpr int ROWPROG_CAP = 96;    // canvas rows one warp can stream per band (its vertical-pass program, 8 B per row)
constexpr int DEFAULT_CLUSTER = 2;  // CTAs sharing one sample

// bytes of one ring slot for a row segment of `seg_bytes`: alignment shift (<= 15) + 16-byte granular copy + tap over-read
__host__ __device__ __forceinline__ int ring_slot_bytes(int seg_bytes) { return (seg_bytes + 15 + 15 + ROWBUF_SLACK) & ~15; }

enum SrcMode { SRC_CROP = 0, SRC_WARP = 1, SRC_PLAIN = 2 };
enum RsMode { RS_COPY = 0, RS_AREA = 1, RS_AREA_INT = 2, RS_LINEAR = 3, RS_CUBIC = 4, RS_LANCZOS = 5 };

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
  int kx;              // widest horizontal INTER_AREA tap count (filled while the tables are built)
  int lin_area;        // RS_LINEAR with cv2's INTER_AREA coefficient rule (INTER_AREA requested while an axis up-scales)
  int fin;             // final rounding of the area sum: 0 rint | 1 integer 2x2 (sum + 2) >> 2 | 2 rint(sum * inv_area)
  int has_t2;
  AffDerived t1, t2, t3;           // label transforms in pipeline order (the half-pixel offset is a constant)
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
  // rotated samples: where the canvas workers leave the cv2.warpAffine canvas (this sample's region of the workspace);
  // NULL = no scratch canvas (no workspace, or it does not fit): the per-pixel path produces canvas pixels on the fly
  uint8_t* cv_ptr;
  int cv_pitch;
  int cv_pad;
  // anti-alias prefilter (downfilter gaussian / hamming, image_geometric_cv2.py:47-62): prefilter_kernel leaves the smoothed
  // canvas at cv_ptr, the resize is then cv2's INTER_LINEAR.  0 = none, else B200AUG_DOWN_*; pf_n taps; pf_sf = the mean
  // scale factor the taps derive from
  int prefilter, pf_n;
  double pf_sf;
  // up-filters cubic / lanczos (image_geometric_cv2.py:65-82,105-119): an up-scaling cv2.warpAffine runs with INTER_CUBIC /
  // INTER_LANCZOS4 (warp_interp = B200AUG_UP_*, warp_tab = cv2's 32 x 32 table of k x k fixed-point taps for that mode);
  // an up-scaling cv2.resize is rs_mode RS_CUBIC / RS_LANCZOS
  int warp_interp, pad_up;
  const int16_t* warp_tab;
};

constexpr size_t WORKER_AREA = 2 * 512 * 8 + 1024 * 2 + 128 + 32 * 48 + 3 * 9136;  // = WK_AREA_BYTES (checked where that is defined)

struct SmemLayout {
  size_t off_tabs, off_tile, off_rowbuf, off_bars, off_prog, off_wtab, total;
  int wpad;  // columns of the dense tap-weight table (the band columns, padded to whole 128-column groups)
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
  o += (size_t)NWARPS * (cap + ROWBUF_SLACK);
  if (o < L.off_tile + WORKER_AREA) o = L.off_tile + WORKER_AREA;  // a canvas worker's tables + staged tiles alias tile + row buffers
  L.off_bars = o;
  o += 2 * sizeof(uint64_t);  // the cluster exchange barrier
  L.off_prog = o;
  o += (size_t)NWARPS * ROWPROG_CAP * sizeof(float2);
  {
    const int tail = ow % (32 * RMAX);
    const int band = (tail != 0 && tail <= TAIL_COLS) ? ow - tail : ow;
    L.wpad = (band + 32 * RMAX - 1) / (32 * RMAX) * (32 * RMAX);  // whole column groups: every lane of a group reads its entries
  }
  L.off_wtab = o;
  o += (size_t)L.wpad * (KMAX + 1) * 4;  // dense tap weights [KMAX][wpad] + first-tap offsets [wpad]
  L.total = o;
  return L;
}

struct KArgs {
  B200AugFusedArgs a;
};

__device__ __forceinline__ int rint_d2i(double v) { return __double2int_rn(v); }

// taps cv2.GaussianBlur derives from sigma for 8-bit images (ksize (0, 0))
__host__ __device__ __forceinline__ int gaussian_ksize_u8(double sigma) {
#ifdef __CUDA_ARCH__
  return __double2int_rn(sigma * 3.0 * 2.0 + 1.0) | 1;
#else
  return (int)nearbyint(sigma * 3.0 * 2.0 + 1.0) | 1;
#endif
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- bulk asynchronous copies (TMA engine, 1-D) completing on an mbarrier: source rows stream into shared memory
// while the warp computes on earlier rows.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Spin until the phase with the given parity completes (`bar32`: shared-memory address of the mbarrier).  A copy that
// never lands would hang the GPU, so the wait traps after ~1 s worth of polls instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar32, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar32), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();
  }
}

// cp.async.wait_group takes an immediate: wait until at most `n` (0..7) of this thread's commit groups are pending
__device__ __forceinline__ void cp_async_wait_pending(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}

// per-CTA timeline (profiling aid): compiled into the TRACE instantiation of the kernel only
template <bool TRACE>
__device__ __forceinline__ void trace_mark(const B200AugFusedArgs& a, int slot) {
  if (TRACE) {
    if (a.trace_out && threadIdx.x == 0) a.trace_out[(size_t)blockIdx.x * 16 + slot] = globaltimer_ns();
  }
}

// ---- thread-block clusters: the CTAs of a cluster share one sample (each resamples a band of rows and stores the
// pixels into every CTA's tile through distributed shared memory, then each finishes its share of the output)

__device__ __forceinline__ void cluster_info(uint32_t& rank, uint32_t& size) {
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(size));
}

// barrier over all threads of the cluster (release/acquire: global and distributed-shared writes before it are visible
// after it); with a single CTA this is a plain __syncthreads()
__device__ __forceinline__ void cluster_sync(uint32_t cl) {
  if (cl > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void st_cluster_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------ plan

// The plan is built by several warps at once: every warp evaluates the short common chain (view box -> focus transform,
// redundantly in all lanes, so no exchange is needed) and then one branch each -- warpAffine matrix, resize decision,
// derived label-transform scalars, flip/rot90 transform, normalisation transform, photometric parameters.  A single
// lane doing all of it serially costs ~8 us of latency per CTA.

struct PlanCore {
  B200AugSrc s;
  int do_flip, rot_dir;
  float angle;
  Aff t1;            // focus transform (identity without F_FOCUS)
  int vx0, vy0, vx1, vy1;
  int src_mode, cw, ch;
  int W, H;          // label frame before normalisation
};

__device__ PlanCore plan_core(const B200AugFusedArgs& a, int b, const float box[4]) {
  const int ow = a.out_w, oh = a.out_h;
  const bool focus = a.flags & B200AUG_F_FOCUS, fliprot = a.flags & B200AUG_F_FLIPROT;
  PlanCore c;
  // ---- loads (issued together)
  if (a.src_table) {
    c.s = a.src_table[b];
  } else {
    c.s = a.src_uniform;
    c.s.ptr += (int64_t)b * a.src_stride;
  }
  c.do_flip = (fliprot && a.do_flip) ? (a.do_flip[b] != 0) : 0;
  c.rot_dir = (fliprot && a.rot_dir) ? (int)a.rot_dir[b] : 0;
  float f = 1.f, rx = 0.f, ry = 0.f, cs = 1.f, sn = 0.f;
  c.angle = 0.f;
  const bool sampled = focus && !a.explicit_view_roi && !a.explicit_tr;
  if (sampled) {
    f = a.scales[b];
    rx = a.translations[2 * b];
    ry = a.translations[2 * b + 1];
    c.angle = a.angles ? a.angles[b] : 0.f;
    if (a.cos_sin) {
      cs = a.cos_sin[2 * b];
      sn = a.cos_sin[2 * b + 1];
    }
  }
  c.t1 = aff_identity();
  c.W = c.s.width;
  c.H = c.s.height;
  c.vx0 = c.vy0 = c.vx1 = c.vy1 = 0;
  bool warp = false;
  if (focus) {
    if (a.explicit_tr) {
      // affine_transform_image_cv2(img, tr, new_size): the caller's transform, always through cv2.warpAffine
      const float* m = a.explicit_tr + 6 * (size_t)b;
      c.t1 = Aff{m[0], m[1], m[2], m[3], m[4], m[5]};
      warp = true;
    } else {
      if (a.explicit_view_roi) {
        // croprescale_image_cv2(img, roi, new_size): the caller's integer box
        const int32_t* v = a.explicit_view_roi + 4 * (size_t)b;
        c.vx0 = v[0]; c.vy0 = v[1]; c.vx1 = v[2]; c.vy1 = v[3];
      } else {
        // GeneralFocusRoi._compute_view_roi, geometric.py:135-156 (float32 elementwise, op for op)
        float bx0 = box[0], by0 = box[1], bx1 = box[2], by1 = box[3];
        float bw = sub(bx1, bx0), bh = sub(by1, by0);
        float cx = mul(0.5f, add(bx1, bx0)), cy = mul(0.5f, add(by1, by0));
        float size = mul(fmaxf(bw, bh), f);
        float bbs = a.beyond_border_shift;
        float wx = add(mul(0.5f, fabsf(sub(size, bw))), mul(bbs, fminf(size, bw)));
        float wy = add(mul(0.5f, fabsf(sub(size, bh))), mul(bbs, fminf(size, bh)));
        float tx = mul(wx, rx), ty = mul(wy, ry);
        float hs = mul(size, 0.5f);
        // torch.round (half to even) -> int32, geometric.py:205
        c.vx0 = (int)rintf(add(sub(cx, hs), tx)); c.vy0 = (int)rintf(add(sub(cy, hs), ty));
        c.vx1 = (int)rintf(add(add(cx, hs), tx)); c.vy1 = (int)rintf(add(add(cy, hs), ty));
      }
      // geometric.py:159-178: tr = (denorm @ rot @ norm) @ range_remap(view -> [0,out])
      Aff tr_roi = aff_range_remap((float)c.vx0, (float)c.vy0, (float)c.vx1, (float)c.vy1, 0.f, 0.f, (float)ow, (float)oh);
      if (a.explicit_view_roi) {
        c.t1 = tr_roi;
      } else {
        Aff nrm = aff_range_remap(0.f, 0.f, (float)ow, (float)oh, -1.f, -1.f, 1.f, 1.f);
        Aff den = aff_range_remap(-1.f, -1.f, 1.f, 1.f, 0.f, 0.f, (float)ow, (float)oh);
        if (!a.cos_sin) cos_sin_rn(c.angle, cs, sn);
        Aff rot = Aff{cs, -sn, 0.f, sn, cs, 0.f};
        c.t1 = aff_compose(aff_compose(aff_compose(den, rot), nrm), tr_roi);
      }
      warp = c.angle != 0.f;
    }
    if (warp) {
      // affine_transform_image_cv2, image_geometric_cv2.py:85-135
      c.src_mode = SRC_WARP;
      double sf = (double)aff_scales(c.t1);
      if (sf > 1.0) {
        c.cw = ow;
        c.ch = oh;
      } else {
        c.cw = rint_d2i((double)ow / sf);  // Python round(): half to even on doubles
        c.ch = rint_d2i((double)oh / sf);
      }
    } else {
      c.src_mode = SRC_CROP;
      c.cw = c.vx1 - c.vx0;
      c.ch = c.vy1 - c.vy0;
    }
    c.W = ow;
    c.H = oh;
  } else {
    c.src_mode = SRC_PLAIN;  // no geometric stage: the source already is the crop
    c.cw = c.s.width;
    c.ch = c.s.height;
  }
  return c;
}

// branch 0: plain fields + the dst->src map of cv2.warpAffine
__device__ void plan_basic_and_warp(const B200AugFusedArgs& a, const PlanCore& c, Plan& P) {
  const int ow = a.out_w, oh = a.out_h;
  P.kx = 0;
  P.src = c.s.ptr;
  P.sw = c.s.width;
  P.sh = c.s.height;
  P.pitch = c.s.pitch;
  P.do_flip = c.do_flip;
  P.rot_dir = c.rot_dir;
  P.src_mode = c.src_mode;
  P.cw = c.cw;
  P.ch = c.ch;
  P.x0 = (c.src_mode == SRC_PLAIN) ? 0 : c.vx0;
  P.y0 = (c.src_mode == SRC_PLAIN) ? 0 : c.vy0;
  P.warp_interp = 0;
  P.pad_up = 0;
  P.warp_tab = nullptr;
  if (c.src_mode == SRC_WARP) {
    Aff M;
    if (c.cw == ow && c.ch == oh && (double)aff_scales(c.t1) > 1.0) {
      M = aff_compose(c.t1, Aff{1.f, 0.f, 0.5f, 0.f, 1.f, 0.5f});
      if (a.upfilter != B200AUG_UP_LINEAR && a.remap_tabs) {  // image_geometric_cv2.py:105-119: the up-filter is the warp's flag
        P.warp_interp = a.upfilter;
        P.warp_tab = a.remap_tabs + (a.upfilter == B200AUG_UP_CUBIC ? 0 : 1024 * 16);
      }
    } else {
      float sc = (float)((double)c.ch / (double)oh);
      M = aff_compose(Aff{sc, 0.f, 0.f, 0.f, sc, 0.f}, c.t1);
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
  }
}

// branch 1: resize decision, image_geometric_cv2.py:65-82 + cv::resize dispatch
__device__ void plan_resize(const B200AugFusedArgs& a, const PlanCore& c, Plan& P) {
  const int ow = a.out_w, oh = a.out_h;
  P.status = B200AUG_S_OK;
  P.fin = 0;
  P.lin_area = 0;
  P.prefilter = 0;
  P.pf_n = 0;
  P.pf_sf = 1.0;
  if (c.cw <= 0 || c.ch <= 0) {
    P.status = B200AUG_S_EMPTY_BOX;
    P.rs_mode = RS_COPY;
  } else if (c.cw == ow && c.ch == oh) {
    P.rs_mode = RS_COPY;
  } else {
    double scale_factor = 0.5 * ((double)ow / (double)c.cw + (double)oh / (double)c.ch);
    P.scale_x = 1.0 / ((double)ow / (double)c.cw);
    P.scale_y = 1.0 / ((double)oh / (double)c.ch);
    if (scale_factor < 1.0 && a.downfilter != B200AUG_DOWN_AREA && (a.flags & B200AUG_F_FOCUS)) {
      // _resize with downfilter gaussian / hamming (image_geometric_cv2.py:47-62,76-81): smooth at canvas resolution, then
      // INTER_LINEAR.  Tap counts: cv2.GaussianBlur's cvRound(sigma * 3 * 2 + 1) | 1 for 8-bit images, sigma = 0.5 / scale;
      // the Hamming window's round(2 / scale + 1) made odd.
      int n;
      if (a.downfilter == B200AUG_DOWN_GAUSSIAN) {
        n = gaussian_ksize_u8(0.5 / scale_factor);  // (the row kernel; the column kernel has 7 taps, see prefilter_kernel)
      } else {
        const double ks = 1.0 / scale_factor;
        n = max(1, rint_d2i(ks * 2.0 + 1.0));
        n |= 1;
      }
      P.rs_mode = RS_LINEAR;
      P.prefilter = a.downfilter;
      P.pf_n = n;
      P.pf_sf = scale_factor;
      if (n > B200AUG_PREFILTER_MAX_TAPS) P.status = B200AUG_S_UNSUPPORTED;
    } else if (scale_factor < 1.0) {
      if (P.scale_x >= 1.0 && P.scale_y >= 1.0) {
        P.iscale_x = rint_d2i(P.scale_x);
        P.iscale_y = rint_d2i(P.scale_y);
        bool fast = fabs(P.scale_x - P.iscale_x) < 2.220446049250313e-16 && fabs(P.scale_y - P.iscale_y) < 2.220446049250313e-16;
        P.rs_mode = fast ? RS_AREA_INT : RS_AREA;
        P.inv_area = (float)(1.0 / (double)(P.iscale_x * P.iscale_y));
        if (fast) P.fin = (P.iscale_x == 2 && P.iscale_y == 2) ? 1 : 2;
      } else {
        // cv::resize(INTER_AREA) with an up-scaling axis: not the area kernel but the 2-tap linear one, with the
        // "area mode" coefficient rule on BOTH axes (imgproc/resize.cpp; checked bit-exact against cv2 4.13 for mixed and
        // pure up-scaling).  Reachable with non-square outputs (the localizer's 288 x 224: view side between 224 and 288);
        // a square output never gets here, the rounded view box is at most one pixel off square.
        P.rs_mode = RS_LINEAR;
        P.lin_area = 1;
      }
    } else {
      // the up-filter (image_geometric_cv2.py:68-75)
      P.rs_mode = (a.upfilter == B200AUG_UP_CUBIC) ? RS_CUBIC : ((a.upfilter == B200AUG_UP_LANCZOS) ? RS_LANCZOS : RS_LINEAR);
    }
  }
}

// horizontal_flip_and_rot_90 label transform, geometric.py:242-252
__device__ Aff fliprot_transform(const PlanCore& c) {
  Aff t2 = aff_identity();
  float w = (float)c.W, h = (float)c.H;
  if (c.rot_dir != 0) {
    t2 = aff_compose(t2, aff_range_remap(-1.f, -1.f, 1.f, 1.f, 0.f, 0.f, w, h));
    // float32(cos), float32(sin) of float32(+-pi/2), as torch evaluates them (affine2d.py:46-47)
    const float c90 = -0x1.777a5cp-25f, s90 = (c.rot_dir > 0) ? 1.f : -1.f;
    t2 = aff_compose(t2, Aff{c90, -s90, 0.f, s90, c90, 0.f});
    t2 = aff_compose(t2, aff_range_remap(0.f, 0.f, w, h, -1.f, -1.f, 1.f, 1.f));
  }
  if (c.do_flip) t2 = aff_compose(t2, aff_range_remap(0.f, 0.f, w, h, w, 0.f, 0.f, h));
  return t2;
}

// normalize_batch label transform, normalization.py:36-40
__device__ __forceinline__ Aff normalize_transform(const PlanCore& c) {
  return aff_range_remap(0.f, 0.f, (float)c.W, (float)c.H, -1.f, -1.f, 1.f, 1.f);
}

// branch 5: photometric parameters of this sample
__device__ void plan_photo(const B200AugPhotoParams& pp, bool enabled, int b, Plan& P) {
  P.n_ops = 0;
  P.blur_pos = -1;
  P.eq_pos = -1;
  P.eq_step0 = 0;
  P.any_noise = 0;
  if (!enabled) return;
  uint8_t op_on[B200AUG_NUM_OPS] = {0, 0, 0, 0, 0, 0}, noise_on[B200AUG_NUM_NOISE] = {0, 0, 0, 0};
  if (pp.apply)
#pragma unroll
    for (int k = 0; k < B200AUG_NUM_OPS; ++k) op_on[k] = pp.apply[(size_t)b * B200AUG_NUM_OPS + k];
  if (pp.noise_apply)
#pragma unroll
    for (int k = 0; k < B200AUG_NUM_NOISE; ++k) noise_on[k] = pp.noise_apply[(size_t)b * B200AUG_NUM_NOISE + k];
  const int bits = pp.bits ? pp.bits[b] : 8;
  const float gamma = pp.gamma ? pp.gamma[b] : 1.f;
  const float contrast = pp.contrast ? pp.contrast[b] : 1.f;
  const float brightness = pp.brightness ? pp.brightness[b] : 1.f;
  for (int k = 0; k < pp.n_order; ++k) {
    const int op = pp.order[k];
    bool on = false;
#pragma unroll
    for (int q = 0; q < B200AUG_NUM_OPS; ++q) on = on || (q == op && op_on[q]);
    if (on) {
      if (op == B200AUG_OP_BLUR) P.blur_pos = P.n_ops;
      if (op == B200AUG_OP_EQUALIZE) P.eq_pos = P.n_ops;
      P.ops[P.n_ops++] = op;
    }
  }
  P.bits = bits;
  P.gamma = gamma;
  P.contrast = contrast;
  P.brightness_shift = sub(brightness, 1.f);
#pragma unroll
  for (int q = 0; q < B200AUG_NUM_NOISE; ++q) {
    P.noise_on[q] = noise_on[q];
    P.any_noise |= noise_on[q];
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
  if ((flags & B200AUG_F_HALF_PIXEL) && (category == B200AUG_CAT_POINTS || category == B200AUG_CAT_XYS)) {
    // offset_points_by_half_pixel: the affine [[1,0,.5],[0,1,.5]] has det = scales = 1, so only x and y move
    v[0] = add(v[0], 0.5f);
    v[1] = add(v[1], 0.5f);
  }
  if (flags & B200AUG_F_FOCUS) transform_item(P.t1, category, v, dim);
  if ((flags & B200AUG_F_FLIPROT) && P.has_t2) transform_item(P.t2, category, v, dim);
  if (flags & B200AUG_F_NORMALIZE) transform_item(P.t3, category, v, dim);
}

// All items of all transformable fields of one sample form one flat list that the threads [t0, t0 + nt) of the CTA share (one
// pass for the pose pipeline's 68 + 1 + 1 + 1 items): the fields are not walked one after the other.  Runs in plan_kernel.
__device__ __noinline__ void transform_labels(const B200AugFusedArgs& a, const Plan& P, int b, int t0, int nt) {
  const int tl = (int)threadIdx.x - t0;
  if (tl < 0 || tl >= nt) return;
  // an odd number of mirroring transforms permutes the left/right landmarks
  const bool flip_parity = (((a.flags & B200AUG_F_FOCUS) && P.t1.det < 0.f) + (P.has_t2 && P.t2.det < 0.f)) & 1;
  int total = 0;
  for (int f = 0; f < a.n_fields; ++f) {
    const B200AugField& F = a.fields[f];
    if (!F.out || !F.in) continue;
    if (F.category == B200AUG_CAT_BACKTRANSFORM) {
      // "image_backtransform": BT @ tr^-1 for every stage of this call, in pipeline order (affinetrafo.py:137-147)
      if (tl == 0) {
        const float* in = F.in + (size_t)b * 6;
        Aff bt = Aff{in[0], in[1], in[2], in[3], in[4], in[5]};
        if (a.flags & B200AUG_F_FOCUS) bt = aff_compose(bt, aff_inv(P.t1.m));
        if ((a.flags & B200AUG_F_FLIPROT) && P.has_t2) bt = aff_compose(bt, aff_inv(P.t2.m));
        if (a.flags & B200AUG_F_NORMALIZE) bt = aff_compose(bt, aff_inv(P.t3.m));
        float* out = F.out + (size_t)b * 6;
        out[0] = bt.a00; out[1] = bt.a01; out[2] = bt.a02; out[3] = bt.a10; out[4] = bt.a11; out[5] = bt.a12;
      }
      continue;
    }
    if (F.category == B200AUG_CAT_GENERAL || F.dim > 4) {
      const float* in = F.in + (size_t)b * F.count * F.dim;
      float* out = F.out + (size_t)b * F.count * F.dim;
      if (in != out)
        for (int i = tl; i < F.count * F.dim; i += nt) out[i] = in[i];
      continue;
    }
    total += F.count;
  }
  for (int t = tl; t < total; t += nt) {
    // locate item t: field f, index i
    int f = 0, i = t;
    for (; f < a.n_fields; ++f) {
      const B200AugField& F = a.fields[f];
      if (!F.out || !F.in || F.category == B200AUG_CAT_GENERAL || F.category == B200AUG_CAT_BACKTRANSFORM || F.dim > 4) continue;
      if (i < F.count) break;
      i -= F.count;
    }
    const B200AugField& F = a.fields[f];
    const bool is_roi_from_lm = (a.flags & B200AUG_F_ROI_FROM_LANDMARKS) && f == a.roi_field;
    if (is_roi_from_lm) continue;  // written by warp 2 of build_plan
    const int dim = F.dim, cnt = F.count, cat = F.category;
    const float* in = F.in + (size_t)b * cnt * dim;
    float* out = F.out + (size_t)b * cnt * dim;
    const int si = (cat == B200AUG_CAT_POINTS && cnt == 68 && flip_parity) ? flip_map68(i) : i;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < dim) v[k] = in[si * dim + k];
    run_item_chain(P, a.flags, cat, v, dim);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < dim) out[i * dim + k] = v[k];
  }
}

// PutRoiFromLandmarks (batch/misc.py:22-25): [min_xy, max_xy] over the landmarks, by one warp.
__device__ void landmark_box(const float* pts, int cnt, int dim, bool half_pixel, const AffDerived* tr, float box[4]) {
  const int lane = threadIdx.x & 31;
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = lane; i < cnt; i += 32) {
    float v[3] = {pts[i * dim], pts[i * dim + 1], 0.f};
    if (half_pixel) { v[0] = add(v[0], 0.5f); v[1] = add(v[1], 0.5f); }
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

// cv::interpolateCubic (A = -0.75) and cv::interpolateLanczos4, float32 / double exactly as OpenCV evaluates them
// (oracle/cv2_model.py:cubic_coeffs, lanczos4_coeffs)
__host__ __device__ inline void cubic_coeffs(float x, float* c) {
  const float A = -0.75f;
#ifdef __CUDA_ARCH__
  const float x1 = __fadd_rn(x, 1.f), xm = __fsub_rn(1.f, x);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), __fmul_rn(5.f, A)), x1), __fmul_rn(8.f, A)), x1), __fmul_rn(4.f, A));
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), x), __fadd_rn(A, 3.f)), x), x), 1.f);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), xm), __fadd_rn(A, 3.f)), xm), xm), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
#else
  volatile float t;  // (volatile: every operation rounds to float32 on its own, whatever the host compiler would contract)
  const float x1 = x + 1.f, xm = 1.f - x;
  t = A * x1; t = t - 5.f * A; t = t * x1; t = t + 8.f * A; t = t * x1; t = t - 4.f * A; c[0] = t;
  t = (A + 2.f) * x; t = t - (A + 3.f); t = t * x; t = t * x; t = t + 1.f; c[1] = t;
  t = (A + 2.f) * xm; t = t - (A + 3.f); t = t * xm; t = t * xm; t = t + 1.f; c[2] = t;
  t = 1.f - c[0]; t = t - c[1]; t = t - c[2]; c[3] = t;
#endif
}

__host__ __device__ inline void lanczos4_coeffs(float x, float* c) {
  const double s45 = 0.70710678118654752440084436210485;
  const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  if (x < 1.1920929e-07f) {  // FLT_EPSILON
    for (int i = 0; i < 8; ++i) c[i] = 0.f;
    c[3] = 1.f;
    return;
  }
  const double pi4 = 3.1415926535897932384626433832795 * 0.25;
  const double y0 = -((double)x + 3) * pi4, s0 = sin(y0), c0 = cos(y0);
#ifdef __CUDA_ARCH__
  float sum = 0.f;
  for (int i = 0; i < 8; ++i) {
    const double y = -((double)x + 3 - i) * pi4;
    c[i] = (float)__ddiv_rn(__dadd_rn(__dmul_rn(cs[i][0], s0), __dmul_rn(cs[i][1], c0)), __dmul_rn(y, y));
    sum = __fadd_rn(sum, c[i]);
  }
  sum = __fdiv_rn(1.f, sum);
  for (int i = 0; i < 8; ++i) c[i] = __fmul_rn(c[i], sum);
#else
  volatile float sum = 0.f;
  for (int i = 0; i < 8; ++i) {
    const double y = -((double)x + 3 - i) * pi4;
    volatile double num = cs[i][0] * s0;
    volatile double num2 = cs[i][1] * c0;
    c[i] = (float)((num + num2) / (y * y));
    sum = sum + c[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; ++i) { volatile float v = c[i] * sum; c[i] = v; }
#endif
}

// cv::resize INTER_CUBIC / INTER_LANCZOS4 taps of one output coordinate (oracle/cv2_model.py:resize_taps): first tap index
// (before the border clamp) and the k 11-bit taps, packed two shorts per word
__device__ void kernel_tab_entry(int d, double scale, int k, int& first, uint32_t (&packed)[4]) {
  float f = (float)((d + 0.5) * scale - 0.5);
  const int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  float c[8];
  if (k == 4) cubic_coeffs(f, c);
  else lanczos4_coeffs(f, c);
  int t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = (i < k) ? min(max(__float2int_rn(__fmul_rn(c[i], 2048.f)), -32768), 32767) : 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) packed[i] = ((uint32_t)t[2 * i] & 0xffffu) | ((uint32_t)t[2 * i + 1] << 16);
  first = s - (k / 2 - 1);
}

// cv::resize INTER_LINEAR taps (oracle/cv2_model.py:linear_tab + border rules)
__device__ void linear_tab_entry(int d, double scale, int ssize, int dsize, bool area_mode, bool is_x, int& i0, int& i1, int& w0, int& w1) {
  float f;
  int s;
  if (!area_mode) {
    f = (float)((d + 0.5) * scale - 0.5);
    s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
  } else {  // INTER_AREA asked for, an axis up-scales: sx = floor(dx * scale), fx = (dx + 1) - (sx + 1) * inv_scale, wrapped into [0, 1)
    const double inv_scale = (double)dsize / (double)ssize;
    s = (int)floor(d * scale);
    f = (float)((double)(d + 1) - (double)(s + 1) * inv_scale);
    f = (f <= 0.f) ? 0.f : __fsub_rn(f, floorf(f));
  }
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

// ------------------------------------------------------------------------------------------------ canvas

// One canvas pixel, scalar: the zero-padded crop (image_geometric_cv2.py:28-44) or cv2.warpAffine INTER_LINEAR /
// BORDER_CONSTANT(0) in fixed point (oracle/cv2_model.py:warp_affine_linear_u8).  This is the specification every
// fast path below must reproduce; the per-pixel fallback calls it directly.
__device__ __forceinline__ int bilinear_q5(int p00, int p01, int p10, int p11, int fx, int fy) {
  const int top = (32 - fx) * p00 + fx * p01;
  const int bot = (32 - fx) * p10 + fx * p11;
  return ((32 - fy) * top + fy * bot + 512) >> 10;
}

__device__ int canvas_px(const Plan& P, int x, int y) {
  if (P.src_mode != SRC_WARP) {
    const int sx = P.x0 + x, sy = P.y0 + y;
    // (coherent load: the source may be the scratch canvas written earlier in this kernel)
    return (sx >= 0 && sx < P.sw && sy >= 0 && sy < P.sh) ? (int)P.src[(size_t)sy * P.pitch + sx] : 0;
  }
  const int X0 = rint_d2i(__dmul_rn(__dadd_rn(__dmul_rn(P.mi[1], (double)y), P.mi[2]), 1024.0)) + 16;
  const int Y0 = rint_d2i(__dmul_rn(__dadd_rn(__dmul_rn(P.mi[4], (double)y), P.mi[5]), 1024.0)) + 16;
  const int ad = rint_d2i(__dmul_rn(__dmul_rn(P.mi[0], (double)x), 1024.0));
  const int bd = rint_d2i(__dmul_rn(__dmul_rn(P.mi[3], (double)x), 1024.0));
  const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
  const int ix = X >> 5, iy = Y >> 5, fx = X & 31, fy = Y & 31;
  if (P.warp_interp != 0) {
    // INTER_CUBIC / INTER_LANCZOS4 (oracle/cv2_model.py:warp_affine_cubic_or_lanczos_u8): k x k taps of cv2's fixed-point
    // table for the 1/32-pixel phase, taps outside the image count as 0 (BORDER_CONSTANT), (sum + 2^14) >> 15
    const int k = (P.warp_interp == B200AUG_UP_CUBIC) ? 4 : 8, o = k / 2 - 1;
    const int16_t* w = P.warp_tab + (size_t)(fy * 32 + fx) * (k * k);
    int sum = 0;
    for (int r = 0; r < k; ++r) {
      const int yy = iy - o + r;
      if ((unsigned)yy >= (unsigned)P.sh) continue;
      const uint8_t* row = P.src + (ptrdiff_t)yy * P.pitch;
      for (int q = 0; q < k; ++q) {
        const int xx = ix - o + q;
        if ((unsigned)xx < (unsigned)P.sw) sum += (int)__ldg(row + xx) * (int)__ldg(w + r * k + q);
      }
    }
    return min(max((sum + (1 << 14)) >> 15, 0), 255);
  }
  const bool r0 = (iy >= 0) && (iy < P.sh), r1 = (iy + 1 >= 0) && (iy + 1 < P.sh);
  const bool c0 = (ix >= 0) && (ix < P.sw), c1 = (ix + 1 >= 0) && (ix + 1 < P.sw);
  const uint8_t* p = P.src + (ptrdiff_t)iy * P.pitch + ix;
  const int p00 = (r0 && c0) ? __ldg(p) : 0;
  const int p01 = (r0 && c1) ? __ldg(p + 1) : 0;
  const int p10 = (r1 && c0) ? __ldg(p + P.pitch) : 0;
  const int p11 = (r1 && c1) ? __ldg(p + P.pitch + 1) : 0;
  return bilinear_q5(p00, p01, p10, p11, fx, fy);
}

__device__ __forceinline__ uint8_t sat_u8_rint(float v) {
  int r = __float2int_rn(v);
  return (uint8_t)min(max(r, 0), 255);
}

__device__ __forceinline__ float area_alpha(int t, int n, bool hf, bool hl, float af, float am, float al) {
  return (t == 0 && hf) ? af : ((t == n - 1 && hl) ? al : am);
}

// One output pixel of the resize, scalar, for every resampler (the fallback path and the reference semantics of the
// fast path): cv2.resize INTER_AREA (general + integer factor), INTER_LINEAR, or a plain copy.
__device__ __noinline__ uint8_t scalar_out_px(const Plan& P, const Tabs& T, int ow, int dx, int dy) {
  switch (P.rs_mode) {
    case RS_AREA: {
      const int xs = T.start[dx], xnf = T.n[dx], xn = xnf & 0xffff;
      const bool xhf = xnf & (1 << 30), xhl = xnf & (1u << 31);
      const float xaf = T.a[dx], xam = T.b[dx], xal = T.c[dx];
      const int ys = T.start[ow + dy], ynf = T.n[ow + dy], yn = ynf & 0xffff;
      const bool yhf = ynf & (1 << 30), yhl = ynf & (1u << 31);
      const float yaf = T.a[ow + dy], yam = T.b[ow + dy], yal = T.c[ow + dy];
      float acc = 0.f;
      for (int k = 0; k < yn; ++k) {
        float h = 0.f;
        for (int t = 0; t < xn; ++t)
          h = __fadd_rn(h, __fmul_rn((float)canvas_px(P, xs + t, ys + k), area_alpha(t, xn, xhf, xhl, xaf, xam, xal)));
        const float beta = area_alpha(k, yn, yhf, yhl, yaf, yam, yal);
        acc = (k == 0) ? __fmul_rn(beta, h) : __fadd_rn(acc, __fmul_rn(beta, h));
      }
      return sat_u8_rint(acc);
    }
    case RS_AREA_INT: {
      int acc = 0;
      for (int k = 0; k < P.iscale_y; ++k)
        for (int t = 0; t < P.iscale_x; ++t) acc += canvas_px(P, dx * P.iscale_x + t, dy * P.iscale_y + k);
      return (P.iscale_x == 2 && P.iscale_y == 2) ? (uint8_t)((acc + 2) >> 2) : sat_u8_rint(__fmul_rn((float)acc, P.inv_area));
    }
    case RS_LINEAR: {
      const int x0 = T.start[dx], x1 = T.n[dx];
      const int a0 = __float_as_int(T.a[dx]), a1 = __float_as_int(T.b[dx]);
      const int r0 = T.start[ow + dy], r1 = T.n[ow + dy];
      const int b0 = __float_as_int(T.a[ow + dy]), b1 = __float_as_int(T.b[ow + dy]);
      const int h0 = canvas_px(P, x0, r0) * a0 + canvas_px(P, x1, r0) * a1;
      const int h1 = canvas_px(P, x0, r1) * a0 + canvas_px(P, x1, r1) * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      return (uint8_t)min(max(v, 0), 255);
    }
    case RS_CUBIC:
    case RS_LANCZOS: {
      // cv2.resize INTER_CUBIC / INTER_LANCZOS4 (oracle/cv2_model.py:resize_cubic_or_lanczos_u8): rows in int32 with 11-bit
      // taps, source indices clamped to the canvas; columns: Lanczos in integers, cubic in float32 as OpenCV's vector loop
      const int k = (P.rs_mode == RS_CUBIC) ? 4 : 8;
      const int xs = T.start[dx], ys = T.start[ow + dy];
      const uint32_t xw[4] = {(uint32_t)T.n[dx], __float_as_uint(T.a[dx]), __float_as_uint(T.b[dx]), __float_as_uint(T.c[dx])};
      const uint32_t yw[4] = {(uint32_t)T.n[ow + dy], __float_as_uint(T.a[ow + dy]), __float_as_uint(T.b[ow + dy]), __float_as_uint(T.c[ow + dy])};
      auto tap = [](const uint32_t (&w)[4], int i) { return (int)(int16_t)(w[i >> 1] >> (16 * (i & 1))); };
      int hsum[8];
      for (int r = 0; r < k; ++r) {
        const int yy = min(max(ys + r, 0), P.ch - 1);
        int h = 0;
        for (int q = 0; q < k; ++q) h += canvas_px(P, min(max(xs + q, 0), P.cw - 1), yy) * tap(xw, q);
        hsum[r] = h;
      }
      if (k == 4) {
        const float sc = 1.f / (2048.f * 2048.f);
        float acc = __fmul_rn((float)hsum[3], __fmul_rn((float)tap(yw, 3), sc));
        acc = __fmaf_rn((float)hsum[2], __fmul_rn((float)tap(yw, 2), sc), acc);
        acc = __fmaf_rn((float)hsum[1], __fmul_rn((float)tap(yw, 1), sc), acc);
        acc = __fmaf_rn((float)hsum[0], __fmul_rn((float)tap(yw, 0), sc), acc);
        return sat_u8_rint(acc);
      }
      long long v = 0;
      for (int r = 0; r < 8; ++r) v += (long long)hsum[r] * tap(yw, r);
      return (uint8_t)min(max((v + (1ll << 21)) >> 22, 0ll), 255ll);
    }
    default: return (uint8_t)canvas_px(P, dx, dy);  // RS_COPY
  }
}

// ------------------------------------------------------------------------------------------------ row staging (shared memory)

// Zero-padded crop row `y` of the canvas, columns [lo, hi), into `buf` (used for rows that touch the image border;
// interior rows are read straight from global memory by the horizontal pass).  Returns the offset of column `lo`.
__device__ __forceinline__ int stage_crop_row(const Plan& P, int y, int lo, int hi, uint8_t* buf, int lane) {
  const int sy = P.y0 + y;
  const int sx_lo = P.x0 + lo;
  const bool row_in = (sy >= 0) && (sy < P.sh);
  for (int i = lane; i < hi - lo; i += 32) {
    const int sx = sx_lo + i;
    buf[i] = (row_in && sx >= 0 && sx < P.sw) ? P.src[(size_t)sy * P.pitch + sx] : (uint8_t)0;  // coherent load: the
  }                                                    // source may be the scratch canvas written earlier in this kernel
  return 0;
}

// ------------------------------------------------------------------------------------------------ INTER_AREA fast path

// Horizontal pass of one staged canvas row for the lane's columns: h[j] = sum_t w[j][t] * S[xoff[j] + t], taps in table
// order, float32 unfused -- padded taps have weight +0 and leave the sum unchanged.  `base32` is the shared-memory
// address of canvas column 0 of the row.
template <int K>
__device__ __forceinline__ void hrow(uint32_t base32, const int (&xoff)[RMAX], const float (&w)[RMAX][K], float (&h)[RMAX]) {
  uint32_t px[RMAX][K];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    const uint32_t a = base32 + (uint32_t)xoff[j];
#pragma unroll
    for (int t = 0; t < K; ++t) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(px[j][t]) : "r"(a + t) : "memory");
  }
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    float s = __fmul_rn(__uint2float_rn(px[j][0]), w[j][0]);
#pragma unroll
    for (int t = 1; t < K; ++t) s = __fadd_rn(s, __fmul_rn(__uint2float_rn(px[j][t]), w[j][t]));
    h[j] = s;
  }
}

// The 2-tap kernels (cv2.resize INTER_LINEAR, or INTER_AREA with an up-scaling axis): h[j] = p0 * a0 + p1 * a1 in integers
// (11-bit coefficients, oracle/cv2_model.py:resize_linear_u8); weights and results travel as bit patterns in the float
// arrays of the INTER_AREA path.  The second tap is the next byte: where cv2 clamps it onto the first (frame border) its
// weight is 0.
__device__ __forceinline__ void hrow_linear(uint32_t base32, const int (&xoff)[RMAX], const float (&w)[RMAX][2], float (&h)[RMAX]) {
  uint32_t px[RMAX][2];
#pragma unroll
  for (int j = 0; j < RMAX; ++j) {
    const uint32_t a = base32 + (uint32_t)xoff[j];
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(px[j][0]) : "r"(a) : "memory");
    asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(px[j][1]) : "r"(a) : "memory");
  }
#pragma unroll
  for (int j = 0; j < RMAX; ++j)
    h[j] = __int_as_float((int)px[j][0] * __float_as_int(w[j][0]) + (int)px[j][1] * __float_as_int(w[j][1]));
}

__device__ __forceinline__ uint32_t cvt_rni_sat_u8(float v) {
  uint32_t r;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

struct TileMap {  // tile index of output pixel (dy, dx) = o + dy * sa + dx * sb (flip / rot90 folded in)
  int o, sa, sb;
};

// geometric.py:256-264: flip(-1), then swapaxes + flip for the 90-degree rotations
__device__ __forceinline__ TileMap make_tile_map(const Plan& P, int ow, int oh) {
  TileMap m;
  if (P.rot_dir == 0) { m.o = 0; m.sa = ow; m.sb = 1; }
  else if (P.rot_dir == 1) { m.o = ow - 1; m.sa = -1; m.sb = ow; }
  else { m.o = (oh - 1) * ow; m.sa = 1; m.sb = -ow; }
  if (P.do_flip) { m.o += (ow - 1) * m.sb; m.sb = -m.sb; }
  return m;
}

// What every band of the CTA needs from cv2's resize tables, worked out ONCE per CTA by all threads (each warp used to
// redo its share with lanes = rows / columns):
//   prog[i], canvas row Rc0 + i of the CTA's output rows [rows_lo, rows_hi): (+-beta_a, beta_b) of the vertical pass --
//     acc += beta_a * h; a minus sign on beta_a: the output row is complete -- store it, move on, restart the sum as
//     beta_b * h (beta_b != 0 only when this canvas row is also the next output row's first tap);
//   wtab[t * wpad + dx]: weight of tap t of band column dx, dense and zero padded (also zero for taps outside the frame:
//     zero padding = weight +0, so only the in-frame part of a row is ever fetched);  xbt[dx]: first tap of the column
//     relative to its 128-column group's fetched segment.
// (The caller checked that the CTA's canvas rows fit the program.)
__device__ __noinline__ void area_precompute(int K, int ow, int oh, int ow_band, int rows_lo, int rows_hi, int cap, bool linear) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(ow, oh, cap);
  const Plan& P = *reinterpret_cast<const Plan*>(smem);
  const int* const tstart = reinterpret_cast<const int*>(smem + L.off_tabs);
  const int* const tn = tstart + L.ntab;
  const float* const ta = reinterpret_cast<const float*>(tn + L.ntab);
  const float* const tb = ta + L.ntab;
  const float* const tc = tb + L.ntab;
  float2* const prog = reinterpret_cast<float2*>(smem + L.off_prog);
  float* const wtab = reinterpret_cast<float*>(smem + L.off_wtab);
  int* const xbt = reinterpret_cast<int*>(wtab + (size_t)KMAX * L.wpad);
  const int tid = threadIdx.x;
  const int cfl = max(0, -P.x0), cfh = min(P.cw, P.sw - P.x0);
  if (linear) {
    // 2-tap kernels: taps (x0, x0 + 1) with cv2's integer coefficients (as bit patterns); no vertical program -- the band
    // walks cv2's row table directly
    for (int dx = tid; dx < L.wpad; dx += NTHREADS) {
      const bool valid = dx < ow_band;
      const int g0 = dx & ~(32 * RMAX - 1);
      const int dxc = valid ? dx : g0;
      const int xs = tstart[dxc];
      xbt[dx] = xs - max(tstart[g0], cfl);
      const bool in0 = valid && xs >= cfl && xs < cfh, in1 = valid && xs + 1 >= cfl && xs + 1 < cfh;
      wtab[dx] = in0 ? ta[dxc] : 0.f;              // (the bit pattern of integer 0 is +0.f)
      wtab[L.wpad + dx] = in1 ? tb[dxc] : 0.f;
    }
    return;
  }
  auto row_end = [&](int d) { return tstart[ow + d] + (tn[ow + d] & 0xffff) - 1; };
  const int Rc0 = tstart[ow + rows_lo], Rc1 = row_end(rows_hi - 1), n_rows = min(Rc1 - Rc0 + 1, NWARPS * ROWPROG_CAP);
  const float inv_sy = (float)(1.0 / P.scale_y);
  for (int i = tid; i < n_rows; i += NTHREADS) {
    const int r = Rc0 + i;
    // the lowest output row of this CTA that still uses canvas row r (estimate, then walk)
    int d = min(max(rows_lo + (int)((float)i * inv_sy), rows_lo), rows_hi - 1);
    while (d > rows_lo && row_end(d - 1) >= r) --d;
    while (d < rows_hi - 1 && row_end(d) < r) ++d;
    const int nf = tn[ow + d], yn = nf & 0xffff, k = r - tstart[ow + d];
    float ba = 0.f, bb = 0.f;
    bool emit = false;
    if (k >= 0 && k < yn) {
      ba = area_alpha(k, yn, nf & (1 << 30), nf & (1u << 31), ta[ow + d], tb[ow + d], tc[ow + d]);
      emit = (k == yn - 1);
      if (emit && d + 1 < rows_hi && tstart[ow + d + 1] == r) {
        const int nf2 = tn[ow + d + 1];
        bb = area_alpha(0, nf2 & 0xffff, nf2 & (1 << 30), nf2 & (1u << 31), ta[ow + d + 1], tb[ow + d + 1], tc[ow + d + 1]);
      }
    }
    prog[i] = make_float2(emit ? -ba : ba, bb);
  }
  for (int dx = tid; dx < L.wpad; dx += NTHREADS) {
    const bool valid = dx < ow_band;
    const int g0 = dx & ~(32 * RMAX - 1);
    const int dxc = valid ? dx : g0;
    const int xnf = tn[dxc], xn = xnf & 0xffff, xs = tstart[dxc];
    const bool xhf = xnf & (1 << 30), xhl = xnf & (1u << 31);
    const float xaf = ta[dxc], xam = tb[dxc], xal = tc[dxc];
    xbt[dx] = xs - max(tstart[g0], cfl);
    for (int t = 0; t < K; ++t) {
      const int col = xs + t;
      wtab[t * L.wpad + dx] = (valid && t < xn && col >= cfl && col < cfh) ? area_alpha(t, xn, xhf, xhl, xaf, xam, xal) : 0.f;
    }
  }
}

// cv2.resize INTER_AREA, general (non-integer) factor: each warp owns a band of output rows and streams the canvas
// rows that feed it exactly once, in order.  Per canvas row: horizontal pass for the lane's columns (registers hold
// the column taps), then the vertical accumulation in source-row order; a row shared by two output rows (fractional
// boundary) is used for both.
// Crop rows inside the frame stream through a per-warp ring of D slots (cp.async, 16 bytes per lane, a fixed number of
// 16-byte vectors per row: the aligned superset of the row segment).  The band is walked in three phases so that the
// steady state carries no row classification: rows above the frame (zeros) / the fetched rows / rows below the frame or
// the frame's last row when its over-read would leave the image (staged synchronously, zero padded).
// The canvas is always a crop here: rotated samples were turned into one by the canvas workers.
// DIRECT: the sample's photometric chain is one point function (no equalize / blur / noise) and there is no 90-degree
// rotation, so the finished uint8 pixel goes through the (already built) LUT straight to the float32 output `gimg` --
// no tile, no cluster exchange, no separate output pass.
// LINEAR (K = 2): the same streaming band for cv2's 2-tap kernels -- integer horizontal pass per canvas row, and instead of
// the vertical program the band emits, behind every canvas row r, the output rows whose second tap row is r (their first
// tap row is r - 1 or, clamped at the border, r itself): (((b0 * (hA >> 4)) >> 16) + ((b1 * (hB >> 4)) >> 16) + 2) >> 2.
template <int K, bool LINEAR>
__device__ __noinline__ void area_band(const TileMap tm, int cap, int ow, int oh, int ow_band, int warp, int lane, int cr, int cl,
                                       float* __restrict__ gimg) {
  static_assert(!LINEAR || K == 2, "the 2-tap kernels have two taps");
  const bool DIRECT = gimg != nullptr;  // (a run-time flag: one instantiation per K keeps the hot code small)
  const int cs = 31 - __clz(cl);  // log2 of the cluster size
  const int rows_lo = (cr * oh) >> cs, rows_n = (((cr + 1) * oh) >> cs) - rows_lo;  // this CTA's band of output rows
  const int dy_begin = rows_lo + (warp * rows_n) / NWARPS, dy_end = rows_lo + ((warp + 1) * rows_n) / NWARPS;
  if (dy_begin >= dy_end) return;
  // everything lives in this CTA's dynamic shared memory; deriving the pointers here keeps the address space known
  // (a generic pointer passed into a non-inlined function turns every table read into a generic load)
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(ow, oh, cap);
  const Plan& P = *reinterpret_cast<const Plan*>(smem);
  Tabs T;
  T.start = reinterpret_cast<int*>(smem + L.off_tabs);
  T.n = T.start + L.ntab;
  T.a = reinterpret_cast<float*>(T.n + L.ntab);
  T.b = T.a + L.ntab;
  T.c = T.b + L.ntab;
  const uint32_t tile32 = smem_u32(smem + L.off_tile);
  uint8_t* const rowbuf = smem + L.off_rowbuf + (size_t)warp * (cap + ROWBUF_SLACK);

  // plan fields used per row live in registers (the tile stores would otherwise force reloads from shared memory)
  const uint8_t* const src = P.src;
  const int pitch = P.pitch, x0 = P.x0, y0 = P.y0, sw = P.sw, sh = P.sh, cw = P.cw;
  const int fin = P.fin;
  const float inv_area = P.inv_area;
  // canvas columns [cfl, cfh) lie inside the frame; taps outside read zeros (zero padding), which is the same as giving
  // them weight +0 -- so only the in-frame part of a row is ever fetched
  const int cfl = max(0, -x0), cfh = min(cw, sw - x0);
  const uint32_t rowbuf32 = smem_u32(rowbuf);

  // ---- vertical-pass program (area_precompute): this band's canvas rows R0 .. R1 inside the CTA's program.  When the
  // band's first canvas row is also the last tap of the output row above (which belongs to another warp), its entry
  // (-beta_last, beta_first) must count as (beta_first, 0) here: `first_shared` patches the first entry on the fly.
  const int last = dy_end - 1;
  const int R0 = T.start[ow + dy_begin];
  const int R1 = LINEAR ? T.n[ow + last] : T.start[ow + last] + (T.n[ow + last] & 0xffff) - 1;
  const float2* const prog = reinterpret_cast<const float2*>(smem + L.off_prog) + (LINEAR ? 0 : R0 - T.start[ow + rows_lo]);
  const bool band_first_shared = !LINEAR && dy_begin > rows_lo && T.start[ow + dy_begin - 1] + (T.n[ow + dy_begin - 1] & 0xffff) - 1 >= R0;
  const bool first_single = (T.n[ow + dy_begin] & 0xffff) == 1;  // (the shared row is the band's first output row's only tap)
  const float* const wtab = reinterpret_cast<const float*>(smem + L.off_wtab);
  const int* const xbt = reinterpret_cast<const int*>(wtab + (size_t)KMAX * L.wpad);
  const int wpad = L.wpad;

  for (int g0 = 0; g0 < ow_band; g0 += 32 * RMAX) {
    const int gcols = min(32 * RMAX, ow_band - g0);
    const int glast = g0 + gcols - 1;
    bool first_shared = band_first_shared;  // (every column group walks the band's rows from the top)
    const int seg_lo = max(T.start[g0], cfl);
    const int seg_hi = max(min(T.start[glast] + (LINEAR ? 2 : (T.n[glast] & 0xffff)), cfh), seg_lo);
    const int seg_bytes = seg_hi - seg_lo;
    // the lane's columns are g0 + lane + 32 j, j < nvalid; their pixels in the tile row being accumulated sit at
    // trow32 + j * cstep (shared-memory address, flip / rot90 folded in)
    const int nvalid = (gcols - lane + 31) >> 5;
    uint32_t trow32 = tile32 + (uint32_t)(tm.o + (g0 + lane) * tm.sb + dy_begin * tm.sa);
    int cstep = 32 * tm.sb;
    // DIRECT: the same pixel as a float32 in global memory, and the LUT behind the plan in shared memory
    float* grow = DIRECT ? gimg + (tm.o + (g0 + lane) * tm.sb + dy_begin * tm.sa) : nullptr;
    const uint32_t lut32 = smem_u32(smem + ((sizeof(Plan) + 15) & ~size_t(15)));
    int xb[RMAX];        // first tap of the column, relative to the segment
    // ring geometry: every fetched row is copied as `nvec` 16-byte vectors starting at the 16-byte boundary at or below
    // its first byte, so the copy ends less than 16 * nvec bytes after the segment's first byte
    // (the caller checked that RING_D slots fit the warp's row buffer and that nvec <= 64)
    const int slot_bytes = ring_slot_bytes(seg_bytes);
    const int nvec = (seg_bytes + 30) >> 4;
    constexpr int D = RING_D;
    const int ring_bytes = D * slot_bytes;

    // rows [f_lo, f_hi] are fetched asynchronously
    int f_lo = INT_MAX, f_hi = INT_MIN;
    if (seg_bytes > 0) {
      const bool last_ok = x0 + seg_lo + 16 * nvec <= sw;  // the copy of the frame's last row stays inside that row
      f_lo = max(R0, -y0);
      f_hi = min(R1, (last_ok ? sh - 1 : sh - 2) - y0);
      if (f_lo > f_hi) { f_lo = INT_MAX; f_hi = INT_MIN; }
    }
    float acc[RMAX];  // INTER_AREA: the running sums | LINEAR: the previous canvas row's horizontal pass (bit patterns)
#pragma unroll
    for (int j = 0; j < RMAX; ++j) acc[j] = 0.f;
    int dy_emit = dy_begin;  // LINEAR: the next output row to emit

    // ring state of row r: byte offset of its slot, the global address of the row's segment
    int s_off = 0;
    uintptr_t ga = reinterpret_cast<uintptr_t>(src) + (ptrdiff_t)(y0 + R0) * pitch + (x0 + seg_lo);
    uintptr_t gf = ga + (ptrdiff_t)D * pitch;  // segment of row r + D, the one fetched while row r is consumed
    uint32_t dst_lane = rowbuf32 + 16u * (uint32_t)lane, rb32 = rowbuf32, lane16 = 16u * (uint32_t)lane;
    int row_sa = tm.sa, slot_b = slot_bytes, ring_b = ring_bytes;
    int my_vecs = (lane < nvec ? 1 : 0) + (lane + 32 < nvec ? 1 : 0);  // 16-byte vectors of a row this lane copies
    // opaque to the optimiser: otherwise it rematerialises these from the kernel parameters inside the row loop
    asm volatile("" : "+r"(dst_lane), "+r"(rb32), "+r"(trow32), "+r"(cstep), "+r"(row_sa));
    asm volatile("" : "+r"(lane16), "+r"(slot_b), "+r"(ring_b), "+r"(my_vecs));
    // one commit group per row, in row order (an empty group for rows that are not fetched): when row r is consumed,
    // the groups of rows <= r + D - 1 have been committed, so "all but the D - 1 newest complete" means row r has landed
    auto issue = [&](int soff, uintptr_t g, bool fetch) {  // whole warp: 16 bytes per lane
      if (fetch) {
        const uintptr_t g16 = (g & ~uintptr_t(15)) + lane16;
        const uint32_t d16 = dst_lane + (uint32_t)soff;
        // (predicated copies: an `if` around the asm compiles to a divergent branch + reconvergence per copy and row)
        asm volatile(
            "{\n.reg .pred p0, p1;\nsetp.gt.s32 p0, %2, 0;\nsetp.gt.s32 p1, %2, 1;\n"
            "@p0 cp.async.cg.shared.global [%0], [%1], 16;\n@p1 cp.async.cg.shared.global [%0+512], [%1+512], 16;\n}" ::"r"(d16),
            "l"(g16), "r"(my_vecs)
            : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // vertical pass (see the program above); a fresh sum starts from +0, and 0 + x is exact
    const float2* pp = prog;
    auto vertical = [&](const float (&h)[RMAX], int r) {
      if (LINEAR) {
        while (dy_emit < dy_end && T.n[ow + dy_emit] == r) {
          const bool same = T.start[ow + dy_emit] == r;  // both taps on this row (clamped at the frame border)
          const int b0 = __float_as_int(T.a[ow + dy_emit]), b1 = __float_as_int(T.b[ow + dy_emit]);
#pragma unroll
          for (int j = 0; j < RMAX; ++j) {
            const int hb = __float_as_int(h[j]), ha = same ? hb : __float_as_int(acc[j]);
            const int v = (((b0 * (ha >> 4)) >> 16) + ((b1 * (hb >> 4)) >> 16) + 2) >> 2;
            const uint32_t q = (uint32_t)min(max(v, 0), 255);
            if (DIRECT) {
              float f;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(lut32 + 4u * q) : "memory");
              if (j < nvalid) grow[j * cstep] = f;
            } else {
              if (j < nvalid) asm volatile("st.shared.u8 [%0], %1;" ::"r"(trow32 + (uint32_t)(j * cstep)), "r"(q) : "memory");
            }
          }
          if (DIRECT) grow += row_sa;
          else trow32 += (uint32_t)row_sa;
          ++dy_emit;
        }
#pragma unroll
        for (int j = 0; j < RMAX; ++j) acc[j] = h[j];
        return;
      }
      float2 pr = *pp++;
      if (first_shared) {  // (only ever true for the band's first canvas row)
        pr = make_float2(first_single ? -pr.y : pr.y, 0.f);
        first_shared = false;
      }
      const float ba = fabsf(pr.x);
#pragma unroll
      for (int j = 0; j < RMAX; ++j) acc[j] = __fadd_rn(acc[j], __fmul_rn(ba, h[j]));
      if (pr.x < 0.f) {
        if (DIRECT) {
#pragma unroll
          for (int j = 0; j < RMAX; ++j) {
            const uint32_t q = cvt_rni_sat_u8(acc[j]);
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(lut32 + 4u * q) : "memory");
            if (j < nvalid) grow[j * cstep] = v;
            acc[j] = __fmul_rn(pr.y, h[j]);
          }
          grow += row_sa;
        } else if (fin == 0) {
#pragma unroll
          for (int j = 0; j < RMAX; ++j) {
            const uint32_t q = cvt_rni_sat_u8(acc[j]);
            if (j < nvalid) asm volatile("st.shared.u8 [%0], %1;" ::"r"(trow32 + (uint32_t)(j * cstep)), "r"(q) : "memory");
            acc[j] = __fmul_rn(pr.y, h[j]);
          }
        } else {  // integer factors: (sum + 2) >> 2 for exact 2 x 2, rint(sum / area) otherwise
#pragma unroll
          for (int j = 0; j < RMAX; ++j) {
            const uint32_t q = (fin == 1) ? (uint32_t)(((int)acc[j] + 2) >> 2) : cvt_rni_sat_u8(__fmul_rn(acc[j], inv_area));
            if (j < nvalid) asm volatile("st.shared.u8 [%0], %1;" ::"r"(trow32 + (uint32_t)(j * cstep)), "r"(q) : "memory");
          }
#pragma unroll
          for (int j = 0; j < RMAX; ++j) acc[j] = __fmul_rn(pr.y, h[j]);
        }
        trow32 += (uint32_t)row_sa;
      }
    };
    auto advance = [&]() {
      ga += pitch;
      gf += pitch;
      s_off += slot_b;
      if (s_off == ring_b) s_off = 0;
    };
    __syncwarp();
    for (int i = 0; i < D; ++i) issue(i * slot_bytes, ga + (ptrdiff_t)i * pitch, R0 + i >= f_lo && R0 + i <= f_hi);
    // (the first rows are on their way: the tap weights are worked out while they fly)
    float w[RMAX][K];
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      const int dx = g0 + 32 * j + lane;  // (< wpad; columns past the band carry zero weights)
      xb[j] = xbt[dx];
#pragma unroll
      for (int t = 0; t < K; ++t) w[j][t] = wtab[t * wpad + dx];
    }

    int r = R0;
#pragma unroll 1
    for (int phase = 0; phase < 3; ++phase) {
      if (phase == 1) {
        // steady state: row r has been fetched into its slot
#pragma unroll 1
        for (; r <= f_hi; ++r) {
          static_assert(RING_D == 4, "wait_group immediate");
          asm volatile("cp.async.wait_group 3;" ::: "memory");
          __syncwarp();  // every lane's 16 bytes of the row are in
          float h[RMAX];
          if (LINEAR) hrow_linear(rb32 + (uint32_t)s_off + ((uint32_t)ga & 15u), xb, reinterpret_cast<const float(&)[RMAX][2]>(w), h);
          else hrow<K>(rb32 + (uint32_t)s_off + ((uint32_t)ga & 15u), xb, w, h);
          vertical(h, r);
          __syncwarp();  // the slot of row r is free again: refill it with row r + D
          issue(s_off, gf, r + D <= f_hi);
          advance();
        }
      } else {
        const int r_end = (phase == 0) ? min(R1, f_lo - 1) : R1;
#pragma unroll 1
        for (; r <= r_end; ++r) {
          float h[RMAX];
          const int sy = y0 + r;
          if (sy < 0 || sy >= sh || seg_bytes <= 0) {  // above / below the frame
#pragma unroll
            for (int j = 0; j < RMAX; ++j) h[j] = 0.f;
          } else {  // the frame's last row: zero-padded, staged synchronously in the row's own (idle) slot
            __syncwarp();
            stage_crop_row(P, r, seg_lo, seg_hi, rowbuf + s_off, lane);
            __syncwarp();
            if (LINEAR) hrow_linear(rb32 + (uint32_t)s_off, xb, reinterpret_cast<const float(&)[RMAX][2]>(w), h);
            else hrow<K>(rb32 + (uint32_t)s_off, xb, w, h);
          }
          vertical(h, r);
          __syncwarp();
          issue(s_off, gf, r + D >= f_lo && r + D <= f_hi);
          advance();
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ photometric helpers

constexpr float R255 = 0x1.010102p-8f;  // float32(1 / 255)

// pointwise stage-1 ops [from, to) of this sample's op list applied to x (oracle/photometric.py)
__device__ __noinline__ float apply_point_ops(const Plan& P, float x, int from, int to, const float* eq_lut) {
  for (int k = from; k < to; ++k) {
    switch (P.ops[k]) {
      case B200AUG_OP_EQUALIZE: {
        // kornia's `result / 255.0` runs on the GPU in the reference's pipeline, where torch divides a tensor by a host
        // scalar as a multiplication with the rounded reciprocal (oracle/photometric.py: R255)
        float im = __fmul_rn(x, 255.f);
        if (P.eq_step0) x = __fmul_rn(im, R255);
        else x = __fmul_rn(eq_lut[min(max((int)im, 0), 255)], R255);
      } break;
      case B200AUG_OP_POSTERIZE: {
        // kornia.enhance.posterize, literally: right shift = uint8(x * 255) / 2^s / 255, left shift = uint8(that * 255) * 2^s / 255
        const int q = (int)__fmul_rn(x, 255.f) & 255;
        const int sh = 8 - P.bits;
        const float right = __fmul_rn(__fmul_rn((float)q, __int_as_float((127 - sh) << 23)), R255);
        const int l = ((int)__fmul_rn(right, 255.f) & 255) << sh;
        x = __fmul_rn((float)(l & 255), R255);
      } break;
      case B200AUG_OP_GAMMA: {
        const float pw = powf(x, P.gamma);
        x = (pw != pw) ? pw : fminf(fmaxf(pw, 0.f), 1.f);  // torch.clamp propagates the NaN of pow(negative, g); fmaxf would not
      } break;
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

// value of pixel p after the LUT prefix and the 5x5 blur: input to the post-blur ops
__device__ __noinline__ float blurred_value(const uint8_t* tile, const float* lut, int ow, int oh, int p) {
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

// The plan of one sample, one branch per warp (see plan_core); also writes the per-sample side outputs (view box, focus
// transform, back-transform, the roi regenerated from landmarks).  Runs either in plan_kernel (ahead of the fused kernel, so
// that its latency chains do not occupy a big CTA slot) or at the top of the fused kernel itself.
__device__ __forceinline__ void build_plan(const B200AugFusedArgs& a, int b, Plan& P, int warp, int lane, bool side_out, PlanCore* shared_core) {
    float box[4] = {0.f, 0.f, 0.f, 0.f};
    const bool lm = (a.flags & B200AUG_F_ROI_FROM_LANDMARKS) && a.landmark_field >= 0;
    const bool half = a.flags & B200AUG_F_HALF_PIXEL;
    if (lm) {
      const B200AugField& F = a.fields[a.landmark_field];
      landmark_box(F.in + (size_t)b * F.count * F.dim, F.count, F.dim, half, nullptr, box);
    } else if (a.roi_field >= 0) {
      const float* r = a.fields[a.roi_field].in + 4 * (size_t)b;
      box[0] = r[0]; box[1] = r[1]; box[2] = r[2]; box[3] = r[3];
    }
    // the common chain (view box -> focus transform) once, by warp 0, handed to the branch warps through shared memory:
    // evaluating it redundantly in every warp saved a barrier but cost 7/8 of this function's issue slots
    if (warp == 0) {
      const PlanCore c0 = plan_core(a, b, box);
      if (lane == 0) *shared_core = c0;
    }
    __syncthreads();
    const PlanCore c = *shared_core;
    const bool focus = a.flags & B200AUG_F_FOCUS;
    switch (warp) {
      case 0:
        if (lane == 0) plan_basic_and_warp(a, c, P);
        break;
      case 1:
        if (lane == 0) plan_resize(a, c, P);
        break;
      case 2: {
        const AffDerived d1 = aff_derive(c.t1);
        if (lane == 0) {
          P.t1 = d1;
          P.cv_ptr = nullptr;
          P.cv_pitch = 0;
          P.cv_pad = 0;
          if (side_out && a.view_roi_out && focus) {
            int32_t* o = a.view_roi_out + 4 * (size_t)b;
            o[0] = c.vx0; o[1] = c.vy0; o[2] = c.vx1; o[3] = c.vy1;
          }
          if (side_out && a.tr_out && focus) {
            float* o = a.tr_out + 6 * (size_t)b;
            o[0] = c.t1.a00; o[1] = c.t1.a01; o[2] = c.t1.a02; o[3] = c.t1.a10; o[4] = c.t1.a11; o[5] = c.t1.a12;
          }
        }
        if (side_out && lm && a.roi_field >= 0 && a.fields[a.roi_field].out) {
          // landmarks mode: the roi label is regenerated from the transformed landmarks after the crop (pipelines.py:347-351)
          const B200AugField& F = a.fields[a.landmark_field];
          float nb[4];
          landmark_box(F.in + (size_t)b * F.count * F.dim, F.count, F.dim, half, focus ? &d1 : nullptr, nb);
          if (lane == 0) {
            if ((a.flags & B200AUG_F_FLIPROT) && (c.do_flip || c.rot_dir != 0)) tf_roi(aff_derive(fliprot_transform(c)), nb);
            if (a.flags & B200AUG_F_NORMALIZE) tf_roi(aff_derive(normalize_transform(c)), nb);
            float* o = a.fields[a.roi_field].out + 4 * (size_t)b;
            o[0] = nb[0]; o[1] = nb[1]; o[2] = nb[2]; o[3] = nb[3];
          }
        }
      } break;
      case 3:
        if (lane == 0) {
          P.has_t2 = (c.do_flip || c.rot_dir != 0);
          if (P.has_t2) P.t2 = aff_derive(fliprot_transform(c));
        }
        break;
      case 4:
        if (lane == 0) P.t3 = aff_derive(normalize_transform(c));
        break;
      case 5:
        if (lane == 0) plan_photo(a.photo, (a.flags & B200AUG_F_PHOTOMETRIC) != 0, b, P);
        break;
      default: break;
    }
  }

// cv2's per-axis resize tables of one sample (x entries first, then y), and Plan::kx
__device__ __forceinline__ void build_tables(const B200AugFusedArgs& a, Plan& P, const Tabs& T, int tid, int lane, int nthr) {
  const int ow = a.out_w, oh = a.out_h;
  const int rs = P.rs_mode;
  if (rs == RS_CUBIC || rs == RS_LANCZOS) {
    for (int i = tid; i < ow + oh; i += nthr) {
      const bool is_x = i < ow;
      int first;
      uint32_t packed[4];
      kernel_tab_entry(is_x ? i : i - ow, is_x ? P.scale_x : P.scale_y, rs == RS_CUBIC ? 4 : 8, first, packed);
      T.start[i] = first; T.n[i] = (int)packed[0]; T.a[i] = __uint_as_float(packed[1]); T.b[i] = __uint_as_float(packed[2]);
      T.c[i] = __uint_as_float(packed[3]);
    }
  } else if (rs == RS_AREA || rs == RS_LINEAR || rs == RS_AREA_INT) {
    // two independent entries per thread and iteration: 129 + 129 entries on 256 threads would otherwise pay a second,
    // nearly empty, round of double-precision latency
    auto tab_entry = [&](int i) {
      const bool is_x = i < ow;
      const int d = is_x ? i : i - ow;
      const double sc = is_x ? P.scale_x : P.scale_y;
      const int ss = is_x ? P.cw : P.ch;
      if (rs == RS_AREA_INT) {
        // integer factor: plain box sums (unit weights, exact in float32); the rounding happens at the end (Plan::fin)
        const int isc = is_x ? P.iscale_x : P.iscale_y;
        T.start[i] = d * isc; T.n[i] = isc; T.a[i] = 0.f; T.b[i] = 1.f; T.c[i] = 0.f;
        if (is_x) atomicMax(&P.kx, isc);
      } else if (rs == RS_AREA) {
        int st, nf; float af, am, al;
        area_tab_entry(d, sc, ss, st, nf, af, am, al);
        T.start[i] = st; T.n[i] = nf; T.a[i] = af; T.b[i] = am; T.c[i] = al;
        if (is_x) atomicMax(&P.kx, nf & 0xffff);
      } else {
        int i0, i1, w0, w1;
        linear_tab_entry(d, sc, ss, is_x ? ow : oh, P.lin_area != 0, is_x, i0, i1, w0, w1);
        T.start[i] = i0; T.n[i] = i1; T.a[i] = __int_as_float(w0); T.b[i] = __int_as_float(w1);
      }
    };
    const int ntab = ow + oh;
    if (rs == RS_AREA) {
      for (int i = tid; i < ntab; i += 2 * nthr) {
        const int i2 = i + nthr;
        const bool two = i2 < ntab;
        const int j = two ? i2 : i;  // (a lone entry is simply computed twice)
        const bool x1 = i < ow, x2 = j < ow;
        int st1, nf1, st2, nf2; float af1, am1, al1, af2, am2, al2;
        area_tab_entry(x1 ? i : i - ow, x1 ? P.scale_x : P.scale_y, x1 ? P.cw : P.ch, st1, nf1, af1, am1, al1);
        area_tab_entry(x2 ? j : j - ow, x2 ? P.scale_x : P.scale_y, x2 ? P.cw : P.ch, st2, nf2, af2, am2, al2);
        T.start[i] = st1; T.n[i] = nf1; T.a[i] = af1; T.b[i] = am1; T.c[i] = al1;
        T.start[j] = st2; T.n[j] = nf2; T.a[j] = af2; T.b[j] = am2; T.c[j] = al2;
        int kmax = x1 ? (nf1 & 0xffff) : 0;
        if (x2) kmax = max(kmax, nf2 & 0xffff);
        kmax = __reduce_max_sync(__activemask(), kmax);
        if (lane == 0 && kmax) atomicMax(&P.kx, kmax);
      }
    } else {
      for (int i = tid; i < ntab; i += nthr) tab_entry(i);
    }
  }
}

// bytes of one sample's record in B200AugFusedArgs::plans: the Plan, then the tables of out_w + out_h entries
__host__ __device__ inline size_t plan_bytes() { return (sizeof(Plan) + 15) & ~size_t(15); }
__host__ __device__ inline size_t plan_tab_bytes(int ow, int oh) { return ((size_t)(ow + oh) * 5 * 4 + 15) & ~size_t(15); }

// canvas geometry of a rotated sample in its workspace region: 16-byte aligned rows with room for the row copies' over-read
__host__ __device__ __forceinline__ int canvas_pitch(int cw) { return (cw + 16 + 15) & ~15; }

// The tail of B200AugFusedArgs::plans behind the B records: work-stealing counters of the canvas workers (one per slice
// of WK_SLICE samples), one completion counter per sample, one "rotated, canvas in the workspace" flag byte per sample.
constexpr int WK_SLICE = 1024;
constexpr int WK_MAX_SLICES = 64;
__host__ __device__ inline size_t plan_tail_bytes(int batch) {
  return (size_t)WK_MAX_SLICES * 4 + (size_t)batch * 4 + (((size_t)batch + 15) & ~size_t(15));
}
struct PlanTail {
  uint32_t* counters;  // [WK_MAX_SLICES] next work item of the slice
  uint32_t* done;      // [B] chunks of the sample's canvas finished (WK_CHUNKS = complete)
  uint8_t* flags;      // [B] 1 = rotated sample with a canvas in the workspace
};
__host__ __device__ inline PlanTail plan_tail(unsigned char* plans, int batch, int64_t plan_stride) {
  unsigned char* t = plans + (size_t)batch * plan_stride;
  PlanTail r;
  r.counters = reinterpret_cast<uint32_t*>(t);
  r.done = r.counters + WK_MAX_SLICES;
  r.flags = reinterpret_cast<uint8_t*>(r.done + batch);
  return r;
}

// Plans + resize tables + LABELS of all samples, one CTA per sample, ahead of the other kernels.  The labels only need the
// plan's transforms, so they are finished here (the big kernels never touch them); the record (plan + tables) goes to
// B200AugFusedArgs::plans when that buffer is given.
__global__ void __launch_bounds__(NTHREADS) plan_kernel(const __grid_constant__ KArgs K, int with_tables) {
  const B200AugFusedArgs& a = K.a;
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntab = with_tables ? a.out_w + a.out_h : 0;
  const size_t tab_bytes = with_tables ? plan_tab_bytes(a.out_w, a.out_h) : 0;
  Plan& P = *reinterpret_cast<Plan*>(smem);
  Tabs T;
  T.start = reinterpret_cast<int*>(smem + plan_bytes());
  T.n = T.start + ntab;
  T.a = reinterpret_cast<float*>(T.n + ntab);
  T.b = T.a + ntab;
  T.c = T.b + ntab;
  // programmatic dependent launch: the next kernel may be scheduled now; it waits (griddepcontrol.wait) for this whole grid
  // to complete and flush before it reads the plans
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  PlanTail tail = {nullptr, nullptr, nullptr};
  if (a.plans) tail = plan_tail(a.plans, a.batch, a.plan_stride);
  if (a.plans && b == 0 && tid < WK_MAX_SLICES) tail.counters[tid] = 0u;  // the canvas workers' work counters
  build_plan(a, b, P, warp, lane, true, reinterpret_cast<PlanCore*>(smem + plan_bytes() + tab_bytes));
  __syncthreads();
  if (tid == 0) {
    // rotated samples: the canvas goes to this sample's workspace region (if it fits), the canvas workers fill it
    bool canvas = false;
    if (P.prefilter && with_tables) {
      // prefiltered samples: prefilter_kernel produces the (cropped or warped) canvas itself and leaves it smoothed in the
      // workspace; without room for it the sample cannot be served
      const int spitch = canvas_pitch(P.cw);
      if (P.status == B200AUG_S_OK) {
        if (with_tables && a.plans && a.workspace && (int64_t)spitch * (P.ch + 1) <= a.workspace_stride) {
          P.cv_ptr = a.workspace + (size_t)b * a.workspace_stride;
          P.cv_pitch = spitch;
        } else {
          P.status = B200AUG_S_UNSUPPORTED;
        }
      }
    } else if (with_tables && a.plans && a.workspace && a.warp_ctas > 0 && P.src_mode == SRC_WARP && P.warp_interp == 0 &&
               P.status == B200AUG_S_OK && P.cw <= DT_CAP && P.ch <= DT_CAP) {  // (the workers' gather is the bilinear one)
      const int spitch = canvas_pitch(P.cw);
      if ((int64_t)spitch * (P.ch + 1) <= a.workspace_stride) {
        P.cv_ptr = a.workspace + (size_t)b * a.workspace_stride;
        P.cv_pitch = spitch;
        canvas = true;
      }
    }
    if (a.plans) {
      tail.flags[b] = canvas ? 1 : 0;
      tail.done[b] = 0u;
    }
    if (a.status_out) a.status_out[b] = P.status;
  }
  if (tid == 32 && a.backtransform_out && (a.flags & B200AUG_F_FOCUS) && (a.flags & B200AUG_F_INSERT_BACKTRANSFORM)) {
    // GeneralFocusRoi(insert_backtransform=True): tr^-1 of the focus stage (geometric.py:226-227), then BT @ tr^-1 for the
    // stages behind it in this call (affinetrafo.py:137-147)
    Aff bt = aff_inv(P.t1.m);
    if ((a.flags & B200AUG_F_FLIPROT) && P.has_t2) bt = aff_compose(bt, aff_inv(P.t2.m));
    if (a.flags & B200AUG_F_NORMALIZE) bt = aff_compose(bt, aff_inv(P.t3.m));
    float* o = a.backtransform_out + 6 * (size_t)b;
    o[0] = bt.a00; o[1] = bt.a01; o[2] = bt.a02; o[3] = bt.a10; o[4] = bt.a11; o[5] = bt.a12;
  }
  // the labels (warps 5-7) next to the resize tables (warps 0-4): both only read the finished plan
  transform_labels(a, P, b, 5 * 32, 3 * 32);
  if (with_tables) {
    if (tid < 5 * 32) build_tables(a, P, T, tid, lane, 5 * 32);
    __syncthreads();
    if (a.plans) {
      const int nvec = (int)((plan_bytes() + tab_bytes) >> 4);
      uint4* dst = reinterpret_cast<uint4*>(a.plans + (size_t)b * a.plan_stride);
      const uint4* src = reinterpret_cast<const uint4*>(smem);
      for (int i = tid; i < nvec; i += NTHREADS) dst[i] = src[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------ anti-alias prefilters

// _apply_antialias_filter (image_geometric_cv2.py:47-62) for the samples plan_kernel marked (Plan::prefilter): the canvas --
// the zero-padded crop, or the cv2.warpAffine image of a rotated sample -- smoothed at its own resolution into the sample's
// workspace region; the fused kernel then resizes it with cv2's INTER_LINEAR (:76-81) like any crop.
//   gaussian: the reference writes cv2.GaussianBlur(img, (0, 0), ks, ks, cv2.BORDER_REPLICATE) with ks = 0.5 / scale, and
//     Python binds that as sigmaX = ks, dst = ks (ignored), sigmaY = BORDER_REPLICATE = 1.0, default border: a
//     cvRound(6 ks + 1) | 1 tap kernel along rows, 7 taps (sigma 1) along columns, BORDER_REFLECT_101.  That is what is
//     reproduced here (oracle/cv2_model.py:REFERENCE_GAUSSIAN_SIGMA_Y).  On 8-bit images the blur is exact integer work:
//     8-bit fixed point taps (error-diffused so that they sum to 256), rows then columns, (sum + 2^15) >> 16
//     (oracle/cv2_model.py:gaussian_kernel_fixed, gaussian_blur_u8);
//   hamming:  cv2.sepFilter2D(img, -1, k, k), BORDER_REFLECT_101, float32: rows s = k0 p0, s = fma(ki, pi, s); columns the
//     same, or for a kernel cv2 classifies as symmetric s = kc hc, s = fma(k[c+i], h[c-i] + h[c+i], s); rint, saturate
//     (oracle/cv2_model.py:sep_filter_u8 -- the arithmetic of cv2's vector loops).  The window comes from the caller's table
//     (B200AugFusedArgs::hamming_taps): whether cv2 sees it as symmetric hangs on the last bit of the host's cos().
// A non-default configuration (the samplers always ask for area / linear, geometric.py:76-77), so this is a plain tiled
// kernel: PF_CTAS CTAs per sample walk its 32 x 32 canvas tiles -- stage tile + halo through canvas_px() with the border
// rule applied to the canvas coordinates, horizontal pass into shared memory, vertical pass, store.
constexpr int PF_TILE = 32;
constexpr int PF_CTAS = 8;  // CTAs per sample
constexpr int PF_MAXR = B200AUG_PREFILTER_MAX_TAPS / 2;
constexpr int PF_SIDE = PF_TILE + 2 * PF_MAXR;  // 94: staged rows / columns at the widest kernel
constexpr double PF_GAUSS_SIGMA_Y = 1.0;        // (see above)

__device__ __forceinline__ int border_reflect101(int p, int len) {  // cv::borderInterpolate(BORDER_REFLECT_101)
  if ((unsigned)p < (unsigned)len) return p;
  if (len == 1) return 0;
  do {
    if (p < 0) p = -p;
    else p = 2 * len - p - 2;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}

// the n fixed-point taps (bit patterns of int32) of cv2's 8-bit Gaussian; one thread
__device__ void gaussian_taps_fixed(double sigma, int n, float* ktab) {
  const int r = n >> 1;
  const double s2 = -0.5 / (sigma * sigma);
  double vals[PF_MAXR + 1], sum = 0.0;
  for (int i = 0; i < r; ++i) {
    const double x = (double)(i - r);
    vals[i] = exp(s2 * x * x);
    sum += vals[i];
  }
  const double norm = 1.0 / (2.0 * sum + 1.0);
  double err = 0.0;
  int tot = 0;
  for (int i = 0; i < r; ++i) {
    const double adj = vals[i] * norm * 256.0 + err;
    const int v0 = rint_d2i(adj);
    err = adj - (double)v0;
    ktab[i] = __int_as_float(v0);
    ktab[n - 1 - i] = __int_as_float(v0);
    tot += v0;
  }
  ktab[r] = __int_as_float(256 - 2 * tot);
}

__global__ void __launch_bounds__(NTHREADS) prefilter_kernel(const __grid_constant__ KArgs K) {
  const B200AugFusedArgs& a = K.a;
  __shared__ __align__(16) unsigned char plan_raw[(sizeof(Plan) + 15) & ~size_t(15)];
  __shared__ uint8_t staged[PF_SIDE][PF_SIDE + 2];
  __shared__ float hbuf[PF_SIDE][PF_TILE + 1];  // gaussian: int32 bit patterns
  __shared__ float ktx[B200AUG_PREFILTER_MAX_TAPS + 1], kty[B200AUG_PREFILTER_MAX_TAPS + 1];  // gaussian: int32 bit patterns
  const int b = blockIdx.x / PF_CTAS, part = blockIdx.x % PF_CTAS, tid = threadIdx.x;
  {
    const uint4* rec = reinterpret_cast<const uint4*>(a.plans + (size_t)b * a.plan_stride);
    uint4* dst = reinterpret_cast<uint4*>(plan_raw);
    for (int i = tid; i < (int)(plan_bytes() >> 4); i += NTHREADS) dst[i] = rec[i];
  }
  __syncthreads();
  const Plan& P = *reinterpret_cast<const Plan*>(plan_raw);
  if (!P.prefilter || P.status != B200AUG_S_OK || P.cv_ptr == nullptr) return;
  const bool gauss = P.prefilter == B200AUG_DOWN_GAUSSIAN;
  const int nx = P.pf_n, ny = gauss ? gaussian_ksize_u8(PF_GAUSS_SIGMA_Y) : nx;
  const int rx = nx >> 1, ry = ny >> 1, cw = P.cw, ch = P.ch;
  bool sym = false;
  if (gauss) {
    if (tid == 0) gaussian_taps_fixed(0.5 / P.pf_sf, nx, ktx);
    if (tid == 32) gaussian_taps_fixed(PF_GAUSS_SIGMA_Y, ny, kty);
  } else {
    if (tid < nx) ktx[tid] = kty[tid] = a.hamming_taps[(size_t)rx * 64 + tid];
    sym = (a.hamming_sym_mask >> rx) & 1ull;
  }
  __syncthreads();
  const int side_x = PF_TILE + 2 * rx, side_y = PF_TILE + 2 * ry;
  const int ntx = (cw + PF_TILE - 1) / PF_TILE, nty = (ch + PF_TILE - 1) / PF_TILE;
  for (int t = part; t < ntx * nty; t += PF_CTAS) {
    const int tx0 = (t % ntx) * PF_TILE, ty0 = (t / ntx) * PF_TILE;
    // ---- stage tile + halo (canvas coordinates folded back into the canvas by the border rule)
    for (int i = tid; i < side_x * side_y; i += NTHREADS) {
      const int yy = i / side_x, xx = i - yy * side_x;
      staged[yy][xx] = (uint8_t)canvas_px(P, border_reflect101(tx0 - rx + xx, cw), border_reflect101(ty0 - ry + yy, ch));
    }
    __syncthreads();
    // ---- rows
    for (int i = tid; i < side_y * PF_TILE; i += NTHREADS) {
      const int yy = i / PF_TILE, x = i % PF_TILE;
      if (gauss) {
        int sacc = 0;
        for (int k = 0; k < nx; ++k) sacc += __float_as_int(ktx[k]) * (int)staged[yy][x + k];
        hbuf[yy][x] = __int_as_float(sacc);
      } else {
        float sacc = __fmul_rn(ktx[0], (float)staged[yy][x]);
        for (int k = 1; k < nx; ++k) sacc = __fmaf_rn(ktx[k], (float)staged[yy][x + k], sacc);
        hbuf[yy][x] = sacc;
      }
    }
    __syncthreads();
    // ---- columns
    for (int i = tid; i < PF_TILE * PF_TILE; i += NTHREADS) {
      const int y = i / PF_TILE, x = i % PF_TILE;
      int q;
      if (gauss) {
        int sacc = 0;
        for (int k = 0; k < ny; ++k) sacc += __float_as_int(kty[k]) * __float_as_int(hbuf[y + k][x]);
        q = (sacc + (1 << 15)) >> 16;
      } else {
        float sacc;
        if (sym) {
          sacc = __fmul_rn(kty[ry], hbuf[y + ry][x]);
          for (int k = 1; k <= ry; ++k) sacc = __fmaf_rn(kty[ry + k], __fadd_rn(hbuf[y + ry - k][x], hbuf[y + ry + k][x]), sacc);
        } else {
          sacc = __fmul_rn(kty[0], hbuf[y][x]);
          for (int k = 1; k < ny; ++k) sacc = __fmaf_rn(kty[k], hbuf[y + k][x], sacc);
        }
        q = __float2int_rn(sacc);
      }
      if (tx0 + x < cw && ty0 + y < ch) P.cv_ptr[(size_t)(ty0 + y) * P.cv_pitch + tx0 + x] = (uint8_t)min(max(q, 0), 255);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ canvas workers

// cv2.warpAffine of every rotated sample's canvas into its workspace region.  Some CTAs of the fused kernel's grid take
// this role instead of a sample (fused_augment_kernel decides which): work items = (rotated sample, quarter of its canvas
// tiles) handed out by an atomic counter, so the stage is spread over the whole machine instead of sitting in front of its
// own sample's resampling (where one rotated + blurred sample used to bound the whole launch).  The samples are found by
// scanning the flag bytes plan_kernel left behind the plan records; every finished item bumps its sample's completion
// counter, which the sample's own CTAs wait for (a worker is always among the first CTAs of the grid and any worker can
// take any item, so the wait cannot deadlock).
//
// One item is a pipeline over 64 x 32 canvas tiles: the canvas has the resolution of the source (image_geometric_cv2.py:
// 121-124 sizes it so), so a tile's source footprint is a rotated 64 x 32 rectangle; its bounding box (<= 74 rows x 81 px)
// is staged in shared memory by all threads with 16-byte cp.async copies, WK_NSTAGE - 1 tiles ahead of the tile being
// gathered, so the loads' latency never sits in front of the arithmetic.  Box rows outside the frame are zero-filled by the
// copy itself (src-size 0 = BORDER_CONSTANT), columns outside it by a fix-up pass over the staged box, so the gather loop
// is the same for every tile: per 32 pixels 2 adds, 3 address ops, 4 byte loads, the 11-op fixed-point blend, 1 store.
// Staged layout: box row r lands at offset B r + 16 ((c0 + r pm) >> 4), c0 = alignment shift
// of row 0, pm = pitch mod 16, so that source pixel (iy, ix) sits at c0 + (iy - by0)(B + pm) + (ix - bx0): linear, no
// per-row alignment fix-up in the gather; B is 96 or 112, whichever spreads one canvas row's taps over more banks.
#ifndef B200AUG_WK_CHUNKS
#define B200AUG_WK_CHUNKS 4
#endif
constexpr int WK_CHUNKS = B200AUG_WK_CHUNKS;  // work items per rotated sample
constexpr int WK_NSTAGE = 3;                 // staged tiles in flight per worker
constexpr int WK_MAX_ITEM_TILES = 32;        // tiles of one work item whose geometry is worked out in one go
constexpr int WT2_W = 64, WT2_H = 32;        // canvas tile of one pipeline step
constexpr int WT2_ROWS = 72;                 // tallest staged bounding box (a 64 x 32 tile turned by 45 degrees: 71 rows)
constexpr int WT2_NCH = 6;                   // 16-byte chunks staged per box row (box width + alignment shift <= 96)
constexpr int WT2_B0 = 96, WT2_B1 = 112;     // candidate row strides (+ pitch mod 16)
constexpr int WT2_STAGE = ((WT2_ROWS - 1) * (WT2_B1 + 15) + 15 + 16 * WT2_NCH + 15) & ~15;
constexpr int WK_OFF_RTAB = DT_CAP * 8, WK_OFF_LIST = 2 * DT_CAP * 8, WK_OFF_SH = WK_OFF_LIST + WK_SLICE * 2,
              WK_OFF_STAGE = WK_OFF_SH + 128 + WK_MAX_ITEM_TILES * 48, WK_AREA_BYTES = WK_OFF_STAGE + WK_NSTAGE * WT2_STAGE;

static_assert(WK_AREA_BYTES == (int)WORKER_AREA, "smem_layout reserves WORKER_AREA bytes for a canvas worker");

struct TileMeta {
  int x_lo, y_lo, tw, th, bx0, bx1, by0, c0, nrows, mode;  // mode 0 staged | 1 staged, columns outside the frame | 2 not staged
  int pad[2];
};
static_assert(sizeof(TileMeta) == 48, "WK_OFF_STAGE");

// geometry of canvas tile (tx, ty): bounding box of its taps in the source, staging decision.  One thread.
__device__ __forceinline__ void wk_tile_meta(const Plan& P, const int2* dtab, const int2* rtab, int tx, int ty, TileMeta* meta) {
  const int x_lo = tx * WT2_W, y_lo = ty * WT2_H;
  const int tw = min(WT2_W, P.cw - x_lo), th = min(WT2_H, P.ch - y_lo);
  // bounding box of the taps from the four tile corners (the fixed-point map is monotone in x and in y)
  const int2 ra = rtab[y_lo], rb = rtab[y_lo + th - 1], dl = dtab[x_lo], dr = dtab[x_lo + tw - 1];
  const int ix0 = (ra.x + dl.x) >> 10, ix1 = (ra.x + dr.x) >> 10, ix2 = (rb.x + dl.x) >> 10, ix3 = (rb.x + dr.x) >> 10;
  const int iy0 = (ra.y + dl.y) >> 10, iy1 = (ra.y + dr.y) >> 10, iy2 = (rb.y + dl.y) >> 10, iy3 = (rb.y + dr.y) >> 10;
  const int bx0 = min(min(ix0, ix1), min(ix2, ix3)), bx1 = max(max(ix0, ix1), max(ix2, ix3)) + 1;
  const int by0 = min(min(iy0, iy1), min(iy2, iy3)), by1 = max(max(iy0, iy1), max(iy2, iy3)) + 1;
  const int nrows = by1 - by0 + 1, bbw = bx1 - bx0 + 1;
  const uintptr_t gbase = reinterpret_cast<uintptr_t>(P.src) + (ptrdiff_t)by0 * P.pitch + bx0;  // (may lie outside the frame)
  const bool staged = nrows <= WT2_ROWS && bbw + 15 <= 16 * WT2_NCH;
  TileMeta m;
  m.x_lo = x_lo; m.y_lo = y_lo; m.tw = tw; m.th = th; m.bx0 = bx0; m.bx1 = bx1; m.by0 = by0; m.c0 = (int)(gbase & 15); m.nrows = nrows;
  m.mode = !staged ? 2 : ((bx0 < 0 || bx1 >= P.sw) ? 1 : 0);
  m.pad[0] = m.pad[1] = 0;
  *meta = m;
}

// the cp.async copies of a tile's source box into its stage: thread = (16-byte chunk k = tid & 7, rows tid >> 3 + 32 j)
__device__ __forceinline__ void wk_issue_tile(const Plan& P, const TileMeta& m, int bstride, uint32_t buf32, int tid) {
  if (m.mode == 2) return;
  const int k = tid & 7;
  if (k >= WT2_NCH) return;
  const int pitch = P.pitch, sh = P.sh, pm = pitch & 15, by0 = m.by0, c0 = m.c0, nrows = m.nrows;
  const uintptr_t src = reinterpret_cast<uintptr_t>(P.src);
  // copies must stay inside the 16-byte blocks that hold frame bytes; everything else is zero-filled
  const uintptr_t frame_lo = src & ~uintptr_t(15), frame_end = src + (size_t)(sh - 1) * pitch + P.sw;
  int row = tid >> 3;
  uintptr_t grow = src + (ptrdiff_t)(by0 + row) * pitch + m.bx0;      // first byte of box row `row`
  uint32_t drow = buf32 + (uint32_t)(row * bstride + 16 * k), sft = (uint32_t)(c0 + row * pm);
  const uint32_t drow_step = 32u * (uint32_t)bstride, sft_step = 32u * (uint32_t)pm;
  const ptrdiff_t grow_step = (ptrdiff_t)32 * pitch;
#pragma unroll 1
  for (; row < nrows; row += 32) {
    const uintptr_t g = (grow & ~uintptr_t(15)) + 16u * (uint32_t)k;
    const bool ok = (unsigned)(by0 + row) < (unsigned)sh && g >= frame_lo && g < frame_end;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + ((sft >> 4) << 4)), "l"(ok ? g : frame_lo), "r"(ok ? 16 : 0)
                 : "memory");
    grow += grow_step;
    drow += drow_step;
    sft += sft_step;
  }
}

__device__ __forceinline__ void wk_compute_tile(const Plan& P, const int2* dtab, const int2* rtab, const TileMeta* meta, int bstride,
                                                uint8_t* buf, int warp, int lane, int tid) {
  uint8_t* const canvas = P.cv_ptr;
  const int spitch = P.cv_pitch;
  const TileMeta m = *meta;  // (registers: the gather loop's inline assembly would otherwise force reloads)
  if (m.mode == 2) {
    // (a box too large to stage: only for strongly down-scaling canvases, which image_geometric_cv2.py never builds)
    const uint8_t* const src = P.src;
    const int pitch = P.pitch, sw = P.sw, sh = P.sh;
    for (int p = tid; p < m.tw * m.th; p += NTHREADS) {
      const int yy = p / m.tw, xx = p - yy * m.tw;
      const int2 ro = rtab[m.y_lo + yy], d = dtab[m.x_lo + xx];
      const int X = (ro.x + d.x) >> 5, Y = (ro.y + d.y) >> 5, ix = X >> 5, iy = Y >> 5;
      const bool r0 = (unsigned)iy < (unsigned)sh, r1 = (unsigned)(iy + 1) < (unsigned)sh;
      const bool q0 = (unsigned)ix < (unsigned)sw, q1 = (unsigned)(ix + 1) < (unsigned)sw;
      const uint8_t* g = src + (ptrdiff_t)iy * pitch + ix;
      const int p00 = (r0 && q0) ? __ldg(g) : 0, p01 = (r0 && q1) ? __ldg(g + 1) : 0;
      const int p10 = (r1 && q0) ? __ldg(g + pitch) : 0, p11 = (r1 && q1) ? __ldg(g + pitch + 1) : 0;
      canvas[(size_t)(m.y_lo + yy) * spitch + m.x_lo + xx] = (uint8_t)bilinear_q5(p00, p01, p10, p11, X & 31, Y & 31);
    }
    return;
  }
  const int rstride = bstride + (P.pitch & 15);
  if (m.mode == 1) {
    // BORDER_CONSTANT: the staged bytes of columns outside [0, sw) belong to neighbouring rows -- zero them
    const int nl = max(0, min(m.bx1, -1) - m.bx0 + 1);
    const int xr = max(P.sw, m.bx0), nr = max(0, m.bx1 - xr + 1), nz = nl + nr;
    for (int i = tid; i < m.nrows * nz; i += NTHREADS) {
      const int r = i / nz, j = i - r * nz;
      const int x = (j < nl) ? m.bx0 + j : xr + (j - nl);
      buf[m.c0 + r * rstride + (x - m.bx0)] = 0;
    }
    __syncthreads();
  }
  // shared-memory address of source pixel (iy, ix) = buf + c0 + (iy - by0) * rstride + (ix - bx0).  The low 22 bits of the
  // base, c0 - bx0 and -by0 ride in the column deltas (scaled by 1024 like the coordinates); the high bits of the base (in
  // a cluster the shared-memory window of a CTA carries its rank up there) come back in through the funnel shift that
  // extracts the integer coordinate, so the address costs shift + shift + multiply-add.
  const uint32_t buf32 = smem_u32(buf), hi22 = buf32 >> 22;
  const uint32_t K = (buf32 & 0x3FFFFFu) + (uint32_t)(m.c0 - m.bx0);
  int dxk[2], dyk[2];
  bool act[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int xi = 32 * h + lane;
    const int2 d = dtab[m.x_lo + min(xi, m.tw - 1)];  // (idle lanes repeat the last column: their addresses stay valid)
    dxk[h] = d.x + (int)(K << 10);
    dyk[h] = d.y - (m.by0 << 10);
    act[h] = xi < m.tw;
  }
  const int2* const rorow = rtab + m.y_lo + warp;
  uint8_t* out = canvas + (size_t)(m.y_lo + warp) * spitch + m.x_lo + lane;
  const size_t out_step = (size_t)NWARPS * spitch;
  const int n_it = (m.th - warp + NWARPS - 1) / NWARPS;  // rows warp, warp + 8, ...: at most WT2_H / NWARPS = 4 of them
  static_assert(WT2_H == 4 * NWARPS, "the row loop below is unrolled four times");
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (it >= n_it) break;
    const int2 ro = rorow[it * NWARPS];
    const int rox = ro.x, roy = ro.y;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t sx = (uint32_t)rox + (uint32_t)dxk[h];  // 1/1024 px, offset by K
      const int sy = roy + dyk[h];
      const uint32_t a0 = (uint32_t)(sy >> 10) * (uint32_t)rstride + __funnelshift_r(sx, hi22, 10), a1 = a0 + (uint32_t)rstride;
      uint32_t p00, p01, p10, p11;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p00) : "r"(a0) : "memory");
      asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(p01) : "r"(a0) : "memory");
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(p10) : "r"(a1) : "memory");
      asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(p11) : "r"(a1) : "memory");
      // bilinear_q5 with the 1/32 fractions left in place (fx32 = 32 fx): every term carries a factor 1024
      const int fx = sx & 0x3E0, fy = sy & 0x3E0;
      const int top = (1024 - fx) * (int)p00 + fx * (int)p01, bot = (1024 - fx) * (int)p10 + fx * (int)p11;
      const int q = ((1024 - fy) * top + fy * bot + (512 << 10)) >> 20;
      if (act[h]) out[32 * h] = (uint8_t)q;
    }
    out += out_step;
  }
}

__device__ __noinline__ void canvas_worker(const B200AugFusedArgs& a, Plan& P, unsigned char* area, uint64_t* tr) {
  int2* const dtab = reinterpret_cast<int2*>(area);
  int2* const rtab = reinterpret_cast<int2*>(area + WK_OFF_RTAB);
  uint16_t* const list = reinterpret_cast<uint16_t*>(area + WK_OFF_LIST);
  int* const sh = reinterpret_cast<int*>(area + WK_OFF_SH);  // [0] n_list, [1] item, [2..9] warp counts, [10..17] bases
  TileMeta* const meta = reinterpret_cast<TileMeta*>(area + WK_OFF_SH + 128);  // [WK_MAX_ITEM_TILES], the current item's tiles
  uint8_t* const stage = area + WK_OFF_STAGE;
  const uint32_t stage32 = smem_u32(stage);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PlanTail tail = plan_tail(a.plans, a.batch, a.plan_stride);
  int n_done = 0;
  for (int slice = 0, base = 0; base < a.batch && slice < WK_MAX_SLICES; ++slice, base += WK_SLICE) {
    // ---- the rotated samples of this slice of the launch order, in that order (every worker builds the same list)
    if (tid == 0) sh[0] = 0;
    __syncthreads();
    const int n_here = min(WK_SLICE, a.batch - base);
    for (int i0 = 0; i0 < n_here; i0 += NTHREADS) {
      const int i = i0 + tid;
      const bool on = i < n_here && tail.flags[a.order ? a.order[base + i] : base + i] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, on);
      if (lane == 0) sh[2 + warp] = __popc(m);
      __syncthreads();
      if (tid == 0) {
        int run = sh[0];
        for (int w = 0; w < NWARPS; ++w) { sh[10 + w] = run; run += sh[2 + w]; }
        sh[0] = run;
      }
      __syncthreads();
      if (on) list[sh[10 + warp] + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
    }
    __syncthreads();
    const int n_items = sh[0] * WK_CHUNKS;
    // ---- work stealing over (sample, chunk)
    for (;;) {
      const long long i0 = tr ? clock64() : 0;
      if (tid == 0) sh[1] = (int)atomicAdd(&tail.counters[slice], 1u);
      __syncthreads();
      const int item = sh[1];
      if (item >= n_items) break;
      const long long i1 = tr ? clock64() : 0;
      const int pos = base + list[item / WK_CHUNKS], chunk = item % WK_CHUNKS;
      const int b = a.order ? a.order[pos] : pos;
      {
        const uint4* rec = reinterpret_cast<const uint4*>(a.plans + (size_t)b * a.plan_stride);
        uint4* dst = reinterpret_cast<uint4*>(&P);
        for (int i = tid; i < (int)(plan_bytes() >> 4); i += NTHREADS) dst[i] = rec[i];
      }
      __syncthreads();
      // per-column / per-row fixed-point terms of cv2.warpAffine (oracle/cv2_model.py:warp_affine_linear_u8)
      const int cw = P.cw, ch = P.ch;
      for (int i = tid; i < cw + ch; i += NTHREADS) {
        if (i < cw) {
          dtab[i] = make_int2(rint_d2i(__dmul_rn(__dmul_rn(P.mi[0], (double)i), 1024.0)),
                              rint_d2i(__dmul_rn(__dmul_rn(P.mi[3], (double)i), 1024.0)));
        } else {
          const double y = (double)(i - cw);
          rtab[i - cw] = make_int2(rint_d2i(__dmul_rn(__dadd_rn(__dmul_rn(P.mi[1], y), P.mi[2]), 1024.0)) + 16,
                                   rint_d2i(__dmul_rn(__dadd_rn(__dmul_rn(P.mi[4], y), P.mi[5]), 1024.0)) + 16);
        }
      }
      // Row stride of the staged box: whichever candidate spreads one canvas row's 32 taps over more shared-memory banks.
      // The taps walk a straight line (mi[0] px right, mi[3] px down per canvas pixel), so the bank pattern depends on
      // slope and stride only (every warp evaluates it; the result is the same).
      int bstride = WT2_B0;
      {
        const int pm = P.pitch & 15;
        const int ix = (int)floor(P.mi[0] * (double)lane), iy = (int)floor(P.mi[3] * (double)lane) + 64;
        int best = INT_MAX;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int e = (c ? WT2_B1 : WT2_B0) + pm;
          const unsigned word = (unsigned)(iy * e + ix) >> 2, bank = word & 31u;
          const unsigned same_word = __match_any_sync(0xffffffffu, word), same_bank = __match_any_sync(0xffffffffu, bank);
          const unsigned leaders = __ballot_sync(0xffffffffu, (__ffs(same_word) - 1) == lane);  // one lane per distinct word
          const int degree = __reduce_max_sync(0xffffffffu, __popc(same_bank & leaders));      // wavefronts of this load
          if (degree < best) { best = degree; bstride = c ? WT2_B1 : WT2_B0; }
        }
      }
      __syncthreads();
      const int tiles_x = (cw + WT2_W - 1) / WT2_W, tiles_y = (ch + WT2_H - 1) / WT2_H, n_tiles = tiles_x * tiles_y;
      const int it_begin = (chunk * n_tiles) / WK_CHUNKS, it_end = ((chunk + 1) * n_tiles) / WK_CHUNKS;
      long long i2 = 0;
      // ---- the tile pipeline: the copies of tile t + WK_NSTAGE - 1 are issued before tile t is gathered.  The geometry of
      // the tiles (WK_MAX_ITEM_TILES at a time: all of the item's unless the canvas is very large) is worked out first, one
      // thread per tile.
#pragma unroll 1
      for (int t_begin = it_begin; t_begin < it_end; t_begin += WK_MAX_ITEM_TILES) {
      const int t_end = min(t_begin + WK_MAX_ITEM_TILES, it_end);
      if (tid < t_end - t_begin) {
        const int t = t_begin + tid, ty = t / tiles_x;
        wk_tile_meta(P, dtab, rtab, t - ty * tiles_x, ty, meta + tid);
      }
      __syncthreads();
#pragma unroll 1
      for (int s2 = 0; s2 < WK_NSTAGE - 1; ++s2) {
        if (t_begin + s2 < t_end) wk_issue_tile(P, meta[s2], bstride, stage32 + s2 * WT2_STAGE, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (t_begin == it_begin) i2 = tr ? clock64() : 0;
      int sl = 0, sn = WK_NSTAGE - 1;  // stage of tile t / of tile t + WK_NSTAGE - 1
#pragma unroll 1
      for (int t = t_begin; t < t_end; ++t) {
        const int i = t - t_begin;
        const long long c0 = tr ? clock64() : 0;
        if (t + WK_NSTAGE - 1 < t_end) wk_issue_tile(P, meta[i + WK_NSTAGE - 1], bstride, stage32 + sn * WT2_STAGE, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
        static_assert(WK_NSTAGE == 3, "wait_group immediate");
        const long long c1 = tr ? clock64() : 0;
        asm volatile("cp.async.wait_group 2;" ::: "memory");  // this thread's copies of tile t have landed ...
        __syncthreads();                                      // ... and everybody else's
        const long long c2 = tr ? clock64() : 0;
        wk_compute_tile(P, dtab, rtab, meta + i, bstride, stage + sl * WT2_STAGE, warp, lane, tid);
        const long long c3 = tr ? clock64() : 0;
        __syncthreads();  // the stage is free again
        if (tr && tid == 0) {  // (profiling: cycles in issue / wait / gather / barrier, tiles)
          const long long c4 = clock64();
          tr[8] += (uint64_t)(c1 - c0); tr[9] += (uint64_t)(c2 - c1); tr[10] += (uint64_t)(c3 - c2); tr[11] += (uint64_t)(c4 - c3); tr[12] += 1;
        }
        sl = (sl + 1 == WK_NSTAGE) ? 0 : sl + 1;
        sn = (sn + 1 == WK_NSTAGE) ? 0 : sn + 1;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");  // (only empty groups are left)
      }
      const long long i3 = tr ? clock64() : 0;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();   // every thread's canvas pixels are written ...
      if (tid == 0) {
        __threadfence();  // ... and visible device-wide (cumulative over the barrier) before the completion counter moves
        atomicAdd(&tail.done[b], 1u);
        if (tr) {  // (profiling: cycles for the work-item fetch / plan + tables + tile geometry + first copies / tail, items)
          const long long i4 = clock64();
          tr[13] += (uint64_t)(i1 - i0); tr[14] += (uint64_t)(i2 - i1); tr[15] += (uint64_t)(i4 - i3); tr[7] += 1;
        }
      }
      ++n_done;
    }
    __syncthreads();
  }
  if (tr && tid == 0) {
    tr[4] = globaltimer_ns();
    tr[2] = (uint64_t)n_done;
  }
}

// ------------------------------------------------------------------------------------------------ the kernel

// dep_mode: 1 = launched behind plan_kernel with a programmatic dependent launch: every CTA waits for it
// (griddepcontrol.wait) before it reads its plan record; 0 = plain stream order (the plan is built in here)
template <bool TRACE>
__global__ void __launch_bounds__(NTHREADS, 3) fused_augment_kernel(const __grid_constant__ KArgs K, int cap, int dep_mode) {
  const B200AugFusedArgs& a = K.a;
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t cr_u, cl_u;
  cluster_info(cr_u, cl_u);
  const int cr = (int)cr_u, cl = (int)cl_u;  // rank in / size of the cluster that shares this sample
  // launch order -> sample: the caller may schedule expensive samples (blurred, noisy) first so that the cheap ones
  // fill the tail of the grid
  const int cs = 31 - __clz(cl);  // log2 of the cluster size (1, 2, 4 or 8: checked at launch), divisions become shifts
  const int ow = a.out_w, oh = a.out_h, npix = ow * oh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // a.warp_ctas CTAs of the grid (a multiple of the cluster size) produce the rotated samples' canvases instead of taking a
  // sample.  They are interleaved with the sample clusters at the head of the grid -- worker, sample, sample, ... -- so
  // that the first wave puts one worker next to two sample CTAs on every SM (three workers on one SM would all queue on
  // its shared-memory pipe); cluster 0 is always a worker.
  bool worker;
  int slot;
  {
    const int nw = a.warp_ctas >> cs, ns = a.batch, g = min(nw, ns >> 1), c = (int)(blockIdx.x >> cs);
    if (c < 3 * g) {
      worker = (c % 3) == 0;
      slot = c - (c / 3 + 1);
    } else {
      const int rem = c - 3 * g;
      worker = rem < nw - g;
      slot = 2 * g + rem - (nw - g);
    }
  }
  const int b = worker ? 0 : (a.order ? a.order[slot] : slot);
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
  uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + L.off_bars);  // cluster exchange barrier

  trace_mark<TRACE>(a, 0);
  if (TRACE) {
    if (a.trace_out && tid == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.trace_out[(size_t)blockIdx.x * 16 + 5] = smid;
    }
  }
  if (worker) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // plan_kernel is done (no-op without a programmatic launch)
    canvas_worker(a, P, smem + L.off_tile, (TRACE && a.trace_out) ? a.trace_out + (size_t)blockIdx.x * 16 : nullptr);
    return;
  }
  if (tid == 0) {
    mbar_init(xbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- the plan: precomputed by plan_kernel or built here (the labels are plan_kernel's business either way) -----------
  const bool preplanned = a.plans != nullptr;
  if (preplanned) {
    if (dep_mode == 1) asm volatile("griddepcontrol.wait;" ::: "memory");  // plan_kernel is done
    const unsigned char* rec = a.plans + (size_t)b * a.plan_stride;
    const int nv_plan = (int)(plan_bytes() >> 4), nv_tab = (int)(plan_tab_bytes(ow, oh) >> 4);
    for (int i = tid; i < nv_plan + nv_tab; i += NTHREADS) {
      const uint32_t d = (i < nv_plan) ? smem_u32(smem) + 16u * i : smem_u32(smem + L.off_tabs) + 16u * (i - nv_plan);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(rec + 16 * (size_t)i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    build_plan(a, b, P, warp, lane, false, reinterpret_cast<PlanCore*>(lut));  // (the LUT area is idle until the plan exists)
  }
  __syncthreads();

  trace_mark<TRACE>(a, 1);
  // Samples whose photometric chain is a single point function (no equalize, blur or noise: more than half of the training
  // draws, and every evaluation sample) get their LUT now: the INTER_AREA pass can then write float32 output directly
  // (`direct` below).  The cluster barrier in front of the resampling orders these writes before their first use.
  const bool lut_early = (a.flags & B200AUG_F_NORMALIZE) && a.image_f32_out && P.blur_pos < 0 && P.eq_pos < 0 && !P.any_noise;
  if (lut_early) {
    float xv = apply_point_ops(P, __fmul_rn((float)tid, 0.00390625f), 0, P.n_ops, eq_lut);
    if ((a.flags & B200AUG_F_PHOTOMETRIC) && a.photo.clip) xv = fminf(fmaxf(xv, 0.f), 1.f);
    if (a.flags & B200AUG_F_WHITEN) xv = __fsub_rn(xv, 0.5f);
    lut[tid] = xv;
  }
  // this CTA's band of output rows, and the bytes the other CTAs of the cluster will bulk-copy into this tile
  const int rows_lo = (cr * oh) >> cs, rows_hi = ((cr + 1) * oh) >> cs;
  int rx_bytes = 0;
  if (cl > 1 && P.status == B200AUG_S_OK && P.rot_dir == 0) {
    for (int q = 0; q < cl; ++q) {
      if (q == cr) continue;
      const int lo = ((q * oh) >> cs) * ow, hi = (((q + 1) * oh) >> cs) * ow, alo = min((lo + 15) & ~15, hi), ahi = max(hi & ~15, alo);
      rx_bytes += ahi - alo;
    }
    if (tid == 0 && rx_bytes)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(xbar)), "r"(rx_bytes) : "memory");
  }

  // ---- resize tables (unless they came with the precomputed plan) -------------------------------------------------------
  const int rs = P.rs_mode;
  if (!preplanned) build_tables(a, P, T, tid, lane, NTHREADS);
  trace_mark<TRACE>(a, 8);
  if (a.flags & B200AUG_F_PHOTOMETRIC) {
    hist8[tid] = 0;
    binhist[tid] = 0;
  }
  // ---- rotated samples: the canvas workers leave the cv2.warpAffine canvas in the workspace; from here on the sample is a
  // plain crop of it (image_geometric_cv2.py:121-134: warpAffine at source resolution, then cv2.resize)
  // (prefiltered samples: prefilter_kernel, ahead of this kernel in the stream, left the smoothed canvas there)
  if ((P.src_mode == SRC_WARP || P.prefilter) && P.cv_ptr != nullptr) {
    if (tid == 0 && !P.prefilter) {
      const uint32_t* done = plan_tail(a.plans, a.batch, a.plan_stride).done + b;
      uint32_t v;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done) : "memory");
        if (v >= (uint32_t)WK_CHUNKS) break;
        __nanosleep(256);
      }
    }
    __syncthreads();
    if (tid == 0) {
      P.src = P.cv_ptr;
      P.pitch = P.cv_pitch;
      P.sw = P.cv_pitch;  // columns [cw, pitch) are padding the row copies may touch but no tap ever reads
      P.sh = P.ch + 1;
      P.x0 = 0;
      P.y0 = 0;
      P.src_mode = SRC_CROP;
    }
  }
  trace_mark<TRACE>(a, 7);
  __syncthreads();

  trace_mark<TRACE>(a, 2);
  trace_mark<TRACE>(a, 6);
  // ---- resample into the uint8 tile -----------------------------------------------------------------------
  const TileMap tm = make_tile_map(P, ow, oh);
  // columns [0, ow_band) are resampled by the warps' bands, a short last group by the per-pixel path
  const int ow_tail = ow % (32 * RMAX);
  const int ow_band = (ow_tail != 0 && ow_tail <= TAIL_COLS) ? ow - ow_tail : ow;
  const bool lin = rs == RS_LINEAR;  // cv2's 2-tap kernels stream through the same bands
  bool fast = (P.status == B200AUG_S_OK) && (((rs == RS_AREA || rs == RS_AREA_INT) && P.kx <= KMAX) || lin) && (P.src_mode != SRC_WARP) &&
              ow_band > 0;
  if (fast && !lin) {
    // the CTA's canvas rows must fit its vertical-pass program
    // (the widest share of any CTA of the cluster: `fast` -- and with it `direct` -- must come out the same in all of them)
    const int rows_per_cta = ((oh + cl - 1) >> cs) + 1;
    const int sy_ceil = (rs == RS_AREA_INT) ? P.iscale_y : (int)ceil(P.scale_y);
    if ((rows_per_cta + 1) * sy_ceil + 2 > NWARPS * ROWPROG_CAP) fast = false;
  }
  if (fast) {
    // every column group's canvas segment must fit the per-warp row buffer (only staged rows need it)
    for (int g0 = 0; g0 < ow_band; g0 += 32 * RMAX) {
      const int glast = min(g0 + 32 * RMAX, ow_band) - 1;
      const int seg = T.start[glast] + (lin ? 2 : (T.n[glast] & 0xffff)) - T.start[g0];
      if (RING_D * ring_slot_bytes(seg) > cap + ROWBUF_SLACK || seg + 30 > 64 * 16) fast = false;
    }
  }
  // direct: no tile, no exchange, no output pass -- the resampling pass writes the float32 crop itself
  const bool direct = fast && lut_early && ((rs == RS_AREA && P.fin == 0) || lin) && P.rot_dir == 0;
  // Every CTA of the cluster must be running (its exchange barrier initialised) before the tile exchange touches its shared
  // memory: arrive now, wait only in front of the exchange -- the resampling in between never waits for the partner.
  // (direct samples have no exchange; `direct` comes out the same in all CTAs of the cluster.)
  if (cl > 1 && !direct) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  if (fast) {
    const int kx = P.kx;
    area_precompute(kx <= 3 ? 3 : (kx == 4 ? 4 : 6), ow, oh, ow_band, rows_lo, rows_hi, cap, lin);
    __syncthreads();
    float* const gimg = direct ? a.image_f32_out + (size_t)b * npix : nullptr;
#define B200AUG_BAND(KK) area_band<KK, false>(tm, cap, ow, oh, ow_band, warp, lane, cr, cl, gimg)
    if (lin) area_band<2, true>(tm, cap, ow, oh, ow_band, warp, lane, cr, cl, gimg);
    else if (kx <= 3) B200AUG_BAND(3);
    else if (kx == 4) B200AUG_BAND(4);
    else B200AUG_BAND(6);
#undef B200AUG_BAND
    const int tw = ow - ow_band;
    if (tw > 0 && rs == RS_AREA) {
      // the tail column(s) of a general-factor INTER_AREA crop: scalar_out_px() specialised (plan fields in registers, crop
      // source only), one thread per pixel
      const uint8_t* const src = P.src;
      const int pitch = P.pitch, sw = P.sw, sh = P.sh, x0 = P.x0, y0 = P.y0;
      for (int p = tid; p < (rows_hi - rows_lo) * tw; p += NTHREADS) {
        const int dy = rows_lo + p / tw, dx = ow_band + p % tw;
        const int xs = T.start[dx], xnf = T.n[dx], xn = xnf & 0xffff;
        const bool xhf = xnf & (1 << 30), xhl = xnf & (1u << 31);
        const float xaf = T.a[dx], xam = T.b[dx], xal = T.c[dx];
        const int ys = T.start[ow + dy], ynf = T.n[ow + dy], yn = ynf & 0xffff;
        const bool yhf = ynf & (1 << 30), yhl = ynf & (1u << 31);
        const float yaf = T.a[ow + dy], yam = T.b[ow + dy], yal = T.c[ow + dy];
        float acc = 0.f;
        // The first TAIL_ROWS canvas rows of the pixel: every tap is loaded unconditionally from a clamped (always valid)
        // address and masked afterwards, so the TAIL_ROWS x KMAX loads are independent and in flight together (a guarded
        // load per tap compiles to a branch per tap, i.e. one L2 round trip after the other: 5 us of the step).
        constexpr int TAIL_ROWS = 4;
        uint32_t px[TAIL_ROWS][KMAX];
#pragma unroll
        for (int k = 0; k < TAIL_ROWS; ++k) {
          const int sy = y0 + ys + k;
          const bool row_in = k < yn && (unsigned)sy < (unsigned)sh;
          const uint8_t* row = src + (ptrdiff_t)min(max(sy, 0), sh - 1) * pitch;
#pragma unroll
          for (int t = 0; t < KMAX; ++t) {
            const int sx = x0 + xs + t;
            const uint32_t v = row[min(max(sx, 0), sw - 1)];  // (coherent load: the source may be the scratch canvas)
            px[k][t] = (row_in && t < xn && (unsigned)sx < (unsigned)sw) ? v : 0u;
          }
        }
#pragma unroll
        for (int k = 0; k < TAIL_ROWS; ++k) {
          if (k < yn) {
            float h = 0.f;
#pragma unroll
            for (int t = 0; t < KMAX; ++t)
              if (t < xn) h = __fadd_rn(h, __fmul_rn((float)px[k][t], area_alpha(t, xn, xhf, xhl, xaf, xam, xal)));
            const float beta = area_alpha(k, yn, yhf, yhl, yaf, yam, yal);
            acc = (k == 0) ? __fmul_rn(beta, h) : __fadd_rn(acc, __fmul_rn(beta, h));
          }
        }
        for (int k = TAIL_ROWS; k < yn; ++k) {  // (scale factors above 3: the remaining rows one tap after the other)
          const int sy = y0 + ys + k;
          const bool row_in = (unsigned)sy < (unsigned)sh;
          const uint8_t* row = src + (ptrdiff_t)sy * pitch + (x0 + xs);
          float h = 0.f;
          for (int t = 0; t < xn; ++t) {
            const bool in = row_in && (unsigned)(x0 + xs + t) < (unsigned)sw;
            const float px1 = in ? (float)row[t] : 0.f;
            h = __fadd_rn(h, __fmul_rn(px1, area_alpha(t, xn, xhf, xhl, xaf, xam, xal)));
          }
          const float beta = area_alpha(k, yn, yhf, yhl, yaf, yam, yal);
          acc = __fadd_rn(acc, __fmul_rn(beta, h));
        }
        const uint8_t q = sat_u8_rint(acc);
        if (direct) gimg[tm.o + dy * tm.sa + dx * tm.sb] = lut[q];
        else tile[tm.o + dy * tm.sa + dx * tm.sb] = q;
      }
    } else {
      for (int p = tid; p < (rows_hi - rows_lo) * tw; p += NTHREADS) {
        const int dy = rows_lo + p / tw, dx = ow_band + p % tw;
        const uint8_t q = scalar_out_px(P, T, ow, dx, dy);
        if (direct) gimg[tm.o + dy * tm.sa + dx * tm.sb] = lut[q];
        else tile[tm.o + dy * tm.sa + dx * tm.sb] = q;
      }
    }
  } else if (P.status == B200AUG_S_OK && rs == RS_LINEAR && P.src_mode != SRC_WARP) {
    // cv2's 2-tap kernels (INTER_LINEAR up-scaling, or INTER_AREA with an up-scaling axis) from a crop: scalar_out_px()
    // specialised -- plan fields in registers, no resampler / source dispatch per tap; consecutive threads = consecutive
    // output columns, so the taps of a warp fall into a few cache lines
    const uint8_t* const src = P.src;
    const int pitch = P.pitch, sw = P.sw, sh = P.sh, x0 = P.x0, y0 = P.y0;
    for (int p = rows_lo * ow + tid; p < rows_hi * ow; p += NTHREADS) {
      const int dy = p / ow, dx = p - dy * ow;
      const int sxa = x0 + T.start[dx], sxb = x0 + T.n[dx], sya = y0 + T.start[ow + dy], syb = y0 + T.n[ow + dy];
      const int a0 = __float_as_int(T.a[dx]), a1 = __float_as_int(T.b[dx]);
      const int b0 = __float_as_int(T.a[ow + dy]), b1 = __float_as_int(T.b[ow + dy]);
      const bool xa = (unsigned)sxa < (unsigned)sw, xb = (unsigned)sxb < (unsigned)sw;
      const bool ya = (unsigned)sya < (unsigned)sh, yb = (unsigned)syb < (unsigned)sh;
      const uint8_t* ra = src + (ptrdiff_t)sya * pitch;
      const uint8_t* rb = src + (ptrdiff_t)syb * pitch;
      // (coherent loads: the source may be the scratch canvas written earlier in this kernel)
      const int p00 = (ya && xa) ? ra[sxa] : 0, p01 = (ya && xb) ? ra[sxb] : 0;
      const int p10 = (yb && xa) ? rb[sxa] : 0, p11 = (yb && xb) ? rb[sxb] : 0;
      const int h0 = p00 * a0 + p01 * a1, h1 = p10 * a0 + p11 * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      tile[tm.o + dy * tm.sa + dx * tm.sb] = (uint8_t)min(max(v, 0), 255);
    }
  } else if (P.status == B200AUG_S_OK) {
    // per-pixel path: integer-factor area, plain copy, very wide taps / canvases, rotated samples without a scratch canvas
    for (int p = rows_lo * ow + tid; p < rows_hi * ow; p += NTHREADS) {
      const int dy = p / ow, dx = p - dy * ow;
      tile[tm.o + dy * tm.sa + dx * tm.sb] = scalar_out_px(P, T, ow, dx, dy);
    }
  } else {
    for (int p = tid; p < npix; p += NTHREADS) tile[p] = 0;
  }
  if (direct) {  // the crop is already in global memory; only the labels are left (both CTAs of the cluster take this exit)
    trace_mark<TRACE>(a, 3);
    trace_mark<TRACE>(a, 9);
    trace_mark<TRACE>(a, 4);
    return;
  }
  // ---- cluster exchange: every CTA resampled a band of rows into its own tile; now each sends its band to the others.
  // Without a 90-degree rotation the band is a contiguous byte range of the tile: its 16-byte aligned interior goes as
  // one bulk copy through distributed shared memory (completing on the receiver's mbarrier), the ragged ends as byte
  // stores.  Rotated-by-90 samples (1 %) send pixel by pixel.
  if (cl > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (cl > 1 && P.status == B200AUG_S_OK) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // tile writes (generic proxy) before the bulk copy reads them
    __syncthreads();
    const uint32_t tile32 = smem_u32(tile), xbar32 = smem_u32(xbar);
    if (P.rot_dir == 0) {
      const int lo = rows_lo * ow, hi = rows_hi * ow, alo = min((lo + 15) & ~15, hi), ahi = max(hi & ~15, alo);
      for (int q = 0; q < cl; ++q) {
        if (q == cr) continue;
        const uint32_t rt = map_to_rank(tile32, q);
        if (tid == 0 && ahi > alo)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rt + alo),
                       "r"(tile32 + alo), "r"(ahi - alo), "r"(map_to_rank(xbar32, q))
                       : "memory");
        for (int i = lo + tid; i < alo; i += NTHREADS) st_cluster_u8(rt + i, tile[i]);
        for (int i = ahi + tid; i < hi; i += NTHREADS) st_cluster_u8(rt + i, tile[i]);
      }
    } else {
      for (int q = 0; q < cl; ++q) {
        if (q == cr) continue;
        const uint32_t rt = map_to_rank(tile32, q);
        for (int p = rows_lo * ow + tid; p < rows_hi * ow; p += NTHREADS) {
          const int dy = p / ow, dx = p - dy * ow, idx = tm.o + dy * tm.sa + dx * tm.sb;
          st_cluster_u8(rt + idx, tile[idx]);
        }
      }
    }
    if (rx_bytes) mbar_wait(xbar32, 0);  // the other CTAs' bands have landed in this tile
  }
  cluster_sync(cl);  // every CTA's tile now holds the whole crop; no distributed-shared-memory access after this point
  trace_mark<TRACE>(a, 3);
  trace_mark<TRACE>(a, 9);

  // ---- uint8 output (geometric stages only) --------------------------------------------------------------
  if (!(a.flags & B200AUG_F_NORMALIZE)) {
    uint8_t* out = a.image_u8_out + (size_t)b * npix;
    for (int p = ((cr * npix) >> cs) + tid; p < (((cr + 1) * npix) >> cs); p += NTHREADS) out[p] = tile[p];
    trace_mark<TRACE>(a, 4);
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
  const bool whiten = a.flags & B200AUG_F_WHITEN;
  const bool photo = a.flags & B200AUG_F_PHOTOMETRIC;
  const bool clip = photo && a.photo.clip;
  // without blur and noise the rest of the chain is a point function too: fold clip and whiten into the LUT
  const bool folded = (P.blur_pos < 0) && !P.any_noise;
  if (folded) {
    if (clip) xv = fminf(fmaxf(xv, 0.f), 1.f);
    if (whiten) xv = __fsub_rn(xv, 0.5f);
  }
  lut[tid] = xv;
  __syncthreads();

  if (eq_after_blur) {
    // rare: equalize sits after the blur -> histogram of the blurred (+ intermediate ops) values
    for (int p = tid; p < npix; p += NTHREADS) {
      float x = blurred_value(tile, lut, ow, oh, p);
      x = apply_point_ops(P, x, n_pre + 1, P.eq_pos, eq_lut);
      atomicAdd(&binhist[eq_bin(x)], 1u);
    }
    __syncthreads();
    build_eq_lut();
    __syncthreads();
  }

  trace_mark<TRACE>(a, 10);
  // ---- output pass ----------------------------------------------------------------------------------------
  float* out = a.image_f32_out + (size_t)b * npix;
  const int Q = (npix + 3) >> 2;
  const int g_lo = (cr * Q) >> cs, g_hi = ((cr + 1) * Q) >> cs;  // this CTA's share of the output
  if (folded) {
    for (int g = g_lo + tid; g < g_hi; g += NTHREADS) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = g + i * Q;
        if (p < npix) out[p] = lut[tile[p]];
      }
    }
    trace_mark<TRACE>(a, 4);
    return;
  }
  const uint64_t sid = a.photo.sample_offset + (uint64_t)b;
  const uint2 key = make_uint2((uint32_t)a.photo.seed, (uint32_t)(a.photo.seed >> 32));
  const bool blur = P.blur_pos >= 0;
  // Blur: separable, strip by strip.  This CTA's pixels are the four ranges [g_lo, g_hi) + i Q; for each range the
  // horizontal pass of its rows (+ 2 halo rows either side, reflect border) goes into the warps' idle row buffers as float32,
  // then the vertical pass and the ops behind the blur produce the pre-noise value, parked in the output buffer itself.
  float* const hb = reinterpret_cast<float*>(smem + L.off_rowbuf);
  const int hb_rows = (int)(((size_t)NWARPS * (cap + ROWBUF_SLACK)) / (sizeof(float) * (size_t)ow));
  const bool blur_strips = blur && hb_rows >= 6;
  if (blur_strips) {
    // float32(exp(-t^2/(2 sigma^2))) / float32 sum, identical to oracle/photometric.py:gaussian_kernel1d
    const float g0 = 0x1.ebd752p-4f, g1 = 0x1.defcdep-3f, g2 = 0x1.2b1778p-2f;
    const int strip = hb_rows - 4;
    for (int i = 0; i < 4; ++i) {
      const int p_lo = g_lo + i * Q, p_hi = min(g_hi + i * Q, npix);
      if (p_lo >= p_hi) continue;
      const int y_last = (p_hi - 1) / ow;
      for (int ry0 = p_lo / ow; ry0 <= y_last; ry0 += strip) {
        const int ry1 = min(ry0 + strip - 1, y_last), nr = ry1 - ry0 + 5;
        __syncthreads();  // the previous strip's vertical pass is done with hb
        // horizontal pass over the strip's rows (incl. halo), one value per thread and step; (row, column) advance without
        // a division, interior columns need no reflection
        {
          const int step_r = NTHREADS / ow, step_x = NTHREADS - step_r * ow;
          int rr = tid / ow, x = tid - rr * ow;
          for (int v = tid; v < nr * ow; v += NTHREADS) {
            const uint8_t* row = tile + reflect_idx(ry0 - 2 + rr, oh) * ow;
            const bool inner = x >= 2 && x + 2 < ow;
            const int xm2 = inner ? x - 2 : reflect_idx(x - 2, ow), xm1 = inner ? x - 1 : reflect_idx(x - 1, ow);
            const int xp1 = inner ? x + 1 : reflect_idx(x + 1, ow), xp2 = inner ? x + 2 : reflect_idx(x + 2, ow);
            float t = __fmul_rn(g0, lut[row[xm2]]);
            t = __fadd_rn(t, __fmul_rn(g1, lut[row[xm1]]));
            t = __fadd_rn(t, __fmul_rn(g2, lut[row[x]]));
            t = __fadd_rn(t, __fmul_rn(g1, lut[row[xp1]]));
            t = __fadd_rn(t, __fmul_rn(g0, lut[row[xp2]]));
            hb[v] = t;
            x += step_x;
            rr += step_r;
            if (x >= ow) { x -= ow; ++rr; }
          }
        }
        __syncthreads();
        // vertical pass + the ops behind the blur; only the pixels of [p_lo, p_hi) are this CTA's
        const bool post_ops = P.blur_pos + 1 < P.n_ops;
        const int q_lo = max(p_lo, ry0 * ow), q_hi = min(p_hi, (ry1 + 1) * ow);
        for (int p = q_lo + tid; p < q_hi; p += NTHREADS) {
          const float* c = hb + (p - ry0 * ow);  // horizontal-pass value of the pixel two rows up
          float acc = __fmul_rn(g0, c[0]);
          acc = __fadd_rn(acc, __fmul_rn(g1, c[ow]));
          acc = __fadd_rn(acc, __fmul_rn(g2, c[2 * ow]));
          acc = __fadd_rn(acc, __fmul_rn(g1, c[3 * ow]));
          acc = __fadd_rn(acc, __fmul_rn(g0, c[4 * ow]));
          out[p] = post_ops ? apply_point_ops(P, acc, P.blur_pos + 1, P.n_ops, eq_lut) : acc;
        }
      }
    }
    __syncthreads();  // the parked values are read back by other threads of this CTA
  }
  for (int g = g_lo + tid; g < g_hi; g += NTHREADS) {
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = g + i * Q;
      float v = 0.f;
      if (p < npix) {
        if (blur_strips) v = out[p];
        else if (blur) v = apply_point_ops(P, blurred_value(tile, lut, ow, oh, p), P.blur_pos + 1, P.n_ops, eq_lut);
        else v = lut[tile[p]];
      }
      x[i] = v;
    }
    if (P.any_noise) {
#pragma unroll
      for (int s = 0; s < B200AUG_NUM_NOISE; ++s) {
        if (!P.noise_on[s]) continue;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)s, (uint32_t)sid, 0x6E6F6973u ^ (uint32_t)(sid >> 32)), key);
        float z[4];
        box_muller(r.x, r.y, z[0], z[1]);
        box_muller(r.z, r.w, z[2], z[3]);
        const float sd = a.photo.noise_std[s];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = __fadd_rn(x[i], __fmul_rn(sd, z[i]));
        if (a.photo.noise_clip[s])  // RandomGaussianNoiseWithClipping
#pragma unroll
          for (int i = 0; i < 4; ++i) x[i] = fminf(fmaxf(x[i], 0.f), 1.f);
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
  trace_mark<TRACE>(a, 4);
}


// ------------------------------------------------------------------------------------------------ stand-alone photometric

// KorniaImageDistortions on float32 images that are already on the device (batch/intensity.py:30-40 called on its own,
// pipelines.py:508-527): the same op semantics as the fused kernel, but on arbitrary float input, so nothing collapses
// into a uint8 LUT.  One CTA per image; the image ping-pongs between `out` and `tmp` (both stay in L2).  Point ops are
// deferred ("pending") and evaluated on the fly by whichever pass needs their result next: the equalize histogram, the
// materialisation in front of the blur, or the final noise / clip pass.
__device__ __forceinline__ float photo_f32_value(const Plan& P, const float* cur, int p, int from, int to, const float* eq_lut) {
  return apply_point_ops(P, cur[p], from, to, eq_lut);
}

__global__ void __launch_bounds__(NTHREADS) photometric_f32_kernel(const float* in, float* out, float* tmp,
                                                                   int w, int h, const __grid_constant__ B200AugPhotoParams pp,
                                                                   float bias) {
  __shared__ Plan P;
  __shared__ float eq_lut[256];
  __shared__ unsigned binhist[256];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, npix = w * h;
  const float* src = in + (size_t)b * npix;
  float* A = out + (size_t)b * npix;
  float* B = tmp ? tmp + (size_t)b * npix : nullptr;
  if (tid == 0) plan_photo(pp, true, b, P);
  __syncthreads();
  const float* cur = src;
  int from = 0;  // ops [from, k) are pending on `cur`
  for (int k = 0; k < P.n_ops; ++k) {
    const int op = P.ops[k];
    if (op == B200AUG_OP_EQUALIZE) {
      binhist[tid] = 0;
      __syncthreads();
      for (int p = tid; p < npix; p += NTHREADS) {
        const float x = photo_f32_value(P, cur, p, from, k, eq_lut);
        const float im = __fmul_rn(x, 255.f);
        if (im >= 0.f && im <= 255.f) atomicAdd(&binhist[eq_bin(x)], 1u);  // torch.histc ignores out-of-range values
      }
      __syncthreads();
      if (warp == 0) {
        unsigned loc[8], run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { loc[i] = binhist[lane * 8 + i]; run += loc[i]; }
        unsigned incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        int my_last = -1;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (loc[i]) my_last = lane * 8 + i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_last = max(my_last, __shfl_xor_sync(0xffffffffu, my_last, o));
        const unsigned last_nz = (my_last >= 0) ? binhist[my_last] : 0u;
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned step = (total - last_nz) / 255u;
        if (lane == 0) P.eq_step0 = (step == 0);
        if (step) {
          unsigned c = incl - run;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int bin = lane * 8 + i;
            eq_lut[bin] = (bin == 0) ? 0.f : (float)min((c + step / 2u) / step, 255u);
            c += loc[i];
          }
        }
      }
      __syncthreads();  // from here on op k is a point op like the others
    } else if (op == B200AUG_OP_BLUR) {
      // materialise the pending ops, then the separable 5x5 (horizontal pass inside the vertical one, same order of
      // float operations as two full passes)
      float* M = (cur == A) ? B : A;
      for (int p = tid; p < npix; p += NTHREADS) M[p] = photo_f32_value(P, cur, p, from, k, eq_lut);
      __syncthreads();
      float* D = (M == A) ? B : A;
      const float g[5] = {0x1.ebd752p-4f, 0x1.defcdep-3f, 0x1.2b1778p-2f, 0x1.defcdep-3f, 0x1.ebd752p-4f};
      for (int p = tid; p < npix; p += NTHREADS) {
        const int y = p / w, x = p - y * w;
        int xs[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) xs[i] = reflect_idx(x + i - 2, w);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const float* row = M + reflect_idx(y + j - 2, h) * w;
          float t = __fmul_rn(g[0], row[xs[0]]);
#pragma unroll
          for (int i = 1; i < 5; ++i) t = __fadd_rn(t, __fmul_rn(g[i], row[xs[i]]));
          acc = (j == 0) ? __fmul_rn(g[0], t) : __fadd_rn(acc, __fmul_rn(g[j], t));
        }
        D[p] = acc;
      }
      __syncthreads();
      cur = D;
      from = k + 1;
    }
  }
  // final pass: pending point ops, the noise stages, clip, bias (whiten), in the quad layout of the Philox stream
  const int Q = (npix + 3) >> 2;
  const uint64_t sid = pp.sample_offset + (uint64_t)b;
  const uint2 key = make_uint2((uint32_t)pp.seed, (uint32_t)(pp.seed >> 32));
  for (int g = tid; g < Q; g += NTHREADS) {
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = g + i * Q;
      x[i] = (p < npix) ? photo_f32_value(P, cur, p, from, P.n_ops, eq_lut) : 0.f;
    }
    if (P.any_noise) {
#pragma unroll
      for (int s = 0; s < B200AUG_NUM_NOISE; ++s) {
        if (!P.noise_on[s]) continue;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)s, (uint32_t)sid, 0x6E6F6973u ^ (uint32_t)(sid >> 32)), key);
        float z[4];
        box_muller(r.x, r.y, z[0], z[1]);
        box_muller(r.z, r.w, z[2], z[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = __fadd_rn(x[i], __fmul_rn(pp.noise_std[s], z[i]));
        if (pp.noise_clip[s])  // RandomGaussianNoiseWithClipping (torch.clip keeps a NaN)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (x[i] == x[i]) x[i] = fminf(fmaxf(x[i], 0.f), 1.f);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = g + i * Q;
      if (p < npix) {
        float v = x[i];
        if (pp.clip && v == v) v = fminf(fmaxf(v, 0.f), 1.f);  // (torch.clip keeps a NaN; arbitrary float input can carry one)
        A[p] = __fadd_rn(v, bias);
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
    if (F.category == B200AUG_CAT_BACKTRANSFORM) {  // BT @ tr^-1, affinetrafo.py:137-147
      if (threadIdx.x == 0) {
        const Aff bt = aff_compose(Aff{in[0], in[1], in[2], in[3], in[4], in[5]}, aff_inv(D.m));
        out[0] = bt.a00; out[1] = bt.a01; out[2] = bt.a02; out[3] = bt.a10; out[4] = bt.a11; out[5] = bt.a12;
      }
      continue;
    }
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

// ------------------------------------------------------------------------------------------------ rotation labels (eval side)

// torchquaternion.mult (neuralnets/torchquaternion.py:23-48), xyzw: the 4x4 matrix form of u times the (w,i,j,k) vector of v
__device__ __forceinline__ void quat_mult(const float* u, const float* v, float* o) {
  const float ui = u[0], uj = u[1], uk = u[2], uw = u[3], vi = v[0], vj = v[1], vk = v[2], vw = v[3];
  const float w = add(add(add(mul(uw, vw), mul(-ui, vi)), mul(-uj, vj)), mul(-uk, vk));
  const float i = add(add(add(mul(ui, vw), mul(uw, vi)), mul(-uk, vj)), mul(uj, vk));
  const float j = add(add(add(mul(uj, vw), mul(uk, vi)), mul(uw, vj)), mul(-ui, vk));
  const float k = add(add(add(mul(uk, vw), mul(-uj, vi)), mul(ui, vj)), mul(uw, vk));
  o[0] = i; o[1] = j; o[2] = k; o[3] = w;
}

// torchquaternion.from_matrix (torchquaternion.py:94-168): four candidate solutions, the best conditioned one is picked
// (first maximum of the clamped square-root arguments), then positivereal() (q * sign(w); sign(0) = 0 as in torch)
__device__ void quat_from_matrix(const float (&m)[3][3], float* q) {
  float arg[4];
  arg[0] = add(add(add(-m[0][0], -m[1][1]), m[2][2]), 1.f);   // k
  arg[1] = add(add(add(-m[0][0], m[1][1]), -m[2][2]), 1.f);   // j
  arg[2] = add(add(add(m[0][0], -m[1][1]), -m[2][2]), 1.f);   // i
  arg[3] = add(add(add(m[0][0], m[1][1]), m[2][2]), 1.f);     // w
  int pick = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    arg[t] = fmaxf(arg[t], 1.0e-6f);
    if (arg[t] > arg[pick]) pick = t;
  }
  const float d = mul(sqrtf(arg[pick]), 0.5f);
  auto val = [&](float a, float b) { return fdiv(mul(0.25f, add(a, b)), d); };
  float qi, qj, qk, qw;
  switch (pick) {
    case 0: qw = val(m[1][0], -m[0][1]); qi = val(m[2][0], m[0][2]); qj = val(m[1][2], m[2][1]); qk = d; break;
    case 1: qw = val(m[0][2], -m[2][0]); qi = val(m[1][0], m[0][1]); qk = val(m[1][2], m[2][1]); qj = d; break;
    case 2: qw = val(m[2][1], -m[1][2]); qj = val(m[1][0], m[0][1]); qk = val(m[0][2], m[2][0]); qi = d; break;
    default: qi = val(m[2][1], -m[1][2]); qj = val(m[0][2], -m[2][0]); qk = val(m[1][0], -m[0][1]); qw = d; break;
  }
  const float sg = (qw > 0.f) ? 1.f : ((qw < 0.f) ? -1.f : 0.f);
  q[0] = mul(qi, sg); q[1] = mul(qj, sg); q[2] = mul(qk, sg); q[3] = mul(qw, sg);
}

// PerspectiveCorrector._make_look_at_matrix (eval.py:531-544), columns x, y, z; note y / |x| as in the reference (:542)
__device__ void look_at_matrix(float px, float py, float pz, float (&m)[3][3]) {
  const float n = sqrtf(add(add(mul(px, px), mul(py, py)), mul(pz, pz)));
  const float zx = fdiv(px, n), zy = fdiv(py, n), zz = fdiv(pz, n);
  // cross((0,1,0), z) = (zz, 0, -zx)
  float xx = zz, xy = 0.f, xz = -zx;
  const float nx = sqrtf(add(add(mul(xx, xx), mul(xy, xy)), mul(xz, xz)));
  xx = fdiv(xx, nx); xy = fdiv(xy, nx); xz = fdiv(xz, nx);
  float yx = sub(mul(zy, xz), mul(zz, xy)), yy = sub(mul(zz, xx), mul(zx, xz)), yz = sub(mul(zx, xy), mul(zy, xx));
  const float nx2 = sqrtf(add(add(mul(xx, xx), mul(xy, xy)), mul(xz, xz)));
  yx = fdiv(yx, nx2); yy = fdiv(yy, nx2); yz = fdiv(yz, nx2);
  m[0][0] = xx; m[1][0] = xy; m[2][0] = xz;
  m[0][1] = yx; m[1][1] = yy; m[2][1] = yz;
  m[0][2] = zx; m[1][2] = zy; m[2][2] = zz;
}

// PerspectiveCorrector.corrected_rotation (eval.py:491-529): one thread per sample
__global__ void corrected_rotation_kernel(const float* __restrict__ half_sizes, int64_t size_stride, float div_x, float div_y, float f,
                                          const float* __restrict__ coord, int64_t coord_stride, const float* __restrict__ pose,
                                          float* __restrict__ out, float* __restrict__ look_at_out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xn, yn, zn = f;
  if (half_sizes) {
    const float hx = half_sizes[i * size_stride], hy = half_sizes[i * size_stride + 1];
    xn = fdiv(sub(coord[i * coord_stride], hx), div_x);
    yn = fdiv(sub(coord[i * coord_stride + 1], hy), div_y);
  } else {  // coord rows are the look-at positions themselves
    xn = coord[i * coord_stride];
    yn = coord[i * coord_stride + 1];
    zn = coord[i * coord_stride + 2];
  }
  float m[3][3];
  look_at_matrix(xn, yn, zn, m);
  if (look_at_out)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) look_at_out[(size_t)i * 9 + r * 3 + c] = m[r][c];
  if (out) {
    float qm[4], qo[4];
    quat_from_matrix(m, qm);
    const float qp[4] = {pose[4 * (size_t)i], pose[4 * (size_t)i + 1], pose[4 * (size_t)i + 2], pose[4 * (size_t)i + 3]};
    quat_mult(qm, qp, qo);
#pragma unroll
    for (int k = 0; k < 4; ++k) out[4 * (size_t)i + k] = qo[k];
  }
}

// torchquaternion.tomatrix (torchquaternion.py:70-91; the 6D / matrix target of losses.py:53-58) and from_matrix
__global__ void quat_matrix_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int to_matrix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (to_matrix) {
    const float qi = in[4 * (size_t)i], qj = in[4 * (size_t)i + 1], qk = in[4 * (size_t)i + 2], qw = in[4 * (size_t)i + 3];
    float* o = out + 9 * (size_t)i;
    o[0] = sub(1.f, mul(2.f, add(mul(qj, qj), mul(qk, qk))));
    o[3] = mul(2.f, add(mul(qi, qj), mul(qk, qw)));
    o[6] = mul(2.f, sub(mul(qi, qk), mul(qj, qw)));
    o[1] = mul(2.f, sub(mul(qi, qj), mul(qk, qw)));
    o[4] = sub(1.f, mul(2.f, add(mul(qi, qi), mul(qk, qk))));
    o[7] = mul(2.f, add(mul(qj, qk), mul(qi, qw)));
    o[2] = mul(2.f, add(mul(qi, qk), mul(qj, qw)));
    o[5] = mul(2.f, sub(mul(qj, qk), mul(qi, qw)));
    o[8] = sub(1.f, mul(2.f, add(mul(qi, qi), mul(qj, qj))));
  } else {
    float m[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) m[r][c] = in[9 * (size_t)i + 3 * r + c];
    quat_from_matrix(m, out + 4 * (size_t)i);
  }
}

// PutRoiFromLandmarks(extend_to_forehead=True), batch/misc.py:14-26: the roi is the xy bounding box of ALL vertices of the
// posed deformable face model -- PosedDeformableHead (modelcomponents.py:85-94): local = vertices + sum_k base[k] * shape[k]
// (ScaledBfmModule.forward, bfm.py:91-95), rotated by the pose quaternion, scaled by coord[2], shifted by coord[:2]
// (rigid_transformation_25d, modelcomponents.py:38-56).  One CTA takes HR_SPB samples and walks the vertex list once for all
// of them, so a deformation base entry is read once per HR_SPB samples; without shape parameters (what the reference does for
// every dataset: it looks for a key "shapeparams" that no sample has) the inner loop disappears and the whole thing is a
// min / max over 38 k rotated points.  `xy_offset` = the half-pixel shift offset_points_by_half_pixel would have applied to
// coord in front of this transform (normalization.py:83-90), 0 when the caller already did.
constexpr int HR_SPB = 8;
constexpr int HR_MAXK = 64;

__global__ void __launch_bounds__(NTHREADS) head_roi_kernel(const float* __restrict__ vertices, const float* __restrict__ deform_base,
                                                            int V, int K, const float* __restrict__ shapeparams,
                                                            const float* __restrict__ coord, const float* __restrict__ quat,
                                                            float xy_offset, float* __restrict__ roi_out, int B) {
  __shared__ float sp[HR_MAXK][HR_SPB];
  __shared__ float pose[HR_SPB][12];  // rotation matrix (row-major), scale, tx, ty
  __shared__ float red[NWARPS][HR_SPB][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s0 = blockIdx.x * HR_SPB, ns = min(HR_SPB, B - s0);
  const bool deform = K > 0 && shapeparams != nullptr && deform_base != nullptr;
  for (int i = tid; i < HR_MAXK * HR_SPB; i += NTHREADS) {
    const int k = i / HR_SPB, s = i % HR_SPB;
    sp[k][s] = (deform && k < K && s < ns) ? shapeparams[(size_t)(s0 + s) * K + k] : 0.f;
  }
  if (tid < HR_SPB) {
    const int b = s0 + min(tid, ns - 1);
    const float qi = quat[4 * (size_t)b], qj = quat[4 * (size_t)b + 1], qk = quat[4 * (size_t)b + 2], qw = quat[4 * (size_t)b + 3];
    float* o = pose[tid];
    o[0] = sub(1.f, mul(2.f, add(mul(qj, qj), mul(qk, qk))));
    o[1] = mul(2.f, sub(mul(qi, qj), mul(qk, qw)));
    o[2] = mul(2.f, add(mul(qi, qk), mul(qj, qw)));
    o[3] = mul(2.f, add(mul(qi, qj), mul(qk, qw)));
    o[4] = sub(1.f, mul(2.f, add(mul(qi, qi), mul(qk, qk))));
    o[5] = mul(2.f, sub(mul(qj, qk), mul(qi, qw)));
    o[9] = coord[3 * (size_t)b + 2];
    o[10] = add(coord[3 * (size_t)b], xy_offset);
    o[11] = add(coord[3 * (size_t)b + 1], xy_offset);
  }
  __syncthreads();
  float lo_x[HR_SPB], lo_y[HR_SPB], hi_x[HR_SPB], hi_y[HR_SPB];
#pragma unroll
  for (int s = 0; s < HR_SPB; ++s) {
    lo_x[s] = lo_y[s] = INFINITY;
    hi_x[s] = hi_y[s] = -INFINITY;
  }
  for (int v = tid; v < V; v += NTHREADS) {
    const float vx = vertices[3 * (size_t)v], vy = vertices[3 * (size_t)v + 1], vz = vertices[3 * (size_t)v + 2];
    float ax[HR_SPB], ay[HR_SPB], az[HR_SPB];
#pragma unroll
    for (int s = 0; s < HR_SPB; ++s) ax[s] = ay[s] = az[s] = 0.f;
    if (deform) {
      for (int k = 0; k < K; ++k) {
        const float* bp = deform_base + ((size_t)k * V + v) * 3;
        const float bx = bp[0], by = bp[1], bz = bp[2];
#pragma unroll
        for (int s = 0; s < HR_SPB; ++s) {
          const float w = sp[k][s];
          ax[s] = __fmaf_rn(bx, w, ax[s]);
          ay[s] = __fmaf_rn(by, w, ay[s]);
          az[s] = __fmaf_rn(bz, w, az[s]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < HR_SPB; ++s) {
      const float x = add(ax[s], vx), y = add(ay[s], vy), z = add(az[s], vz);
      const float* o = pose[s];
      const float rx = __fmaf_rn(o[2], z, __fmaf_rn(o[1], y, mul(o[0], x)));
      const float ry = __fmaf_rn(o[5], z, __fmaf_rn(o[4], y, mul(o[3], x)));
      const float px = add(mul(rx, o[9]), o[10]), py = add(mul(ry, o[9]), o[11]);
      lo_x[s] = fminf(lo_x[s], px); hi_x[s] = fmaxf(hi_x[s], px);
      lo_y[s] = fminf(lo_y[s], py); hi_y[s] = fmaxf(hi_y[s], py);
    }
  }
#pragma unroll
  for (int s = 0; s < HR_SPB; ++s) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo_x[s] = fminf(lo_x[s], __shfl_xor_sync(0xffffffffu, lo_x[s], d));
      lo_y[s] = fminf(lo_y[s], __shfl_xor_sync(0xffffffffu, lo_y[s], d));
      hi_x[s] = fmaxf(hi_x[s], __shfl_xor_sync(0xffffffffu, hi_x[s], d));
      hi_y[s] = fmaxf(hi_y[s], __shfl_xor_sync(0xffffffffu, hi_y[s], d));
    }
    if (lane == 0) {
      red[warp][s][0] = lo_x[s]; red[warp][s][1] = lo_y[s]; red[warp][s][2] = hi_x[s]; red[warp][s][3] = hi_y[s];
    }
  }
  __syncthreads();
  if (tid < HR_SPB * 4 && tid / 4 < ns) {
    const int s = tid / 4, c = tid % 4;
    float r = red[0][s][c];
    for (int w = 1; w < NWARPS; ++w) r = (c < 2) ? fminf(r, red[w][s][c]) : fmaxf(r, red[w][s][c]);
    roi_out[4 * (size_t)(s0 + s) + c] = r;
  }
}

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

extern "C" int b200aug_fused_occupancy(int out_w, int out_h, int rowbuf_capacity, int cluster_size, int* ctas_per_sm, int* active_clusters) {
  if (out_w <= 0 || out_h <= 0 || !ctas_per_sm || !active_clusters) return B200AUG_E_INVALID_ARG;
  int cap = rowbuf_capacity > 0 ? rowbuf_capacity : DEFAULT_ROWBUF;
  cap = (cap + 15) & ~15;
  const size_t smem = smem_layout(out_w, out_h, cap).total;
  if (smem > 227 * 1024) return B200AUG_E_SMEM;
  const int cl = cluster_size > 0 ? cluster_size : DEFAULT_CLUSTER;
  cudaError_t e = cudaFuncSetAttribute(fused_augment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, fused_augment_kernel<false>, NTHREADS, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 * 8 * cl);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(active_clusters, fused_augment_kernel<false>, &cfg);
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}

extern "C" int64_t b200aug_plan_stride(int out_w, int out_h) {
  if (out_w <= 0 || out_h <= 0) return 0;
  return (int64_t)(plan_bytes() + plan_tab_bytes(out_w, out_h));
}

extern "C" int64_t b200aug_plan_buffer_bytes(int batch, int out_w, int out_h) {
  if (batch < 0 || out_w <= 0 || out_h <= 0) return 0;
  return (int64_t)batch * b200aug_plan_stride(out_w, out_h) + (int64_t)plan_tail_bytes(batch);
}

extern "C" int64_t b200aug_workspace_stride(int max_side) {
  if (max_side <= 0) return 0;
  const int64_t spitch = (max_side + 16 + 15) & ~15;
  return ((spitch * (max_side + 1)) + 255) & ~int64_t(255);
}

extern "C" int b200aug_remap_table(int upfilter, int16_t* out) {
  // cv::initInterTab2D(fixpt) (oracle/cv2_model.py:remap_table): plain host code
  if (!out || (upfilter != B200AUG_UP_CUBIC && upfilter != B200AUG_UP_LANCZOS)) return B200AUG_E_INVALID_ARG;
  const int k = (upfilter == B200AUG_UP_CUBIC) ? 4 : 8, k2 = k / 2;
  float t1[32][8];
  for (int i = 0; i < 32; ++i) {
    volatile float x = (float)i * (1.0f / 32);
    if (k == 4) cubic_coeffs(x, t1[i]);
    else lanczos4_coeffs(x, t1[i]);
  }
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      int16_t* it = out + (size_t)(i * 32 + j) * k * k;
      int isum = 0;
      for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
          volatile float v = t1[i][a] * t1[j][b];
          volatile float sv = v * 32768.f;
          long r = lrintf(sv);
          r = r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
          it[a * k + b] = (int16_t)r;
          isum += (int)r;
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int Mk1 = k2, Mk2 = k2, mk1 = k2, mk2 = k2;
        for (int a = k2; a < k2 + 2; ++a)
          for (int b = k2; b < k2 + 2; ++b) {
            if (it[a * k + b] < it[mk1 * k + mk2]) { mk1 = a; mk2 = b; }
            else if (it[a * k + b] > it[Mk1 * k + Mk2]) { Mk1 = a; Mk2 = b; }
          }
        if (diff < 0) it[Mk1 * k + Mk2] = (int16_t)(it[Mk1 * k + Mk2] - diff);
        else it[mk1 * k + mk2] = (int16_t)(it[mk1 * k + mk2] - diff);
      }
    }
  return B200AUG_OK;
}

extern "C" int b200aug_hamming_table(const double* windows, float* taps_out, uint64_t* sym_mask_out) {
  if (!windows || !taps_out || !sym_mask_out) return B200AUG_E_INVALID_ARG;
  uint64_t mask = 0;
  for (int r = 0; r < 32; ++r) {
    const int n = 2 * r + 1;
    bool sym = true;
    for (int i = 0; i < 64; ++i) {
      const double v = (r > 0 && i < n) ? windows[r * 64 + i] : 0.0;
      taps_out[r * 64 + i] = (float)v;
      if (r > 0 && i < n && v != windows[r * 64 + (n - 1 - i)]) sym = false;  // cv::getKernelType: exact mirror equality
    }
    if (sym && r > 0) mask |= 1ull << r;
  }
  *sym_mask_out = mask;
  return B200AUG_OK;
}

extern "C" int b200aug_upload_row_bands(uint8_t* dev_frames, const uint8_t* host_frames, int64_t frame_stride, int32_t pitch,
                                        int32_t batch, const int32_t* row_lo, const int32_t* row_hi, void* stream) {
  if (!dev_frames || !host_frames || frame_stride <= 0 || pitch <= 0 || batch < 0 || !row_lo || !row_hi) return B200AUG_E_INVALID_ARG;
  std::vector<void*> dsts, srcs;
  std::vector<size_t> sizes;
  dsts.reserve(batch); srcs.reserve(batch); sizes.reserve(batch);
  for (int i = 0; i < batch; ++i) {
    const int64_t lo = row_lo[i], hi = row_hi[i];
    if (lo < 0 || hi * pitch > frame_stride) return B200AUG_E_INVALID_ARG;
    if (hi <= lo) continue;
    const int64_t off = (int64_t)i * frame_stride + lo * pitch;
    dsts.push_back(dev_frames + off);
    srcs.push_back(const_cast<uint8_t*>(host_frames) + off);
    sizes.push_back((size_t)((hi - lo) * pitch));
  }
  if (dsts.empty()) return B200AUG_OK;
  // one batched submission (CUDA 12.8+); it is refused on the legacy default stream, where the copies go one by one
  static const bool one_by_one = getenv("B200AUG_UPLOAD_ONE_BY_ONE") != nullptr;
  cudaError_t e = cudaErrorNotSupported;
  if (stream && !one_by_one) {
    cudaMemcpyAttributes attr = {};
    attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    size_t attr_idx = 0, fail = 0;
    e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr, &attr_idx, 1, &fail, (cudaStream_t)stream);
    if (e != cudaSuccess) (void)cudaGetLastError();
  }
  if (e != cudaSuccess) {
    for (size_t i = 0; i < dsts.size(); ++i) {
      e = cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyHostToDevice, (cudaStream_t)stream);
      if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
    }
  }
  return B200AUG_OK;
}

extern "C" int b200aug_upload_boxes(uint8_t* dev_frames, const uint8_t* host_frames, int64_t frame_stride, int32_t pitch,
                                    int32_t batch, const int32_t* boxes, void* stream) {
  if (!dev_frames || !host_frames || frame_stride <= 0 || pitch <= 0 || batch < 0 || !boxes) return B200AUG_E_INVALID_ARG;
  static thread_local std::vector<cudaMemcpy3DBatchOp> ops;
  ops.clear();
  ops.reserve(batch);
  for (int i = 0; i < batch; ++i) {
    const int64_t x0 = boxes[4 * i], y0 = boxes[4 * i + 1], x1 = boxes[4 * i + 2], y1 = boxes[4 * i + 3];
    if (x0 < 0 || y0 < 0 || x1 > pitch || y1 * pitch > frame_stride) return B200AUG_E_INVALID_ARG;
    if (x1 <= x0 || y1 <= y0) continue;
    const int64_t off = (int64_t)i * frame_stride + y0 * pitch + x0;
    cudaMemcpy3DBatchOp op = {};
    op.src.type = cudaMemcpyOperandTypePointer;
    op.src.op.ptr.ptr = const_cast<uint8_t*>(host_frames) + off;
    op.src.op.ptr.rowLength = (size_t)pitch;
    op.src.op.ptr.layerHeight = (size_t)(y1 - y0);
    op.src.op.ptr.locHint.type = cudaMemLocationTypeHost;
    op.dst.type = cudaMemcpyOperandTypePointer;
    op.dst.op.ptr.ptr = dev_frames + off;
    op.dst.op.ptr.rowLength = (size_t)pitch;
    op.dst.op.ptr.layerHeight = (size_t)(y1 - y0);
    op.dst.op.ptr.locHint.type = cudaMemLocationTypeDevice;
    op.extent = make_cudaExtent((size_t)(x1 - x0), (size_t)(y1 - y0), 1);
    op.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    ops.push_back(op);
  }
  if (ops.empty()) return B200AUG_OK;
  cudaError_t e = cudaErrorNotSupported;
  static const bool one_by_one = getenv("B200AUG_UPLOAD_ONE_BY_ONE") != nullptr;
  if (stream && !one_by_one) {
    int dev = 0;
    (void)cudaGetDevice(&dev);
    for (auto& op : ops) op.dst.op.ptr.locHint.id = dev;
    size_t fail = 0;
    e = cudaMemcpy3DBatchAsync(ops.size(), ops.data(), &fail, 0, (cudaStream_t)stream);
    if (e != cudaSuccess) (void)cudaGetLastError();
  }
  if (e != cudaSuccess) {  // (legacy default stream, or a driver without the batched entry point)
    for (const auto& op : ops) {
      e = cudaMemcpy2DAsync(op.dst.op.ptr.ptr, (size_t)pitch, op.src.op.ptr.ptr, (size_t)pitch, op.extent.width, op.extent.height,
                            cudaMemcpyHostToDevice, (cudaStream_t)stream);
      if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
    }
  }
  return B200AUG_OK;
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
    if (c == B200AUG_CAT_BACKTRANSFORM && (d != 6 || f[i].count != 1)) return B200AUG_E_INVALID_ARG;
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
  if ((a.flags & B200AUG_F_FOCUS) && !a.explicit_view_roi && !a.explicit_tr) {
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
  const bool want_image = (a.flags & B200AUG_F_NORMALIZE) ? (a.image_f32_out != nullptr) : (a.image_u8_out != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  auto fail = [&](cudaError_t err) { g_last_cuda_error = (int)err; return B200AUG_E_CUDA; };
  KArgs K;
  K.a = a;
  int cl = a.cluster_size > 0 ? a.cluster_size : DEFAULT_CLUSTER;
  if (cl != 1 && cl != 2 && cl != 4 && cl != 8) return B200AUG_E_INVALID_ARG;

  // ---- kernel 1: plans, resize tables and labels of every sample (one small CTA each).  The record of a sample goes to
  // a.plans when the caller provides that scratch; otherwise the fused kernel rebuilds the plan for itself.
  const size_t rec = plan_bytes() + plan_tab_bytes(a.out_w, a.out_h);
  const bool records = want_image && a.plans != nullptr && rec <= 48 * 1024;  // (very large outputs: tables built in the fused kernel)
  if (a.plans && want_image) {
    if (a.plan_stride < (int64_t)rec || (a.plan_stride & 15) || (reinterpret_cast<uintptr_t>(a.plans) & 15)) return B200AUG_E_INVALID_ARG;
  }
  if (!records) K.a.plans = nullptr;
  // canvas workers: the first warp_ctas CTAs of the fused grid (a multiple of the cluster size); none without records +
  // workspace + a rotation parameter (rotated samples then take the per-pixel path)
  static thread_local int attr_dev = -1;     // per host thread and device: function attributes, SM count
  static thread_local int attr_smem = 0, n_sm = 0;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(e);
  if (dev != attr_dev) {
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return fail(e);
  }
  {
    int w = 0;
    if (records && a.workspace && (a.flags & B200AUG_F_FOCUS) && (a.angles || a.explicit_tr) && !a.explicit_view_roi && a.warp_ctas >= 0)
      w = (a.warp_ctas > 0) ? a.warp_ctas : (3 * n_sm) / 2;
    K.a.warp_ctas = (w + cl - 1) / cl * cl;
  }
  // anti-alias prefilters keep their smoothed canvases in the workspace and are planned by plan_kernel
  const bool prefilter = want_image && (a.flags & B200AUG_F_FOCUS) && a.downfilter != B200AUG_DOWN_AREA;
  if (a.downfilter < B200AUG_DOWN_AREA || a.downfilter > B200AUG_DOWN_HAMMING) return B200AUG_E_INVALID_ARG;
  if (a.upfilter < B200AUG_UP_LINEAR || a.upfilter > B200AUG_UP_LANCZOS) return B200AUG_E_INVALID_ARG;
  if (a.upfilter != B200AUG_UP_LINEAR && (a.flags & B200AUG_F_FOCUS) && want_image && !a.remap_tabs) return B200AUG_E_INVALID_ARG;
  if (prefilter && a.downfilter == B200AUG_DOWN_HAMMING && !a.hamming_taps) return B200AUG_E_INVALID_ARG;
  if (prefilter && (!records || !a.workspace)) return B200AUG_E_UNSUPPORTED;
  if (a.phase < B200AUG_PHASE_ALL || a.phase > B200AUG_PHASE_MAIN) return B200AUG_E_INVALID_ARG;
  if (a.phase == B200AUG_PHASE_MAIN && !records) return B200AUG_E_INVALID_ARG;  // nothing to pick the plans up from
  if (a.phase != B200AUG_PHASE_MAIN) {
    const size_t psmem = plan_bytes() + (records ? plan_tab_bytes(a.out_w, a.out_h) : 0) + ((sizeof(PlanCore) + 15) & ~size_t(15));
    plan_kernel<<<a.batch, NTHREADS, psmem, st>>>(K, records ? 1 : 0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e);
  }
  if (!want_image || a.phase == B200AUG_PHASE_PLAN) return B200AUG_OK;  // label-only call / plan phase

  const size_t smem = smem_layout(a.out_w, a.out_h, cap).total;
  if (smem > 227 * 1024) return B200AUG_E_SMEM;
  if (dev != attr_dev || (int)smem > attr_smem) {
    e = cudaFuncSetAttribute(fused_augment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fused_augment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    static const int carve = getenv("B200AUG_CARVEOUT") ? atoi(getenv("B200AUG_CARVEOUT")) : -1;
    if (carve >= 0) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fused_augment_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fused_augment_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(plan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    if (e != cudaSuccess) return fail(e);
    attr_dev = dev;
    attr_smem = (int)smem;
  }

  if (prefilter) {  // (a no-op for samples whose plan asks for no prefilter: those that do not down-scale)
    prefilter_kernel<<<a.batch * PF_CTAS, NTHREADS, 0, st>>>(K);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e);
  }
  // (MAIN alone, or behind prefilter_kernel: plain stream / event order)
  const int dep_mode = (records && a.phase == B200AUG_PHASE_ALL && !prefilter) ? 1 : 0;
  cudaLaunchAttribute pdl;
  pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl.val.programmaticStreamSerializationAllowed = 1;

  // ---- kernel 2: canvas workers + one cluster per sample (resampling, photometric chain)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(K.a.warp_ctas + a.batch * cl));
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (dep_mode != 0) {  // start while the kernel in front drains (see fused_augment_kernel: dep_mode)
    attr[1] = pdl;
    cfg.numAttrs = 2;
  }
  if (a.trace_out) e = cudaLaunchKernelEx(&cfg, fused_augment_kernel<true>, K, cap, dep_mode);
  else e = cudaLaunchKernelEx(&cfg, fused_augment_kernel<false>, K, cap, dep_mode);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(e);
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

extern "C" int b200aug_photometric_f32(const float* in, float* out, float* tmp, int batch, int width, int height,
                                       const B200AugPhotoParams* photo, float bias, void* stream) {
  if (!in || !out || !photo || batch < 0 || width <= 0 || height <= 0) return B200AUG_E_INVALID_ARG;
  if (photo->n_order < 0 || photo->n_order > B200AUG_NUM_OPS || (photo->n_order > 0 && !photo->apply)) return B200AUG_E_INVALID_ARG;
  bool blur = false;
  for (int k = 0; k < photo->n_order; ++k) {
    if (photo->order[k] < 0 || photo->order[k] >= B200AUG_NUM_OPS) return B200AUG_E_INVALID_ARG;
    blur = blur || photo->order[k] == B200AUG_OP_BLUR;
  }
  if (blur && (!tmp || tmp == out || tmp == in)) return B200AUG_E_INVALID_ARG;  // the blur ping-pongs between out and tmp
  if (blur && in == out) return B200AUG_E_INVALID_ARG;
  if (batch == 0) return B200AUG_OK;
  photometric_f32_kernel<<<batch, NTHREADS, 0, (cudaStream_t)stream>>>(in, out, tmp, width, height, *photo, bias);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}

extern "C" int b200aug_corrected_rotation(const float* half_sizes, int64_t size_stride, float div_x, float div_y, float f,
                                          const float* coord, int64_t coord_stride, const float* pose, float* out,
                                          float* look_at_out, int batch, void* stream) {
  if (!coord || batch < 0 || (!out && !look_at_out) || (out && !pose) || size_stride < 0 || coord_stride < (half_sizes ? 2 : 3))
    return B200AUG_E_INVALID_ARG;
  if (batch == 0) return B200AUG_OK;
  corrected_rotation_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(half_sizes, size_stride, div_x, div_y, f, coord,
                                                                                  coord_stride, pose, out, look_at_out, batch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}

extern "C" int b200aug_head_roi(const float* vertices, const float* deform_base, int32_t n_vertices, int32_t n_params,
                                const float* shapeparams, const float* coord, const float* quat, float xy_offset, float* roi_out,
                                int32_t batch, void* stream) {
  if (!vertices || n_vertices <= 0 || !coord || !quat || !roi_out || batch < 0) return B200AUG_E_INVALID_ARG;
  if (n_params < 0 || n_params > HR_MAXK) return B200AUG_E_INVALID_ARG;
  if (shapeparams && n_params > 0 && !deform_base) return B200AUG_E_INVALID_ARG;
  if (batch == 0) return B200AUG_OK;
  head_roi_kernel<<<(batch + HR_SPB - 1) / HR_SPB, NTHREADS, 0, (cudaStream_t)stream>>>(vertices, deform_base, n_vertices, n_params,
                                                                                          shapeparams, coord, quat, xy_offset, roi_out, batch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}

extern "C" int b200aug_quat_matrix(const float* in, float* out, int batch, int to_matrix, void* stream) {
  if (!in || !out || batch < 0) return B200AUG_E_INVALID_ARG;
  if (batch == 0) return B200AUG_OK;
  quat_matrix_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, out, batch, to_matrix);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}
