/*
 * b200aug.h -- C ABI of the B200-native augmentation / label-transform hot path.
 *
 * The reference (opentrack/neuralnet-tracker-traincode) has no native or FFI interface for this path: the
 * boundary is the Python protocol `Callable[[Batch], Batch]`.  Each entry point below names the
 * reference function(s) (file:line, relative to the reference checkout) whose arithmetic it replaces; the Python
 * mirror of the reference API (neuralnet-tracker-traincode_b200/trackertraincode_b200) binds them with ctypes,
 * see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors); nothing is allocated or freed here;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and is
 *     asynchronous; calls on distinct streams are thread-safe;
 *   - return value: 0 on success, a B200AUG_E_* code otherwise (b200aug_strerror() explains it);
 *   - per-sample problems (empty crop box, unsupported resampler) do not fail the launch: they are reported in the
 *     optional device array `status_out[B]` and the sample's image is zero-filled;
 *   - images are single-channel (the pose pipeline loads monochrome, dshdf5pose.py:201); layouts:
 *       source  uint8  [H, W]      row pitch in bytes, one descriptor per sample (ragged batches allowed)
 *       crop    uint8  [B, 1, oh, ow]   or   float32 [B, 1, oh, ow]
 *       roi [x0,y0,x1,y1] - coord [x,y,size] - quaternion [i,j,k,w] - points [n, 2|3]   (all float32)
 */
#ifndef B200AUG_H_
#define B200AUG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200AUG_ABI_VERSION 7

/* error codes */
#define B200AUG_OK 0
#define B200AUG_E_INVALID_ARG 1   /* NULL where a pointer is required, non-positive sizes, bad flag combination */
#define B200AUG_E_UNSUPPORTED 2   /* valid request the kernels do not cover (e.g. rot90 on non-square output) */
#define B200AUG_E_SMEM 3          /* requested row-buffer capacity does not fit in 227 KB of shared memory */
#define B200AUG_E_CUDA 4          /* the CUDA runtime reported an error at launch; see b200aug_last_cuda_error() */

/* per-sample status codes written to status_out */
#define B200AUG_S_OK 0
#define B200AUG_S_EMPTY_BOX 1     /* view box with non-positive width/height (cv2.resize would throw) */
#define B200AUG_S_UNSUPPORTED 2   /* an anti-alias prefilter wider than B200AUG_PREFILTER_MAX_TAPS (scale below ~0.05), or a
                                     prefiltered canvas that does not fit its workspace region: the image is zeros */
#define B200AUG_S_ROWBUF 3        /* reserved (ABI 1 reported row-buffer overflow; since ABI 2 such samples take the per-pixel path) */

/* field categories: FieldCategory, trackertraincode/datasets/dshdf5pose.py:21-28 */
#define B200AUG_CAT_GENERAL 0     /* ""    passes through unchanged */
#define B200AUG_CAT_QUAT 1        /* "q"   affinetrafo.py:98-127  transform_rot */
#define B200AUG_CAT_XYS 2         /* "xys" affinetrafo.py:89-95   transform_coord */
#define B200AUG_CAT_ROI 3         /* "roi" affinetrafo.py:75-86   transform_roi */
#define B200AUG_CAT_POINTS 4      /* "pts" affinetrafo.py:37-72   transform_keypoints (68-landmark flip_map) */
#define B200AUG_CAT_BACKTRANSFORM 5 /* the "image_backtransform" entry (a [2,3] affine, count 1, dim 6): every transform tr
                                     rewrites it as BT @ tr^-1, affinetrafo.py:137-147 */

/* stage flags of b200aug_fused_forward, in pipeline order (trackertraincode/pipelines.py:372-383, 508-532) */
#define B200AUG_F_HALF_PIXEL 0x01u         /* offset_points_by_half_pixel       batch/normalization.py:83-90 */
#define B200AUG_F_ROI_FROM_LANDMARKS 0x02u /* PutRoiFromLandmarks before+after  batch/misc.py:9-31 (no forehead ext.) */
#define B200AUG_F_FOCUS 0x04u              /* GeneralFocusRoi.__call__          batch/geometric.py:193-231 */
#define B200AUG_F_FLIPROT 0x08u            /* horizontal_flip_and_rot_90        batch/geometric.py:234-267 */
#define B200AUG_F_NORMALIZE 0x10u          /* normalize_batch                   batch/normalization.py:20-56 */
#define B200AUG_F_PHOTOMETRIC 0x20u        /* KorniaImageDistortions x2         batch/intensity.py:30-64, pipelines.py:508-527 */
#define B200AUG_F_WHITEN 0x40u             /* whiten_batch                      batch/normalization.py:94-99 */
#define B200AUG_F_INSERT_BACKTRANSFORM 0x80u /* GeneralFocusRoi(insert_backtransform=True), batch/geometric.py:226-227: the
                                              focus stage (re)starts the back-transform as tr^-1; later stages of the
                                              same call compose onto it; written to backtransform_out */

/* B200AugFusedArgs::phase */
/* B200AugFusedArgs::downfilter -- DownFilters of tensors/image_geometric_cv2.py:15, used when a crop shrinks (:65-82):
 * "area" = cv2.resize(INTER_AREA); the other two smooth the canvas first (_apply_antialias_filter, :47-62) and resize it
 * with INTER_LINEAR */
#define B200AUG_DOWN_AREA 0
#define B200AUG_DOWN_GAUSSIAN 1   /* the reference's cv2.GaussianBlur call as Python binds its positional arguments (:50):
                                     sigmaX = 0.5 / scale, sigmaY = 1, BORDER_REFLECT_101; 8-bit fixed point, bit-exact */
#define B200AUG_DOWN_HAMMING 2    /* cv2.sepFilter2D with the normalised Hamming window of round(2 / scale + 1) taps (made
                                     odd): float32, the arithmetic of cv2's vector loops */
#define B200AUG_PREFILTER_MAX_TAPS 63
/* B200AugFusedArgs::upfilter -- UpFilters of tensors/image_geometric_cv2.py:16, used when a crop grows: the interpolation of
 * cv2.resize (:68-75) or, for a rotated sample whose transform up-scales, of cv2.warpAffine (:105-119).  OpenCV's own
 * fixed-point kernels, bit-exact (one caveat: x86 wheels route cv2.resize(INTER_CUBIC) through Intel IPP, which differs from
 * OpenCV's kernel by at most one grey level on a few per cent of the pixels). */
#define B200AUG_UP_LINEAR 0
#define B200AUG_UP_CUBIC 1
#define B200AUG_UP_LANCZOS 2

#define B200AUG_PHASE_ALL 0
#define B200AUG_PHASE_PLAN 1
#define B200AUG_PHASE_MAIN 2

#define B200AUG_MAX_FIELDS 8
#define B200AUG_NUM_OPS 6
#define B200AUG_NUM_NOISE 4

/* stage-1 photometric op ids (order of the AugmentationSequential children, pipelines.py:511-518) */
#define B200AUG_OP_EQUALIZE 0
#define B200AUG_OP_POSTERIZE 1
#define B200AUG_OP_GAMMA 2
#define B200AUG_OP_CONTRAST 3
#define B200AUG_OP_BRIGHTNESS 4
#define B200AUG_OP_BLUR 5

/* One source image (device memory). */
typedef struct B200AugSrc {
  const uint8_t* ptr; /* top-left pixel */
  int32_t width;
  int32_t height;
  int32_t pitch;      /* bytes between rows */
  int32_t reserved;
} B200AugSrc;

/* One label tensor [B, count, dim] float32, transformed according to `category`. in == out is allowed for every
 * category except POINTS with count == 68 (the mirror permutation reads other rows).  BACKTRANSFORM: count 1, dim 6. */
typedef struct B200AugField {
  int32_t category;
  int32_t count;      /* items per sample (68 landmarks, 1 roi, ...) */
  int32_t dim;        /* floats per item: roi 4, xys 3, quat 4, points 2 or 3, general any */
  int32_t reserved;
  const float* in;
  float* out;
} B200AugField;

/* Sampled parameters of the two photometric stages for one call (pipelines.py:510-527).  `order` lists the
 * stage-1 children chosen by random_apply (one draw per call), in application order. */
typedef struct B200AugPhotoParams {
  int32_t n_order;
  int32_t order[B200AUG_NUM_OPS];
  int32_t clip;                 /* OnlyClip(p=1): clamp to [0,1] after the noise stages */
  const uint8_t* apply;         /* [B, 6] per-sample Bernoulli masks, indexed by op id */
  const int32_t* bits;          /* [B] posterize bits */
  const float* gamma;           /* [B] */
  const float* contrast;        /* [B] */
  const float* brightness;      /* [B] */
  const uint8_t* noise_apply;   /* [B, 4] */
  float noise_std[B200AUG_NUM_NOISE];
  int32_t noise_clip[B200AUG_NUM_NOISE]; /* RandomGaussianNoiseWithClipping (batch/intensity.py:43-53): clamp to [0,1] right behind
                                   this stage, on the samples it was applied to */
  uint64_t seed;                /* Philox4x32-10 key */
  uint64_t sample_offset;       /* id of sample 0 in the noise stream (rank * local batch + step * global batch ...) */
} B200AugPhotoParams;

typedef struct B200AugFusedArgs {
  int32_t struct_size;          /* sizeof(B200AugFusedArgs), checked */
  int32_t batch;
  int32_t out_w, out_h;         /* crop size (129 x 129 for the pose net) */
  uint32_t flags;               /* B200AUG_F_* */
  int32_t rowbuf_capacity;      /* bytes of per-warp staging (ring of source rows in flight); 0 = default (2800) */

  /* sources: a table of descriptors (ragged batch) or, if NULL, one descriptor + stride (stacked [B,H,W] tensor) */
  const B200AugSrc* src_table;
  B200AugSrc src_uniform;
  int64_t src_stride;           /* bytes between consecutive images when src_table == NULL */

  /* RoiFocusRandomizationParameters (batch/geometric.py:27-32), device arrays */
  const float* scales;          /* [B] */
  const float* angles;          /* [B] radians; pixel path is croprescale iff angle == 0 (geometric.py:210) */
  const float* cos_sin;         /* optional [B,2]: host-evaluated torch.cos/sin(angles) (affine2d.py:46-47) */
  const float* translations;    /* [B,2] */
  float beyond_border_shift;    /* 0.3, geometric.py:104 */
  /* explicit geometry instead of the sampled one (the tensor-level entries of tensors/image_geometric_cv2.py):
   *   explicit_view_roi [B,4] int32: croprescale_image_cv2(img, roi, new_size) (:138-155) -- the integer box is cropped
   *                     (zero padded) and resized; labels follow range_remap(box -> output);
   *   explicit_tr [B,2,3]: affine_transform_image_cv2(img, tr, new_size) (:85-135) -- always the warpAffine path.
   * With either one scales / angles / translations / the roi field are not read. */
  const int32_t* explicit_view_roi;
  const float* explicit_tr;
  /* horizontal_flip_and_rot_90 draws (batch/geometric.py:236-237) */
  const uint8_t* do_flip;       /* [B] or NULL */
  const int8_t* rot_dir;        /* [B] in {-1,0,1} or NULL */

  /* labels */
  int32_t n_fields;
  int32_t roi_field;            /* index into fields of the focus box (roi_variable), -1 if F_ROI_FROM_LANDMARKS only */
  int32_t landmark_field;       /* index of pt3d_68 for F_ROI_FROM_LANDMARKS, else -1 */
  int32_t cluster_size;         /* CTAs (thread-block cluster) that share one sample: 1, 2, 4 or 8; 0 = default (2) */
  B200AugField fields[B200AUG_MAX_FIELDS];

  /* outputs (each may be NULL) */
  int32_t* view_roi_out;        /* [B,4] rounded view box, geometric.py:205 */
  float* tr_out;                /* [B,2,3] focus transform, geometric.py:206-207 */
  float* backtransform_out;     /* [B,2,3] with F_INSERT_BACKTRANSFORM: tr^-1 of the focus stage (geometric.py:226-227) composed
                                   with the inverse of every later stage of this call (affinetrafo.py:137-147) */
  uint8_t* image_u8_out;        /* [B,1,oh,ow] when F_NORMALIZE is not set */
  float* image_f32_out;         /* [B,1,oh,ow] when F_NORMALIZE is set */
  int32_t* status_out;          /* [B] B200AUG_S_* */
  uint64_t* trace_out;          /* [B*cluster_size,16] per-CTA timeline for profiling: %globaltimer (ns) at start / plan built / cluster
                                   synchronised / resample done / end, then %smid, (unused), canvas ready, resize tables
                                   built, (unused), photometric LUT built; the rest 0 */
  /* optional launch order: order[i] = sample processed by the i-th cluster of the grid (a permutation of 0..B-1).  The
   * CTAs are dispatched in grid order, so listing the expensive samples (rotated, blurred, noisy) first lets the cheap
   * ones fill the tail.  NULL = identity. */
  const int32_t* order;
  /* optional scratch for rotated samples: B regions of workspace_stride bytes (see b200aug_workspace_stride()).  The
   * two-stage rotated path (warpAffine canvas, then INTER_AREA; image_geometric_cv2.py:121-134) keeps its canvas here,
   * i.e. in L2: some CTAs of the fused grid ("canvas workers", needs `plans` too) fill the canvases while the others already
   * resample the unrotated samples.  Without it (NULL) or when a canvas does not fit, canvas pixels are produced one at a
   * time instead.  Anti-alias prefilters (downfilter) keep their smoothed canvases here as well. */
  uint8_t* workspace;
  int64_t workspace_stride;
  /* optional scratch for the plans: b200aug_plan_buffer_bytes(batch, out_w, out_h) bytes = B records of plan_stride bytes
   * (= b200aug_plan_stride(out_w, out_h), a multiple of 16) followed by a small tail (work counters and per-sample flags
   * of the canvas workers).  A small kernel computes every sample's plan, cv2 resize tables and LABELS first (it always runs: the
   * labels are its output); with this buffer it also leaves the records for the fused kernel to load, instead of every CTA
   * rebuilding them behind its full register / shared-memory footprint.  NULL = build them inside the fused kernel. */
  uint8_t* plans;
  int64_t plan_stride;
  int32_t warp_ctas;            /* CTAs of the fused grid that produce the rotated samples' canvases instead of taking a sample
                                   (canvas workers); 0 = default (1.5 per SM), < 0 = none */
  int32_t phase;                /* B200AUG_PHASE_*: 0 = the whole call; PLAN = plan_kernel only (plans + tables + labels + side
                                   outputs), MAIN = everything behind it (needs the records a PLAN call with the same arguments
                                   left in `plans`).  A caller that pipelines steps runs PLAN of step s + 1 on a second stream
                                   next to MAIN of step s: plan_kernel is a short latency-bound grid that fits into the tail of
                                   the big kernel. */

  int32_t downfilter;           /* B200AUG_DOWN_*; anything but AREA needs `plans` and `workspace` (else E_UNSUPPORTED) */
  int32_t upfilter;             /* B200AUG_UP_*; anything but LINEAR needs `remap_tabs` */
  /* B200AUG_DOWN_HAMMING: the normalised Hamming windows as float32, row r = (n - 1) / 2 of a [32][64] device table holds
   * the n = 2 r + 1 taps (scipy.signal.windows.hamming(n) / sum, evaluated on the host exactly as the reference does and
   * then rounded to float32, as cv2 does); bit r of hamming_sym_mask says whether the float64 window is exactly mirror
   * symmetric, which is what makes cv2 pick its symmetric column filter (that depends on the host's cos(), so the caller
   * decides).  b200aug_hamming_table() fills both from a given evaluation of the windows. */
  const float* hamming_taps;
  uint64_t hamming_sym_mask;
  /* B200AUG_UP_CUBIC / _LANCZOS: cv2's fixed-point interpolation tables for warpAffine, device memory: 1024 x 16 shorts
   * (cubic) followed by 1024 x 64 shorts (Lanczos), as b200aug_remap_table() computes them */
  const int16_t* remap_tabs;

  B200AugPhotoParams photo;     /* read when F_PHOTOMETRIC is set */
} B200AugFusedArgs;

int b200aug_abi_version(void);
const char* b200aug_strerror(int code);
/* cudaError_t (as int) of the last failed launch on this host thread, 0 if none */
int b200aug_last_cuda_error(void);
/* dynamic shared memory (bytes) one CTA of the fused kernel uses for this geometry; 0 if it cannot fit */
size_t b200aug_fused_smem_bytes(int out_w, int out_h, int rowbuf_capacity);
/* what the driver will co-schedule of the fused kernel for this geometry on the current device: resident CTAs per SM
 * (registers / shared memory) and the number of clusters of `cluster_size` (0 = default) CTAs active at once
 * (cudaOccupancyMaxActiveClusters) */
int b200aug_fused_occupancy(int out_w, int out_h, int rowbuf_capacity, int cluster_size, int* ctas_per_sm, int* active_clusters);
/* bytes of scratch per sample that hold the rotated canvas of a crop box of up to max_side x max_side source pixels */
int64_t b200aug_workspace_stride(int max_side);
/* bytes of one record of B200AugFusedArgs::plans for this output size, and of the whole buffer for `batch` samples */
int64_t b200aug_plan_stride(int out_w, int out_h);
int64_t b200aug_plan_buffer_bytes(int batch, int out_w, int out_h);
/* Packs Hamming windows for B200AugFusedArgs::hamming_taps: windows = HOST float64, the normalised window of n = 2 r + 1
 * taps at windows[r * 64 .. r * 64 + n) for r = 1 .. 31 (row 0 unused); taps_out = HOST float32 [32 * 64] (copy it to the
 * device), *sym_mask_out = the symmetry bits.  Plain host code, no CUDA call. */
int b200aug_hamming_table(const double* windows, float* taps_out, uint64_t* sym_mask_out);
/* cv::initInterTab2D for B200AugFusedArgs::remap_tabs: out = HOST shorts, 1024 x 16 for B200AUG_UP_CUBIC, 1024 x 64 for
 * B200AUG_UP_LANCZOS (phase (fy, fx) at (fy * 32 + fx) * k * k).  Plain host code, no CUDA call. */
int b200aug_remap_table(int upfilter, int16_t* out);

/* Host -> device upload of the rows the fused kernel will read, instead of whole frames (Batch.to(device),
 * datasets/batch.py:161-165 / pipelines.py:508): for each of `batch` stacked frames (host_frames: PINNED host memory,
 * [batch] frames of frame_stride bytes, rows of `pitch` bytes) the rows [row_lo[i], row_hi[i]) are copied to the same
 * offsets of dev_frames (device, same layout) with the copy engine, stream-ordered.  row_lo / row_hi are HOST arrays; rows
 * outside the band keep whatever dev_frames held.  The caller derives the bands from the sampled view boxes. */
int b200aug_upload_row_bands(uint8_t* dev_frames, const uint8_t* host_frames, int64_t frame_stride, int32_t pitch,
                             int32_t batch, const int32_t* row_lo, const int32_t* row_hi, void* stream);

/* The same for boxes: rows [y0, y1) x columns [x0, x1) of frame i (boxes = HOST int32 [batch,4] = x0, y0, x1, y1, already
 * clipped to the frame) -- one batched 2-D copy (cudaMemcpy3DBatchAsync), i.e. only the pixels the view boxes can touch
 * cross PCIe. */
int b200aug_upload_boxes(uint8_t* dev_frames, const uint8_t* host_frames, int64_t frame_stride, int32_t pitch, int32_t batch,
                         const int32_t* boxes, void* stream);

/* The fused hot path.  Two kernels (three with an anti-alias prefilter) on `stream`, the second behind a programmatic
 * dependent launch:
 *   plan_kernel   one small CTA per sample: view box, transforms, cv2 resize tables, all label transforms, side outputs
 *   prefilter_kernel   (downfilter gaussian / hamming only) the smoothed canvases of the down-scaled samples
 *   fused_augment_kernel   one cluster per sample: cv2.resize -> flip/rot90 -> normalise -> photometric chain -> whiten;
 *                 plus the canvas workers (rotated samples, needs plans + workspace): CTAs of the same grid that produce the
 *                 cv2.warpAffine canvases, handing out work items through an atomic counter
 * A call without an image output (labels only) runs plan_kernel alone.
 * Replaces, per the flags: batch/normalization.py:83-90, batch/misc.py:9-31, batch/geometric.py:107-231
 * (+ tensors/image_geometric_cv2.py:28-155 incl. cv2.warpAffine / cv2.resize arithmetic, tensors/affinetrafo.py:37-148),
 * batch/geometric.py:234-267, batch/normalization.py:20-56, batch/intensity.py:30-64, batch/normalization.py:94-99. */
int b200aug_fused_forward(const B200AugFusedArgs* args, void* stream);

/* apply_affine2d (tensors/affinetrafo.py:130-148) on label tensors with an explicit transform per sample
 * (tr [B,2,3], or one [2,3] broadcast when tr_stride == 0). */
int b200aug_apply_affine2d(const float* tr, int64_t tr_stride, int batch, int n_fields, const B200AugField* fields,
                           void* stream);

/* KorniaImageDistortions.__call__ on its own (batch/intensity.py:30-40 with the op lists of pipelines.py:510-527):
 * both photometric stages on float32 images [B,1,h,w] that already live on the device (the fused kernel covers the case
 * where they follow the crop).  Stage 1 = photo->order / apply / factors (n_order = 0 skips it); stage 2 = noise_apply
 * (NULL skips it) + clip.  `bias` is added last (-0.5 fuses whiten_batch, batch/normalization.py:94-99; 0 otherwise).
 * `tmp` [B,h,w] is scratch, required (and distinct from in/out) when the blur is among the ordered ops; in == out is
 * allowed otherwise. */
int b200aug_photometric_f32(const float* in, float* out, float* tmp, int batch, int width, int height,
                            const B200AugPhotoParams* photo, float bias, void* stream);

/* PerspectiveCorrector.corrected_rotation (eval.py:491-529) with _make_look_at_matrix (eval.py:531-544),
 * torchquaternion.from_matrix (neuralnets/torchquaternion.py:94-168) and torchquaternion.mult (:40-48):
 *   xy_n = (coord.xy - half_size) / (div_x, div_y);  M = look_at([xy_n, f]);  out = from_matrix(M) (x) pose   (xyzw).
 * half_sizes: [B,2] (size_stride 2) or one [2] shared by the batch (size_stride 0); coord rows are coord_stride floats
 * apart (>= 2).  The reference divides BOTH axes by `half_image_size_tensor[0]` (eval.py:525) -- the half width for a [2]
 * size, row 0 of the batch for a [B,2] size; the caller passes that divisor pair.  out [B,4] and / or look_at_out [B,3,3]
 * (row-major, columns x, y, z) are written; either may be NULL.  With half_sizes == NULL the rows of coord (>= 3 floats)
 * are taken as the look-at positions themselves (PerspectiveCorrector._make_look_at_matrix on its own). */
int b200aug_corrected_rotation(const float* half_sizes, int64_t size_stride, float div_x, float div_y, float f,
                               const float* coord, int64_t coord_stride, const float* pose, float* out, float* look_at_out,
                               int batch, void* stream);

/* PutRoiFromLandmarks(extend_to_forehead=True) (batch/misc.py:14-26): roi_out[b] = [min_x, min_y, max_x, max_y] over all
 * vertices of the posed deformable face model (PosedDeformableHead, neuralnets/modelcomponents.py:38-56,85-94):
 *   (quat[b] rotates (vertices + sum_k deform_base[k] * shapeparams[b][k])) * coord[b][2] + (coord[b][0:2] + xy_offset).
 * vertices [V,3] and deform_base [K,V,3] (K <= 64) are the scaled BFM arrays of facemodel/bfm.py:49-72, device memory;
 * shapeparams [B,K] may be NULL (= zeros, which is what the reference effectively uses: misc.py:15-17 looks for a key
 * "shapeparams" no dataset has); coord [B,3], quat [B,4] (i, j, k, w).  xy_offset = 0.5 when the labels have not been
 * through offset_points_by_half_pixel yet (batch/normalization.py:83-90), else 0. */
int b200aug_head_roi(const float* vertices, const float* deform_base, int32_t n_vertices, int32_t n_params,
                     const float* shapeparams, const float* coord, const float* quat, float xy_offset, float* roi_out,
                     int32_t batch, void* stream);

/* torchquaternion.tomatrix (to_matrix != 0: in [B,4] xyzw -> out [B,3,3]; torchquaternion.py:70-91, the matrix / 6D
 * rotation target of losses.py:53-58) or torchquaternion.from_matrix (to_matrix == 0: in [B,3,3] -> out [B,4]; :94-168). */
int b200aug_quat_matrix(const float* in, float* out, int batch, int to_matrix, void* stream);

/* JPEG -> grayscale source frames on the device: replaces `imdecode(blob, color=False)` = cv2.imdecode(blob, 0)
 * (datasets/preprocessing.py:42-54) for the JPEG blobs of the HDF5 `varsize_image_buffer` format (datasets/dshdf5.py:59-113).
 * Entropy decoding and IDCT are nvJPEG's (batched, luminance plane only); results match cv2 to the IDCT rounding (+-2 grey
 * levels).  data[i] / lengths[i]: HOST pointers to the JPEG streams; dst[i]: DEVICE pointer to frame i (rows of pitch[i]
 * bytes, size from b200aug_jpeg_info); the arrays data / lengths / dst / pitch themselves are host arrays.  Stream-ordered;
 * one decoder state per host thread.  b200aug_jpeg_last_status() returns the nvjpegStatus_t of the last failure. */
int b200aug_jpeg_info(const uint8_t* data, size_t length, int32_t* width, int32_t* height, int32_t* components);
int b200aug_decode_jpeg_gray(const uint8_t* const* data, const size_t* lengths, int32_t batch, uint8_t* const* dst,
                             const int32_t* pitch, void* stream);
int b200aug_jpeg_last_status(void);
/* 0 = nvJPEG unavailable, 1 = hybrid GPU backend, 2 = the NVJPG hardware engines (env B200AUG_JPEG_BACKEND=hardware) */
int b200aug_jpeg_backend(void);

#ifdef __cplusplus
}
#endif
#endif /* B200AUG_H_ */
