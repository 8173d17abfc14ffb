// JPEG -> grayscale source frames on the GPU: the device-side replacement of the reference's
// `imdecode(blob, color=False)` = cv2.imdecode(blob, 0) (trackertraincode/datasets/preprocessing.py:42-54), which feeds
// the frames of the HDF5 `varsize_image_buffer` format (datasets/dshdf5.py:59-113) into the augmentation path.
//
// Library code: the entropy decoding / IDCT is nvJPEG's (batched decode, luminance plane only -- for a YCbCr JPEG the Y
// plane IS what libjpeg returns for a grayscale decode, up to IDCT rounding).  What this file adds is the C ABI around it:
// caller-owned device frames (stacked or ragged), stream ordering, one handle per host thread.
#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <stdint.h>

#include <cstdlib>
#include <thread>
#include <vector>

#include "b200aug.h"

namespace {

struct JpegCtx {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int batch = 0;      // batch size the state is initialised for
  int status = 0;     // last nvjpegStatus_t that failed
  bool ok = false;
  bool hardware = false;
};

JpegCtx& ctx() {
  static thread_local JpegCtx c;
  if (!c.handle) {
    // backend: the GPU's NVJPG engines when B200AUG_JPEG_BACKEND=hardware (and the library accepts it), else nvJPEG's hybrid
    // decoder (Huffman decoding on host threads, IDCT on the SMs)
    const char* be = getenv("B200AUG_JPEG_BACKEND");
    nvjpegStatus_t s = NVJPEG_STATUS_NOT_INITIALIZED;
    if (be && be[0] == 'h') {
      s = nvjpegCreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &c.handle);
      c.hardware = (s == NVJPEG_STATUS_SUCCESS);
    }
    if (s != NVJPEG_STATUS_SUCCESS) s = nvjpegCreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &c.handle);
    if (s != NVJPEG_STATUS_SUCCESS) s = nvjpegCreateSimple(&c.handle);
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegJpegStateCreate(c.handle, &c.state);
    c.status = (int)s;
    c.ok = (s == NVJPEG_STATUS_SUCCESS);
  }
  return c;
}

}  // namespace

extern "C" int b200aug_jpeg_last_status(void) { return ctx().status; }

extern "C" int b200aug_jpeg_backend(void) { return ctx().ok ? (ctx().hardware ? 2 : 1) : 0; }

extern "C" int b200aug_jpeg_info(const uint8_t* data, size_t length, int32_t* width, int32_t* height, int32_t* components) {
  if (!data || !length || !width || !height) return B200AUG_E_INVALID_ARG;
  JpegCtx& c = ctx();
  if (!c.ok) return B200AUG_E_CUDA;
  int ncomp = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t ss;
  nvjpegStatus_t s = nvjpegGetImageInfo(c.handle, data, length, &ncomp, &ss, w, h);
  if (s != NVJPEG_STATUS_SUCCESS) { c.status = (int)s; return B200AUG_E_INVALID_ARG; }
  *width = w[0];
  *height = h[0];
  if (components) *components = ncomp;
  return B200AUG_OK;
}

extern "C" int b200aug_decode_jpeg_gray(const uint8_t* const* data, const size_t* lengths, int32_t batch, uint8_t* const* dst,
                                        const int32_t* pitch, void* stream) {
  if (batch < 0 || (batch > 0 && (!data || !lengths || !dst || !pitch))) return B200AUG_E_INVALID_ARG;
  if (batch == 0) return B200AUG_OK;
  JpegCtx& c = ctx();
  if (!c.ok) return B200AUG_E_CUDA;
  nvjpegStatus_t s;
  if (c.batch != batch) {
    // the hybrid backend decodes the entropy-coded segments on host threads: give it the cores this process may use
    static const int n_threads = [] {
      const char* e = getenv("B200AUG_JPEG_THREADS");
      int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
      return n < 1 ? 1 : (n > 32 ? 32 : n);
    }();
    s = nvjpegDecodeBatchedInitialize(c.handle, c.state, batch, n_threads, NVJPEG_OUTPUT_Y);
    if (s != NVJPEG_STATUS_SUCCESS) { c.status = (int)s; c.batch = 0; return B200AUG_E_CUDA; }
    c.batch = batch;
  }
  std::vector<nvjpegImage_t> out(batch);
  for (int i = 0; i < batch; ++i) {
    if (!data[i] || !lengths[i] || !dst[i] || pitch[i] <= 0) return B200AUG_E_INVALID_ARG;
    for (int k = 0; k < NVJPEG_MAX_COMPONENT; ++k) { out[i].channel[k] = nullptr; out[i].pitch[k] = 0; }
    out[i].channel[0] = dst[i];
    out[i].pitch[0] = (size_t)pitch[i];
  }
  s = nvjpegDecodeBatched(c.handle, c.state, data, lengths, out.data(), (cudaStream_t)stream);
  if (s != NVJPEG_STATUS_SUCCESS) { c.status = (int)s; return B200AUG_E_CUDA; }
  return B200AUG_OK;
}
