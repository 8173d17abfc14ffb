// Microbenchmark: how fast can one SM stage small row segments global -> shared?
//   mode 0: cp.async.bulk (TMA engine, 1-D), one copy of ROWB bytes per row, issued by the lanes of warp 0, mbarrier completion
//   mode 1: cp.async 16 B (LDGSTS) by all 256 threads, commit / wait_group
// Each CTA streams `tiles` boxes of `rows` rows (pitch 450 B apart, 16-byte aligned supersets) through a 3-deep ring.
// Prints cycles per tile per CTA for 1 / 2 / 3 CTAs per SM.   nvcc -arch=sm_100a -O3 -o bulk_rate bulk_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NST = 3, STAGE = 12288;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) k(const uint8_t* src, size_t span, int tiles, int rows, int rowb, int mode, unsigned long long* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  uint8_t* st = sm + 128;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) for (int i = 0; i < NST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + i)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const uint8_t* base = src + ((size_t)blockIdx.x * 7919 * 450) % span;
  auto issue = [&](int t) {
    const uint8_t* g0 = base + ((size_t)t * 29 * 450) % (span / 2);
    uint8_t* d0 = st + (t % NST) * STAGE;
    if (mode == 0) {
      if (warp == 0) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + t % NST)), "r"(rows * rowb) : "memory");
        __syncwarp();
        for (int r = lane; r < rows; r += 32) {
          const uintptr_t g = (reinterpret_cast<uintptr_t>(g0) + (size_t)r * 450) & ~uintptr_t(15);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d0 + r * rowb)),
                       "l"(g), "r"(rowb), "r"(s32(bar + t % NST)) : "memory");
        }
      }
    } else {
      const int nch = rowb / 16;
      for (int it = tid; it < rows * 8; it += 256) {
        const int r = it >> 3, c = it & 7;
        if (c < nch) {
          const uintptr_t g = ((reinterpret_cast<uintptr_t>(g0) + (size_t)r * 450) & ~uintptr_t(15)) + 16 * c;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d0 + r * rowb + 16 * c)), "l"(g) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  unsigned acc = 0;
  const long long t0 = clock64();
  issue(0); issue(1);
  for (int t = 0; t < tiles; ++t) {
    if (t + 2 < tiles) issue(t + 2);
    else if (mode == 1) asm volatile("cp.async.commit_group;" ::: "memory");
    if (mode == 0) {
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(bar + t % NST)), "r"((t / NST) & 1) : "memory");
    } else {
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncthreads();
    }
    acc += st[(t % NST) * STAGE + tid];  // touch the data
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0) + (acc == 0xffffffffu);
}
int main() {
  const size_t span = 512ull << 20;
  uint8_t* src; cudaMalloc(&src, span + (1 << 20)); cudaMemset(src, 1, span + (1 << 20));
  unsigned long long* out; cudaMallocManaged(&out, 148 * 4 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + NST * STAGE);
  const int tiles = 64;
  for (int rowb : {96, 272})
    for (int rows : {64, 16})
      for (int mode = 0; mode < 2; ++mode)
        for (int per_sm = 1; per_sm <= 3; ++per_sm) {
          k<<<148 * per_sm, 256, 128 + NST * STAGE>>>(src, span, tiles, rows, rowb, mode, out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          double s = 0; for (int i = 0; i < 148 * per_sm; ++i) s += out[i];
          printf("rowb %3d rows %2d mode %s ctas/sm %d: %.0f cycles per tile per CTA (%.1f per row)\n", rowb, rows, mode ? "ldgsts" : "bulk  ", per_sm,
                 s / (148 * per_sm) / tiles, s / (148 * per_sm) / tiles / rows);
        }
  return 0;
}
