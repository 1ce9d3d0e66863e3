// Issue cost of tcgen05.mma / tcgen05.commit from one elected thread (B200): cycles per loop
// iteration for different mixes, 1 or 2 CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cvd_b200/csrc -o scripts/ubench_mma_issue.bin scripts/ubench_mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace cvd::ptx;

template <int MODE>
__global__ void __launch_bounds__(64, 2) k(long long* cyc, int iters) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(smem_u32(&tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idQK = umma_idesc_bf16(128, 64), idPV = umma_idesc_bf16(128, 48);
  if (warp == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it & 3;
      if (elect_one()) {
        if (MODE != 3) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            umma_ss(tmem + (it & 1) * 64, umma_desc_k_sw128(base + ks * 32), umma_desc_k_sw128(base + 16384 + st * 8192 + ks * 32), idQK, ks > 0);
        }
        if (MODE == 1 || MODE == 4 || MODE == 5) umma_commit(smem_u32(&bars[0]));
        if (MODE == 2 || MODE == 4 || MODE == 5) { umma_commit(smem_u32(&bars[1])); }
        if (MODE == 2) umma_commit(smem_u32(&bars[0]));
        if (MODE == 3 || MODE == 4 || MODE == 5) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ts(tmem + 128, tmem + 192 + (it & 1) * 32 + ks * 8, umma_desc_k_sw128(base + 16384 + st * 6144 + ks * 32), idPV, 1);
        }
        if (MODE == 4 || MODE == 5) { umma_commit(smem_u32(&bars[2])); umma_commit(smem_u32(&bars[3])); }
      }
      __syncwarp();
      if (MODE == 5) {   // wait for the commit of the previous iteration (keeps at most ~2 iterations in flight)
        if (it > 0) mbar_wait(smem_u32(&bars[4]), (it - 1) & 1);
        if (elect_one()) umma_commit(smem_u32(&bars[4]));
        __syncwarp();
      }
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(smem_u32(&bars[7]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[7]), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}
template <int MODE> void run(const char* name, int grid) {
  long long* cyc; cudaMalloc(&cyc, 296 * 16);
  const int iters = 2000, smem = 49 * 1024 + 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int r = 0; r < 2; ++r) { k<MODE><<<grid, 64, smem>>>(cyc, iters); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; } }
  long long h[592]; cudaMemcpy(h, cyc, grid * 16, cudaMemcpyDeviceToHost);
  double a = 0, b = 0; for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
  printf("%-46s grid=%3d  issue %.0f cycles/iter, issue+drain %.0f cycles/iter\n", name, grid, a / grid / iters, b / grid / iters);
  cudaFree(cyc);
}
int main() {
  for (int grid : {148, 296}) {
    run<0>("3 MMA ss (128x64x16)", grid);
    run<1>("3 MMA ss + 1 commit", grid);
    run<2>("3 MMA ss + 2 commits", grid);
    run<3>("4 MMA ts (128x48x16)", grid);
    run<4>("3 ss + 2 commits + 4 ts + 2 commits", grid);
    run<5>("same + wait for previous iteration's commit", grid);
  }
  return 0;
}
