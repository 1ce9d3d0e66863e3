// Micro-benchmark of the softmax inner loop in isolation (registers only, no TMEM / barriers), as
// written in csrc/epi_attn_sm100.cu: per 32-key chunk, the scores of chunk k (FADD2 x2, FMNMX3 x2,
// FFMA2, FMNMX3 max per key pair) are interleaved with the exponentials of chunk k-1 (MUFU.EX2 x2,
// F2FP per pair), followed by the warp vote of the lazy-rescale check.
//   MODE 0: as in the kernel      MODE 1: without the running max / vote      MODE 2: exp only (MUFU + F2FP)
//   MODE 3: 0 + tcgen05.st x32 + wait::st per tile   MODE 4: 0 + two tcgen05.ld x32 per tile   MODE 5: both
// Prints cycles per warp-column per SM sub-partition for 1-4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cvd_b200/csrc -o scripts/ubench_softmax.bin scripts/ubench_softmax.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace cvd::ptx;

template <int MODE>
__global__ void k(const float* in, uint32_t* out, long long* cyc, int iters, float As, float bm, float negm) {
  uint32_t xa[32], xb[32], preg[32];
  for (int i = 0; i < 32; ++i) { xa[i] = __float_as_uint(in[threadIdx.x * 32 + i]); xb[i] = xa[i]; }
  uint32_t acc = 0;
  const uint64_t sc = pack2(0.228f, 0.228f), st1 = pack2(-2 * As, -2 * As), st2 = pack2(2 * As, 2 * As);
  auto fused = [&](uint32_t* xs, const uint32_t* xo, uint32_t* po, float t0) -> float {
    float xmax = -INFINITY;
    uint64_t a1 = pack2(bm - t0, bm - t0 - As), a2 = pack2(bm + t0, bm + t0 + As);
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      if (MODE != 2) {
        float a1x, a1y, a2x, a2y; unpack2(a1, a1x, a1y); unpack2(a2, a2x, a2y);
        const float v0 = min3(a1x, a2x, negm), v1 = min3(a1y, a2y, negm);
        float x0, x1; unpack2(fma2(pack2(__uint_as_float(xs[c]), __uint_as_float(xs[c + 1])), sc, pack2(v0, v1)), x0, x1);
        if (MODE != 1) xmax = max3(xmax, x0, x1);
        xs[c] = __float_as_uint(x0); xs[c + 1] = __float_as_uint(x1);
        a1 = add2(a1, st1); a2 = add2(a2, st2);
      }
      po[c >> 1] = pack_bf16(ex2(__uint_as_float(xo[c])), ex2(__uint_as_float(xo[c + 1])));
    }
    return xmax;
  };
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&tmem_slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot + ((uint32_t)((threadIdx.x >> 5) & 3) * 32 << 16) + (threadIdx.x >> 7) * 64;   // lane quarter of this warp, 64 columns per warpgroup
  constexpr bool kSt = (MODE == 3 || MODE == 5), kLd = (MODE == 4 || MODE == 5);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (kLd) { tmem_ld_x32(tmem, xa); tmem_ld_wait(); }
    float xm = fused(xa, xb, preg + 16, (float)it);
    if (MODE != 1 && MODE != 2 && __any_sync(0xffffffffu, xm > 1e30f)) acc += 1;
    if (kSt) { tmem_st_x32(tmem + 32, preg); tmem_st_wait(); }
    if (kLd) { tmem_ld_x32(tmem, xb); tmem_ld_wait(); }
    xm = fused(xb, xa, preg, (float)it + 0.5f);
    if (MODE != 1 && MODE != 2 && __any_sync(0xffffffffu, xm > 1e30f)) acc += 1;
#pragma unroll
    for (int c = 0; c < 32; ++c) acc ^= preg[c];
    if (MODE == 2) {
#pragma unroll
      for (int c = 0; c < 32; ++c) { xa[c] += acc & 1; xb[c] += acc & 1; }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem_slot); }
}
template <int MODE> void run(int warps) {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 148 * 1024 * 32 * 4); cudaMemset(in, 0, 148 * 1024 * 32 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  for (int r = 0; r < 2; ++r) { k<MODE><<<148, warps * 32>>>(in, out, cyc, iters, 0.37f, 4.3f, -1.0f); cudaDeviceSynchronize(); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("mode %d  warps/SM=%2d (%d per SMSP): %.2f cycles per warp-column per SMSP  (per warp: %.0f cycles per 64-key tile)\n", MODE, warps, warps / 4,
         c / (iters * 64.0) / (warps / 4.0), c / iters);
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}
int main() { for (int w : {4, 8}) { run<0>(w); run<3>(w); run<4>(w); run<5>(w); } return 0; }
