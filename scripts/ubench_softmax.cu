// Micro-benchmark of the softmax inner loop in isolation (registers only, no TMEM / barriers):
// [score phase: FADD2 x2, FMNMX3 x2(+max), FFMA2] then [exp phase: MUFU.EX2 x2, F2FP] per key pair,
// 32 keys per chunk, with 1-4 warps per SM sub-partition.  Prints cycles per warp-column.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float min3(float a, float b, float c) { float y; asm("min.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float y; asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { uint32_t y; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

template <int MODE>   // 0: phases sequential (as in the kernel)  1: exp of the previous chunk interleaved with the score of this one
__global__ void k(const float* in, uint32_t* out, long long* cyc, int iters, float As, float bm, float negm) {
  float s[32], xp[32];
  for (int i = 0; i < 32; ++i) { s[i] = in[threadIdx.x * 32 + i]; xp[i] = s[i] * 0.5f; }
  uint32_t acc = 0; float mx = -1e30f;
  const uint64_t sc = pack2(0.228f, 0.228f), st1 = pack2(-2 * As, -2 * As), st2 = pack2(2 * As, 2 * As);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint64_t a1 = pack2(bm - it, bm - it - As), a2 = pack2(bm + it, bm + it + As);
    float x[32];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float a1x, a1y, a2x, a2y; unpack2(a1, a1x, a1y); unpack2(a2, a2x, a2y);
      float v0 = min3(a1x, a2x, negm), v1 = min3(a1y, a2y, negm);
      float x0, x1; unpack2(fma2(pack2(s[c], s[c + 1]), sc, pack2(v0, v1)), x0, x1);
      mx = max3(mx, x0, x1);
      x[c] = x0; x[c + 1] = x1;
      a1 = add2(a1, st1); a2 = add2(a2, st2);
      if (MODE == 1) acc ^= packbf(ex2(xp[c]), ex2(xp[c + 1]));
    }
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 32; c += 2) acc ^= packbf(ex2(x[c]), ex2(x[c + 1]));
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) xp[c] = x[c];
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) s[c] += 1e-3f * (float)(acc & 1);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(mx);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int warps) {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 148 * 1024 * 32 * 4); cudaMemset(in, 0, 148 * 1024 * 32 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  for (int r = 0; r < 2; ++r) { k<MODE><<<148, warps * 32>>>(in, out, cyc, iters, 0.37f, 4.3f, -1.0f); cudaDeviceSynchronize(); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("mode %d  warps/SM=%2d (%d per SMSP): %.2f cycles per warp-column per SMSP  (per warp: %.1f cycles per 32-key chunk)\n", MODE, warps, warps / 4,
         c / (iters * 32.0) / (warps / 4.0), c / iters);
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}
int main() { for (int w : {4, 8, 12, 16}) { run<0>(w); run<1>(w); } return 0; }
