// Pipe-throughput micro-benchmarks on B200: cycles per warp-instruction per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/ubench.cu && /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float min3(float a, float b, float c) { float y; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8]; uint64_t p[8]; uint32_t q[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = (uint64_t)threadIdx.x * 77 + i; q[i] = i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);
      if (MODE == 1) a[i] = min3(a[i], a[(i + 1) & 7], 0.5f);
      if (MODE == 2) p[i] = add2(p[i], p[(i + 1) & 7]);
      if (MODE == 3) p[i] = fma2(p[i], p[(i + 1) & 7], p[(i + 2) & 7]);
      if (MODE == 4) q[i] = packbf(a[i], __uint_as_float(q[i]));
      if (MODE == 5) a[i] = fmaf(a[i], 1.0001f, a[(i + 1) & 7]);
      if (MODE == 6) { a[i] = ex2(a[i]); p[i] = add2(p[i], p[(i + 1) & 7]); a[(i+3)&7] = min3(a[(i+3)&7], a[(i + 4) & 7], 0.5f); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + (float)p[i] + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  k<MODE><<<148, warps_per_sm * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  k<MODE><<<148, warps_per_sm * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double per_smsp = c / (double)(iters * 8) / (warps_per_sm / 4.0);   // cycles per warp-instr per SMSP
  printf("%-28s warps/SM=%2d  cycles per warp-instr per SMSP = %.2f\n", name, warps_per_sm, per_smsp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("MUFU.EX2", 4); run<1>("FMNMX3", 4); run<2>("FADD2", 4); run<3>("FFMA2", 4); run<4>("F2FP.BF16 pack", 4); run<5>("FFMA", 4); run<6>("mix ex2+add2+min3 (3 instr)", 4); }
    if (w == 8) { run<0>("MUFU.EX2", 8); run<1>("FMNMX3", 8); run<2>("FADD2", 8); run<3>("FFMA2", 8); run<4>("F2FP.BF16 pack", 8); run<5>("FFMA", 8); run<6>("mix ex2+add2+min3 (3 instr)", 8); }
    if (w == 16) { run<0>("MUFU.EX2", 16); run<1>("FMNMX3", 16); run<2>("FADD2", 16); run<3>("FFMA2", 16); run<4>("F2FP.BF16 pack", 16); run<5>("FFMA", 16); run<6>("mix ex2+add2+min3 (3 instr)", 16); }
  }
  return 0;
}
