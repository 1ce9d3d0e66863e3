// Which hardware warp slot (%warpid) and SM does each warp of a 160-thread CTA get when three such CTAs share an SM
// (the launch shape of epi_attn_fwd_kernel<40,32>)?  Scheduler = %warpid % 4: shows whether the fifth warp (TMA producer +
// MMA issuer) of every CTA lands on the same scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ubench_warp_slots.bin ubench_warp_slots.cu && ./ubench_warp_slots.bin
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(160, 3) k(int* out, int spin) {
  extern __shared__ unsigned char sm[];
  unsigned smid, warpid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if ((threadIdx.x & 31) == 0) {
    int w = threadIdx.x >> 5;
    out[(blockIdx.x * 5 + w) * 2] = smid;
    out[(blockIdx.x * 5 + w) * 2 + 1] = warpid;
  }
  if (threadIdx.x == 0) sm[0] = 1;
}
int main() {
  const int grid = 444 * 3;
  int* d; cudaMalloc(&d, grid * 5 * 2 * sizeof(int));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  k<<<grid, 160, 60 * 1024>>>(d, 200000);
  cudaDeviceSynchronize();
  int* h = new int[grid * 5 * 2];
  cudaMemcpy(h, d, grid * 5 * 2 * sizeof(int), cudaMemcpyDeviceToHost);
  int hist[5][4] = {};
  for (int b = 0; b < grid; ++b) for (int w = 0; w < 5; ++w) hist[w][h[(b * 5 + w) * 2 + 1] & 3]++;
  for (int w = 0; w < 5; ++w) printf("CTA warp %d -> scheduler (%%warpid %% 4) histogram: %d %d %d %d\n", w, hist[w][0], hist[w][1], hist[w][2], hist[w][3]);
  printf("first CTAs: ");
  for (int b = 0; b < 6; ++b) { printf("[sm %d:", h[b * 10]); for (int w = 0; w < 5; ++w) printf(" %d", h[(b * 5 + w) * 2 + 1]); printf("] "); }
  printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
