// Shared helpers of the cvd_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/cvd_b200.h"

namespace cvd {

// thread-local message behind cvd_last_error()
char* error_buffer();
int fail(int code, const char* fmt, ...);

#define CVD_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::cvd::fail(CVD_E_INVALID, __VA_ARGS__); \
  } while (0)

#define CVD_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ::cvd::fail(CVD_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
  } while (0)

#define CVD_LAUNCH_CHECK(name)                                                           \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return ::cvd::fail(CVD_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// K / V^T operands of the attention kernels: one buffer, or one buffer per source GPU of a view-sharded run (the peers'
// symmetric-memory copies through their NVLink mappings).  Partner rows are numbered globally: row p lives in source
// p / rows_per_src at local row p % rows_per_src.
constexpr int kMaxKvSrc = 8;
struct KvSrc {
  const void* k[kMaxKvSrc];
  const void* v[kMaxKvSrc];
  int n_src;
  int rows_per_src;
};

// opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device (once per kernel and device)
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Pixel-centre coordinate of latent index i in F_mat_size space (epi_module.py:298):
//   x = (S/feat) * i + (S/feat - 1) / 2
__host__ __device__ __forceinline__ float pixel_centre(int i, float step) {
  return step * (float)i + (step - 1.0f) * 0.5f;
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, size_t i) { return __half2float(p[i]); }

template <typename T>
__device__ __forceinline__ void store_from_float(T* p, size_t i, float v);
template <>
__device__ __forceinline__ void store_from_float<float>(float* p, size_t i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, size_t i, float v) {
  p[i] = __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ void store_from_float<__half>(__half* p, size_t i, float v) { p[i] = __float2half_rn(v); }


// Programmatic dependent launch (griddepcontrol): a kernel launched with the attribute may begin while its predecessor in
// the stream is still running -- its CTAs take the slots the predecessor's tail leaves idle, run their prologue, and
// block in pdl_wait() until the predecessor has completed and its memory is visible.  Every kernel launched that way calls
// pdl_wait() before its first global-memory access; pdl_trigger() at the top of a kernel lets ITS successor start early.
// CVD_PDL=0 disables the attribute (plain stream order).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();
// cluster_x > 1: the grid is launched as thread-block clusters of cluster_x consecutive CTAs (gridDim.x % cluster_x == 0)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cluster_x > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_pdl_cluster(kernel, grid, block, smem, st, 1u, args...);
}

}  // namespace cvd
