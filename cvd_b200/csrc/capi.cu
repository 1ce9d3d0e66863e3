// Error reporting and versioning of the C ABI (include/cvd_b200.h).
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace cvd {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CVD_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember the largest value set for every
// (kernel, device) so that a process driving several GPUs configures each of them (grow-only, thread-safe).
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[std::make_pair(kernel, dev)];
  if (bytes > have) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    have = bytes;
  }
  return cudaSuccess;
}

}  // namespace cvd

extern "C" const char* cvd_last_error(void) { return cvd::error_buffer(); }
extern "C" int cvd_abi_version(void) { return 1; }
