// Error reporting and versioning of the C ABI (include/cvd_b200.h).
#include <cstdlib>

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

}  // namespace cvd

extern "C" const char* cvd_last_error(void) { return cvd::error_buffer(); }
extern "C" int cvd_abi_version(void) { return 1; }
