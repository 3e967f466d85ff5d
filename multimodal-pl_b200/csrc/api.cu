// Library-wide state: error string, launch counter, device check.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mmpl {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
}  // namespace mmpl

extern "C" {
int mmpl_version(void) { return 100; }
const char* mmpl_last_error(void) { return mmpl::g_err; }
uint64_t mmpl_launch_count(void) { return mmpl::g_launches.load(); }
int mmpl_check_device(void) {
  int dev = 0, major = 0;
  MMPL_CUDA(cudaGetDevice(&dev));
  MMPL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MMPL_REQUIRE(major == 10, MMPL_E_ARCH, "device compute capability %d.x is not sm_100-class; no fallback path", major);
  return MMPL_OK;
}
}
