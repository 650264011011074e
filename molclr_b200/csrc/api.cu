// Library-level entry points: error reporting, device info.
#include <cstdarg>
#include <cstring>

#include "common.cuh"
#include "molclr_b200.h"

namespace molclr {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
int g_pdl = 1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return -1;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace molclr

extern "C" int molclr_abi_version(void) { return MOLCLR_ABI_VERSION; }
extern "C" const char* molclr_last_error(void) { return molclr::g_err; }
extern "C" uint64_t molclr_launch_count(void) { return molclr::g_launches; }
extern "C" int molclr_set_pdl(int enable) {
  const int before = molclr::g_pdl;
  if (enable >= 0) molclr::g_pdl = enable != 0;
  return before;
}

extern "C" int molclr_device_info(int* sm_count_out, int* cc) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return molclr::cuda_fail(e, "device_info");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (sm_count_out) *sm_count_out = molclr::sm_count();
  if (cc) *cc = major * 10 + minor;
  return 0;
}
