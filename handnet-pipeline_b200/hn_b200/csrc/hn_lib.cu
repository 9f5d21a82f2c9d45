// Library-level entry points: error text, device info, launch accounting.
#include "hn_common.cuh"

#include <atomic>
#include <stdarg.h>

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
int g_sms = 0;
}  // namespace

void hn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void hn_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int hn_num_sms() {
  if (g_sms == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      g_sms = sms;
    else
      return 148;
  }
  return g_sms;
}

extern "C" const char* hn_last_error(void) { return g_err; }
extern "C" int hn_version(void) { return 100; }
extern "C" int64_t hn_launch_count(void) { return g_launches.load(); }

extern "C" int hn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  HN_CHECK_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  HN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  HN_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  HN_CHECK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) {
    hn_set_error("libhandnet_b200 needs an sm_100 device, found sm_%d%d", maj, min);
    return HN_ERR_UNSUPPORTED;
  }
  return HN_OK;
}
