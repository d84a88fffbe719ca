// Error plumbing, launch accounting and device queries of libjvae_sm100.so.
#include "common.cuh"
#include <stdarg.h>

namespace jvae {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace jvae

extern "C" {

const char* jvae_last_error(void) { return jvae::g_err; }

int jvae_abi_version(void) { return JVAE_ABI_VERSION; }

int64_t jvae_launch_count(void) { return (int64_t)jvae::g_launches.load(); }

int jvae_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n) {
    jvae::set_error("jvae_device_info: no CUDA device %d (%s)", device, cudaGetErrorString(e));
    return JVAE_ERR_NOGPU;
  }
  int sms = 0, maj = 0, min = 0;
  JVAE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  JVAE_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device));
  JVAE_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, device));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) {
    jvae::set_error("jvae_device_info: device %d is sm_%d%d, this library is built for sm_100a only", device, maj, min);
    return JVAE_ERR_NOGPU;
  }
  return JVAE_OK;
}

}  // extern "C"
