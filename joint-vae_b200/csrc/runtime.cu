// Error plumbing, launch accounting and device queries of libjvae_sm100.so.
#include "common.cuh"
#include <stdarg.h>
#include <mutex>
#include <vector>

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

// ---------------------------------------------------------------- per-launch profile (bench.py's roofline figures)
// When enabled, the entry points that call prof_begin / prof_end bracket their main kernel launch with CUDA events recorded on
// the launching stream from inside the library: nothing of the host-side call path (argument marshalling, Python) sits between
// the first event and the kernel.  Not capture-safe: enable only around eager launches.
struct ProfRec { int tag; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};

void* prof_begin(int tag, cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return nullptr;
  ProfRec r;
  r.tag = tag;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return nullptr;
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  return (void*)(uintptr_t)g_prof.size();      // 1-based index of the record
}
void prof_end(void* h, cudaStream_t st) {
  if (!h) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[(size_t)(uintptr_t)h - 1].b, st);
}

}  // namespace jvae

extern "C" {

int jvae_profile_enable(int on) {
  jvae::g_prof_on.store(on ? 1 : 0);
  return JVAE_OK;
}

int jvae_profile_drain(int32_t* tags, float* ms, int max_records) {
  std::lock_guard<std::mutex> lk(jvae::g_prof_mu);
  int n = 0;
  for (auto& r : jvae::g_prof) {
    float t = -1.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess) cudaEventElapsedTime(&t, r.a, r.b);
    if (n < max_records && tags && ms) { tags[n] = r.tag; ms[n] = t; ++n; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  jvae::g_prof.clear();
  return n;
}

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
