// Shared helpers for libjvae_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/jvae_b200.h"

namespace jvae {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define JVAE_CHECK_ARG(cond, msg)                                                   \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      jvae::set_error("%s: invalid argument: %s (%s)", __func__, msg, #cond);       \
      return JVAE_ERR_INVALID;                                                      \
    }                                                                               \
  } while (0)

#define JVAE_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      jvae::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(_e));\
      return JVAE_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

#define JVAE_LAUNCH_CHECK()                                                         \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      jvae::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(_e)); \
      return JVAE_ERR_CUDA;                                                         \
    }                                                                               \
    jvae::count_launch();                                                           \
  } while (0)

int sm_count();   // SMs of the current device (cached)

// per-launch profile (runtime.cu): events recorded inside the library right around a kernel launch when enabled
void* prof_begin(int tag, cudaStream_t st);
void prof_end(void* h, cudaStream_t st);

// ---------------------------------------------------------------- small device utilities
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

// block-wide sum; `red` is >= 32 floats of shared memory; result valid in all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// streaming 16-byte load that does not pollute L1
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace jvae
