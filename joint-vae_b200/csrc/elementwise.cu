// Casts and layout changes at the boundary between the reference's NCHW f32 tensors and the
// NHWC bf16 activations the tcgen05 kernels consume.
#include "common.cuh"

namespace jvae {

constexpr int EW_THREADS = 256;

__global__ void __launch_bounds__(EW_THREADS) cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, size_t n) {
  const size_t n8 = n >> 3;
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n8; i += (size_t)gridDim.x * EW_THREADS) {
    const float4 a = reinterpret_cast<const float4*>(s)[2 * i], b = reinterpret_cast<const float4*>(s)[2 * i + 1];
    uint4 o;
    o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
    reinterpret_cast<uint4*>(d)[i] = o;
  }
  if (blockIdx.x == 0)
    for (size_t i = (n8 << 3) + threadIdx.x; i < n; i += EW_THREADS) d[i] = __float2bfloat16(s[i]);
}

__global__ void __launch_bounds__(EW_THREADS) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d, size_t n) {
  const size_t n8 = n >> 3;
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n8; i += (size_t)gridDim.x * EW_THREADS) {
    const uint4 u = reinterpret_cast<const uint4*>(s)[i];
    reinterpret_cast<float4*>(d)[2 * i] = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    reinterpret_cast<float4*>(d)[2 * i + 1] = make_float4(bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w));
  }
  if (blockIdx.x == 0)
    for (size_t i = (n8 << 3) + threadIdx.x; i < n; i += EW_THREADS) d[i] = __bfloat162float(s[i]);
}

// one thread per (image, pixel): channel reads are coalesced across the warp, the c_pad bf16 of a pixel
// are written contiguously
__global__ void __launch_bounds__(EW_THREADS) nchw_to_nhwc_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d,
                                                                   int n, int c, int hw, int c_pad) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * EW_THREADS) {
    const size_t img = i / hw, p = i - img * hw;
    const float* sp = s + img * (size_t)c * hw + p;
    __nv_bfloat16* dp = d + i * c_pad;
    int ch = 0;
    for (; ch < c; ++ch) dp[ch] = __float2bfloat16(sp[(size_t)ch * hw]);
    for (; ch < c_pad; ++ch) dp[ch] = __float2bfloat16(0.f);
  }
}

__global__ void __launch_bounds__(EW_THREADS) nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d,
                                                                   int n, int c, int hw, int c_pad) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * EW_THREADS) {
    const size_t img = i / hw, p = i - img * hw;
    const __nv_bfloat16* sp = s + i * c_pad;
    float* dp = d + img * (size_t)c * hw + p;
    for (int ch = 0; ch < c; ++ch) dp[(size_t)ch * hw] = __bfloat162float(sp[ch]);
  }
}

static int ew_grid(size_t items) {
  const size_t want = (items + EW_THREADS - 1) / EW_THREADS;
  const size_t cap = (size_t)sm_count() * 8;
  size_t g = want < cap ? want : cap;
  return (int)(g ? g : 1);
}

}  // namespace jvae

using namespace jvae;

extern "C" {

int jvae_cast_f32_bf16(const float* src, void* dst, size_t n, void* stream) {
  JVAE_CHECK_ARG(src && dst, "src and dst are required");
  JVAE_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "buffers must be 16-byte aligned");
  if (n == 0) return JVAE_OK;
  cast_f32_bf16_kernel<<<ew_grid(n / 8 + 1), EW_THREADS, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  JVAE_LAUNCH_CHECK();
  return JVAE_OK;
}

int jvae_cast_bf16_f32(const void* src, float* dst, size_t n, void* stream) {
  JVAE_CHECK_ARG(src && dst, "src and dst are required");
  JVAE_CHECK_ARG((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "buffers must be 16-byte aligned");
  if (n == 0) return JVAE_OK;
  cast_bf16_f32_kernel<<<ew_grid(n / 8 + 1), EW_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  JVAE_LAUNCH_CHECK();
  return JVAE_OK;
}

int jvae_nchw_to_nhwc_bf16(const float* src, void* dst, int n, int c, int h, int w, int c_pad, void* stream) {
  JVAE_CHECK_ARG(src && dst, "src and dst are required");
  JVAE_CHECK_ARG(n > 0 && c > 0 && h > 0 && w > 0 && c_pad >= c, "bad dims");
  nchw_to_nhwc_kernel<<<ew_grid((size_t)n * h * w), EW_THREADS, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n, c, h * w, c_pad);
  JVAE_LAUNCH_CHECK();
  return JVAE_OK;
}

int jvae_nhwc_bf16_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int c_pad, void* stream) {
  JVAE_CHECK_ARG(src && dst, "src and dst are required");
  JVAE_CHECK_ARG(n > 0 && c > 0 && h > 0 && w > 0 && c_pad >= c, "bad dims");
  nhwc_to_nchw_kernel<<<ew_grid((size_t)n * h * w), EW_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), dst, n, c, h * w, c_pad);
  JVAE_LAUNCH_CHECK();
  return JVAE_OK;
}

}  // extern "C"
