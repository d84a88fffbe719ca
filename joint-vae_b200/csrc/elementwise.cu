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

// Patch matrix of a convolution on a SMALL map (weight gradient of the 512-channel vgg19 layers at 4 x 4 and 2 x 2): row
// q = (n, y, x) of `out` holds column ci * T + t = X[n, y * s + dy_t, x * s + dx_t, ci] (zero outside the map), i.e. the
// columns follow torch's (Cin, kh, kw) weight order, so dW = dY^T . out is the weight gradient in torch's layout (one TN GEMM
// accumulating straight into .grad).  One thread per (row, 8 channels): T 16-byte loads, a register transpose, T 16-byte stores.
struct TapList { short dy[32], dx[32]; };
struct TapList64 { short dy[64], dx[64]; };
template <int T>
__global__ void __launch_bounds__(EW_THREADS) im2col_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int ldx,
                                                            int Hq, int Wq, int stride, const TapList taps,
                                                            __nv_bfloat16* __restrict__ out, int ldo) {
  const int cg = C >> 3;
  const size_t total = (size_t)N * Hq * Wq * cg;
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * EW_THREADS) {
    const int g = (int)(i % cg);
    const size_t q = i / cg;
    const int xq = (int)(q % Wq), yq = (int)((q / Wq) % Hq), n = (int)(q / ((size_t)Wq * Hq));
    __nv_bfloat16 v[8 * T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int yy = yq * stride + taps.dy[t], xx = xq * stride + taps.dx[t];
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        u = *reinterpret_cast<const uint4*>(x + (((size_t)n * H + yy) * W + xx) * ldx + (size_t)g * 8);
      const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&u);
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c * T + t] = e[c];
    }
    uint4* o = reinterpret_cast<uint4*>(out + q * (size_t)ldo + (size_t)g * 8 * T);
#pragma unroll
    for (int t = 0; t < T; ++t) o[t] = reinterpret_cast<const uint4*>(v)[t];
  }
}

// the same patch matrix with the columns in (tap, channel) order: column t * C + ci; any window size, one 16-byte store per tap
__global__ void __launch_bounds__(EW_THREADS) im2col_tapmajor_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                                                     int ldx, int Hq, int Wq, int stride, int T, const TapList64 taps,
                                                                     __nv_bfloat16* __restrict__ out, int ldo) {
  const int cg = C >> 3;
  const size_t total = (size_t)N * Hq * Wq * cg;
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * EW_THREADS) {
    const int g = (int)(i % cg);
    const size_t q = i / cg;
    const int xq = (int)(q % Wq), yq = (int)((q / Wq) % Hq), n = (int)(q / ((size_t)Wq * Hq));
    uint4* o = reinterpret_cast<uint4*>(out + q * (size_t)ldo + (size_t)g * 8);
    for (int t = 0; t < T; ++t) {
      const int yy = yq * stride + taps.dy[t], xx = xq * stride + taps.dx[t];
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        u = *reinterpret_cast<const uint4*>(x + (((size_t)n * H + yy) * W + xx) * ldx + (size_t)g * 8);
      o[(size_t)t * cg] = u;
    }
  }
}

// gather + augmentation + ToTensor of one batch (include/jvae_b200.h: jvae_batch_u8_to_f32).  One thread per output
// element, x fastest: coalesced f32 stores; the uint8 source rows of an image (<= a few KB) stay in L1 / L2.
struct BatchArgs {
  const uint8_t* src; const long long* index; const unsigned char* flip; const int* crop_ij; float* out;
  int B, H, W, C, oH, oW, crop_pad, flip_first, off_y, off_x;
};

__global__ void __launch_bounds__(EW_THREADS) batch_u8_to_f32_kernel(const BatchArgs a) {
  const size_t per_img = (size_t)a.C * a.oH * a.oW, total = per_img * a.B;
  for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * EW_THREADS) {
    const int b = (int)(i / per_img);
    size_t r = i - (size_t)b * per_img;
    const int c = (int)(r / ((size_t)a.oH * a.oW));
    r -= (size_t)c * a.oH * a.oW;
    const int oy = (int)(r / a.oW), ox = (int)(r - (size_t)oy * a.oW);
    int y = oy + a.off_y, x = ox + a.off_x;          // position in the augmented (H, W) image
    float v = 0.f;
    if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
      const bool fl = a.flip && a.flip[b];
      if (fl && !a.flip_first) x = a.W - 1 - x;      // crop, then flip: the flip acts on the cropped image
      if (a.crop_ij) {                               // padded(y + i, x + j) with edge replication
        y = min(max(y + a.crop_ij[2 * b] - a.crop_pad, 0), a.H - 1);
        x = min(max(x + a.crop_ij[2 * b + 1] - a.crop_pad, 0), a.W - 1);
      }
      if (fl && a.flip_first) x = a.W - 1 - x;       // flip, then crop: the crop reads the flipped image
      const uint8_t u = a.src[(((size_t)a.index[b] * a.H + y) * a.W + x) * a.C + c];
      v = __fdiv_rn((float)u, 255.f);                // ToTensor: byte tensor .div(255) in f32
    }
    a.out[i] = v;
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

int jvae_batch_u8_to_f32(const jvae_batch_cfg* cfg, const void* src, long long n_src, const long long* index, int B,
                         const unsigned char* flip, const int* crop_ij, float* out, void* stream) {
  JVAE_CHECK_ARG(cfg && src && index && out, "cfg, src, index and out are required");
  JVAE_CHECK_ARG(B > 0 && n_src > 0, "empty batch or dataset");
  JVAE_CHECK_ARG(cfg->H > 0 && cfg->W > 0 && cfg->C > 0 && cfg->out_H > 0 && cfg->out_W > 0, "bad image dims");
  JVAE_CHECK_ARG(cfg->crop_pad >= 0 && (cfg->crop_pad > 0 || !crop_ij), "crop offsets given without RandomCrop padding");
  BatchArgs a;
  a.src = reinterpret_cast<const uint8_t*>(src); a.index = index; a.flip = flip; a.crop_ij = crop_ij; a.out = out;
  a.B = B; a.H = cfg->H; a.W = cfg->W; a.C = cfg->C; a.oH = cfg->out_H; a.oW = cfg->out_W;
  a.crop_pad = cfg->crop_pad; a.flip_first = cfg->flip_first; a.off_y = cfg->post_off_y; a.off_x = cfg->post_off_x;
  batch_u8_to_f32_kernel<<<ew_grid((size_t)B * a.C * a.oH * a.oW), EW_THREADS, 0, (cudaStream_t)stream>>>(a);
  JVAE_LAUNCH_CHECK();
  return JVAE_OK;
}

int jvae_im2col_bf16(const void* x, int N, int H, int W, int C, int ld_x, int ntaps, const int16_t* tap_dy, const int16_t* tap_dx,
                     int in_stride, int Hq, int Wq, void* out, int ld_out, int tap_major, void* stream) {
  JVAE_CHECK_ARG(x && out && tap_dy && tap_dx, "x, out and the tap lists are required");
  JVAE_CHECK_ARG(N > 0 && H > 0 && W > 0 && Hq > 0 && Wq > 0 && in_stride > 0, "bad dims");
  JVAE_CHECK_ARG(C > 0 && (C % 8) == 0 && (ld_x % 8) == 0 && ld_x >= C, "C and ld_x must be multiples of 8");
  JVAE_CHECK_ARG(ld_out >= C * ntaps && (ld_out % 8) == 0, "ld_out must cover C * ntaps columns and be a multiple of 8");
  JVAE_CHECK_ARG((((uintptr_t)x | (uintptr_t)out) & 15) == 0, "buffers must be 16-byte aligned");
  if (tap_major) {
    JVAE_CHECK_ARG(ntaps >= 1 && ntaps <= 64, "1..64 taps");
    TapList64 tl;
    for (int t = 0; t < ntaps; ++t) { tl.dy[t] = tap_dy[t]; tl.dx[t] = tap_dx[t]; }
    const size_t total = (size_t)N * Hq * Wq * (C / 8);
    im2col_tapmajor_kernel<<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), N, H, W, C, ld_x, Hq, Wq, in_stride, ntaps, tl,
        reinterpret_cast<__nv_bfloat16*>(out), ld_out);
    JVAE_LAUNCH_CHECK();
    return JVAE_OK;
  }
  if (ntaps != 9 && ntaps != 25 && ntaps != 1 && ntaps != 4) {
    set_error("jvae_im2col_bf16: windows of 1, 4, 9 or 25 taps only (ntaps=%d)", ntaps);
    return JVAE_ERR_UNSUPPORTED;
  }
  TapList tl;
  for (int t = 0; t < ntaps; ++t) { tl.dy[t] = tap_dy[t]; tl.dx[t] = tap_dx[t]; }
  const size_t total = (size_t)N * Hq * Wq * (C / 8);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
  cudaStream_t st = (cudaStream_t)stream;
  switch (ntaps) {
    case 1: im2col_kernel<1><<<ew_grid(total), EW_THREADS, 0, st>>>(xp, N, H, W, C, ld_x, Hq, Wq, in_stride, tl, op, ld_out); break;
    case 4: im2col_kernel<4><<<ew_grid(total), EW_THREADS, 0, st>>>(xp, N, H, W, C, ld_x, Hq, Wq, in_stride, tl, op, ld_out); break;
    case 9: im2col_kernel<9><<<ew_grid(total), EW_THREADS, 0, st>>>(xp, N, H, W, C, ld_x, Hq, Wq, in_stride, tl, op, ld_out); break;
    default: im2col_kernel<25><<<ew_grid(total), EW_THREADS, 0, st>>>(xp, N, H, W, C, ld_x, Hq, Wq, in_stride, tl, op, ld_out); break;
  }
  JVAE_LAUNCH_CHECK();
  return JVAE_OK;
}

}  // extern "C"
