// On-device self test of the tensor-core kernels against naive CUDA-core references computed from the same bf16
// operands.  Test utility only (allocates its own scratch); reports per-case max errors on stdout.
#include "tc_common.cuh"
#include <vector>

namespace jvae {

__global__ void fill_bf16_kernel(__nv_bfloat16* p, size_t n, uint32_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = __float2bfloat16(((float)(h & 0xffff) / 32768.f) - 1.f);
  }
}

// reference D[m,n] = act(sum_k A(m,k) B(n,k) + bias[n]) with explicit element strides
__global__ void ref_gemm_kernel(int M, int N, int K, const __nv_bfloat16* a, long a_sm, long a_sk, const __nv_bfloat16* b,
                                long b_sn, long b_sk, const float* bias, int act, float* out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc = fmaf(__bfloat162float(a[m * a_sm + k * a_sk]), __bfloat162float(b[n * b_sn + k * b_sk]), acc);
  if (bias) acc += bias[n];
  if (act == JVAE_ACT_RELU) acc = fmaxf(acc, 0.f);
  else if (act == JVAE_ACT_SIGMOID) acc = 1.f / (1.f + expf(-acc));
  else if (act == JVAE_ACT_LEAKY) acc = acc > 0.f ? acc : JVAE_LEAKY_SLOPE * acc;
  out[(size_t)m * N + n] = acc;
}

__global__ void max_err_kernel(const float* a, const float* b, size_t n, float* out) {
  float e = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float d = fabsf(a[i] - b[i]);
    if (!(d == d)) d = 1e30f;
    e = fmaxf(e, d);
  }
  e = warp_max(e);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(e));
}

static int round8(int v) { return (v + 7) & ~7; }

static int gemm_case(int mode, int M, int N, int K, int act, bool use_bias, int verbose) {
  const char* names[] = {"NT", "NN", "TN"};
  const bool a_mn = mode == JVAE_GEMM_TN, b_mn = mode != JVAE_GEMM_NT;
  const int lda = round8(a_mn ? M : K), ldb = round8(b_mn ? N : K);
  const size_t a_elems = (size_t)(a_mn ? K : M) * lda, b_elems = (size_t)(b_mn ? K : N) * ldb;
  __nv_bfloat16 *a = nullptr, *b = nullptr;
  float *bias = nullptr, *out = nullptr, *ref = nullptr, *err = nullptr;
  cudaMalloc(&a, a_elems * 2); cudaMalloc(&b, b_elems * 2);
  cudaMalloc(&bias, (size_t)N * 4); cudaMalloc(&out, (size_t)M * N * 4); cudaMalloc(&ref, (size_t)M * N * 4);
  cudaMalloc(&err, 4);
  fill_bf16_kernel<<<64, 256>>>(a, a_elems, 0x1234u + mode);
  fill_bf16_kernel<<<64, 256>>>(b, b_elems, 0x9876u + N);
  std::vector<float> hb(N);
  for (int i = 0; i < N; ++i) hb[i] = 0.01f * (float)((i * 37) % 101) - 0.5f;
  cudaMemcpy(bias, hb.data(), (size_t)N * 4, cudaMemcpyHostToDevice);
  cudaMemset(out, 0xff, (size_t)M * N * 4);
  cudaMemset(err, 0, 4);
  int rc = jvae_gemm_bf16(mode, M, N, K, a, lda, b, ldb, use_bias ? bias : nullptr, act, nullptr, out, N, nullptr, 0, nullptr);
  float h_err = -1.f;
  if (rc == 0) {
    dim3 g((N + 127) / 128, M);
    ref_gemm_kernel<<<g, 128>>>(M, N, K, a, a_mn ? 1 : lda, a_mn ? lda : 1, b, b_mn ? 1 : ldb, b_mn ? ldb : 1,
                                use_bias ? bias : nullptr, act, ref);
    max_err_kernel<<<64, 256>>>(out, ref, (size_t)M * N, err);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("[selftest] gemm %s M=%d N=%d K=%d: CUDA error %s\n", names[mode], M, N, K, cudaGetErrorString(e));
      rc = -2;
    } else {
      cudaMemcpy(&h_err, err, 4, cudaMemcpyDeviceToHost);
    }
  }
  const float tol = 2e-3f * sqrtf((float)K) + 1e-3f;
  const bool ok = rc == 0 && h_err >= 0.f && h_err <= tol;
  if (verbose || !ok)
    printf("[selftest] gemm %s M=%d N=%d K=%d act=%d bias=%d: rc=%d max_abs_err=%g (tol %g) %s%s\n", names[mode], M, N, K,
           act, (int)use_bias, rc, h_err, tol, ok ? "OK" : "FAIL", rc ? jvae_last_error() : "");
  cudaFree(a); cudaFree(b); cudaFree(bias); cudaFree(out); cudaFree(ref); cudaFree(err);
  return ok ? 0 : 1;
}

int conv_selftest(int verbose);  // conv.cu

}  // namespace jvae

using namespace jvae;

extern "C" int jvae_selftest(int verbose) {
  int fails = 0;
  int sms = 0, maj = 0, min = 0;
  if (jvae_device_info(0, &sms, &maj, &min) != 0) {
    printf("[selftest] %s\n", jvae_last_error());
    return 1;
  }
  if (verbose) printf("[selftest] device 0: %d SMs, sm_%d%d\n", sms, maj, min);
  const int shapes[][3] = {{128, 128, 64}, {128, 64, 64}, {256, 128, 256}, {128, 128, 512}, {300, 200, 136},
                           {1000, 16, 128}, {64, 784, 512}, {8704, 4096, 128}};
  for (int mode = 0; mode < 3; ++mode)
    for (auto& s : shapes) {
      fails += gemm_case(mode, s[0], s[1], s[2], JVAE_ACT_NONE, false, verbose);
      if (fails > 6) {
        printf("[selftest] too many failures, stopping\n");
        return fails;
      }
    }
  fails += gemm_case(JVAE_GEMM_NT, 256, 256, 128, JVAE_ACT_RELU, true, verbose);
  fails += gemm_case(JVAE_GEMM_NT, 130, 72, 64, JVAE_ACT_SIGMOID, true, verbose);
  fails += gemm_case(JVAE_GEMM_NT, 200, 48, 96, JVAE_ACT_LEAKY, true, verbose);
  fails += conv_selftest(verbose);
  if (verbose) printf("[selftest] %d failure(s)\n", fails);
  return fails;
}
