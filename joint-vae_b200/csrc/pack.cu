// Weight re-packing: fp32 parameters in the reference's torch layouts -> every bf16 operand arrangement the tcgen05
// kernels read, for ALL layers of a stack in ONE launch (include/jvae_b200.h: jvae_pack_weights).
//
// The optimizer updates the flat fp32 parameter buffer in place every step, so the packed copies are rebuilt every step:
// this kernel is the whole cost of that (one read of the parameters, one write per arrangement).
#include "common.cuh"

namespace jvae {

constexpr int PK_THREADS = 256;
constexpr int PK_VEC = 8;                       // elements per thread (one 16-byte bf16 store)
constexpr int PK_PER_BLOCK = PK_THREADS * PK_VEC;

__global__ void __launch_bounds__(PK_THREADS) pack_weights_kernel(const jvae_pack_job* __restrict__ jobs, int n_jobs,
                                                                   const int32_t* __restrict__ taps) {
  // job of this block: the last job whose first_block <= blockIdx.x (jobs are sorted by first_block)
  __shared__ jvae_pack_job job;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    job = jobs[lo];
  }
  __syncthreads();
  const jvae_pack_job& J = job;
  const long long per_row = (long long)J.T * J.cols_pad;
  const long long total = (long long)J.rows_pad * per_row;
  const long long e0 = ((long long)(blockIdx.x - J.first_block) * PK_THREADS + threadIdx.x) * PK_VEC;
  if (e0 >= total) return;
  const int r = (int)(e0 / per_row);
  const long long rem = e0 - (long long)r * per_row;
  const int t = (int)(rem / J.cols_pad);
  const int c_first = (int)(rem - (long long)t * J.cols_pad);      // cols_pad % 8 == 0: the 8 elements share (r, t)
  float v[PK_VEC];
#pragma unroll
  for (int i = 0; i < PK_VEC; ++i) v[i] = 0.f;
  if (r < J.rows) {
    const int r1 = r / J.R0, r0 = r - r1 * J.R0;
    const float* base = J.src + (long long)r1 * J.s_r1 + (long long)r0 * J.s_r0 + (long long)taps[J.tap_off + t] * J.s_t;
#pragma unroll
    for (int i = 0; i < PK_VEC; ++i) {
      const int c = c_first + i;
      if (c < J.cols) {
        const int c1 = c / J.C0, c0 = c - c1 * J.C0;
        v[i] = __ldg(base + (long long)c1 * J.s_c1 + (long long)c0 * J.s_c0);
      }
    }
  }
  if (J.dst_f32) {
    float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(J.dst) + e0);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(J.dst) + e0) = o;
  }
}

}  // namespace jvae

using namespace jvae;

extern "C" {

int jvae_pack_job_blocks(long long rows_pad, int T, int cols_pad) {
  const long long total = rows_pad * (long long)T * cols_pad;
  return (int)((total + PK_PER_BLOCK - 1) / PK_PER_BLOCK);
}

int jvae_pack_weights(const jvae_pack_job* jobs_dev, int n_jobs, const int32_t* taps_dev, int total_blocks, void* stream) {
  JVAE_CHECK_ARG(jobs_dev && taps_dev, "job table and tap table are required");
  JVAE_CHECK_ARG(n_jobs > 0 && total_blocks > 0, "empty job table");
  pack_weights_kernel<<<total_blocks, PK_THREADS, 0, (cudaStream_t)stream>>>(jobs_dev, n_jobs, taps_dev);
  JVAE_LAUNCH_CHECK();
  return JVAE_OK;
}

}  // extern "C"
