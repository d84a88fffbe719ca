// Hardware probe (diagnostic, not on the product path): which start addresses does a K-major swizzled UMMA shared-memory
// descriptor accept?  The halo-tile convolution kernel wants to read the SAME TMA-loaded NHWC tile once per filter tap
// through descriptors whose start address is shifted by whole pixel rows (64 B or 128 B) and whose 8-row groups are a
// halo-row apart (SBO != 8 * row pitch).  This probe runs D = A_shifted * Sel^T for every (swizzle, shift, group stride,
// base-offset policy) and counts mismatches against the expected gather, so the kernel design rests on measured
// behaviour instead of a reading of the PTX manual.
#include "tc_common.cuh"
#include <vector>

namespace jvae {

struct ProbeCfg { int shift, group_rows, policy; };   // policy 0: base_offset = 0; 1: base_offset = (start >> 7) & 7

__global__ void __launch_bounds__(128, 1)
probe_desc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int cblk,
                  const ProbeCfg* cfgs, int ncfg, int* mismatches) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t row_bytes = (uint32_t)cblk * 2u;
  uint8_t* sa = smem;                               // 512 rows
  uint8_t* sb = smem + 512u * row_bytes;            // 16 rows (1024-aligned)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 16u * row_bytes + 1024);
  uint64_t* mbar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mbar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 512u * row_bytes + 16u * row_bytes);
    tma_load_2d(sa, &tmap_a, bar, 0, 0);
    tma_load_2d(sa + 256u * row_bytes, &tmap_a, bar, 0, 256);
    tma_load_2d(sb, &tmap_b, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  __syncthreads();
  const uint32_t swz = (cblk == 64) ? SWZ_128B : (cblk == 32 ? SWZ_64B : SWZ_32B);
  const uint32_t idesc = make_idesc_bf16(128, 16, false, false);
  for (int c = 0; c < ncfg; ++c) {
    const ProbeCfg cfg = cfgs[c];
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t start = smem_u32(sa) + (uint32_t)cfg.shift * row_bytes;
      uint64_t a_desc = make_smem_desc(start, 16, (uint32_t)cfg.group_rows * row_bytes, swz);
      if (cfg.policy == 1) a_desc |= (uint64_t)((start >> 7) & 7u) << 49;
      const uint64_t b_desc = make_smem_desc(smem_u32(sb), 16, 8u * row_bytes, swz);
      for (int k = 0; k < cblk / 16; ++k) umma_bf16(tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, k != 0);
      umma_commit(mbar);
    }
    mbar_wait(mbar, (uint32_t)(c & 1));
    tc_fence_after();
    uint32_t r[16];
    tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    const int row = warp * 32 + lane;
    const int mem_row = (row / 8) * cfg.group_rows + (row % 8) + cfg.shift;
    int bad = 0;
    for (int n = 0; n < 16; ++n) {
      const int k = n * (cblk / 16);
      const float want = (float)((mem_row * 3 + k) % 251);
      if (__uint_as_float(r[n]) != want) ++bad;
    }
    if (bad) atomicAdd(&mismatches[c], bad);
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc<32>(tmem);
}

int probe_descriptors(int verbose) {
  int total_unexpected = 0;
  for (int cblk : {64, 32, 16}) {
    const int rows = 512;
    std::vector<__nv_bfloat16> ha((size_t)rows * cblk), hb((size_t)16 * cblk);
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < cblk; ++k) ha[(size_t)r * cblk + k] = __float2bfloat16((float)((r * 3 + k) % 251));
    for (int n = 0; n < 16; ++n)
      for (int k = 0; k < cblk; ++k) hb[(size_t)n * cblk + k] = __float2bfloat16(k == n * (cblk / 16) ? 1.f : 0.f);
    std::vector<ProbeCfg> cfgs;
    for (int g : {8, 12, 16, 20})
      for (int pol = 0; pol < 2; ++pol)
        for (int s = 0; s <= 9; ++s) cfgs.push_back({s, g, pol});
    __nv_bfloat16 *da, *db; ProbeCfg* dc; int* dm;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dc, cfgs.size() * sizeof(ProbeCfg)); cudaMalloc(&dm, cfgs.size() * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, cfgs.data(), cfgs.size() * sizeof(ProbeCfg), cudaMemcpyHostToDevice);
    cudaMemset(dm, 0, cfgs.size() * 4);
    CUtensorMap ta, tb;
    uint64_t dims_a[2] = {(uint64_t)cblk, (uint64_t)rows}, dims_b[2] = {(uint64_t)cblk, 16};
    uint64_t strides[1] = {(uint64_t)cblk * 2};
    uint32_t box_a[2] = {(uint32_t)cblk, 256}, box_b[2] = {(uint32_t)cblk, 16};
    if (make_tmap_bf16(&ta, da, 2, dims_a, strides, box_a, nullptr, cblk * 2)) return -1;
    if (make_tmap_bf16(&tb, db, 2, dims_b, strides, box_b, nullptr, cblk * 2)) return -1;
    const size_t smem = (size_t)528 * cblk * 2 + 4096;
    cudaFuncSetAttribute(probe_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_desc_kernel<<<1, 128, smem>>>(ta, tb, cblk, dc, (int)cfgs.size(), dm);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[probe] cblk=%d CUDA error %s\n", cblk, cudaGetErrorString(e)); return -2; }
    std::vector<int> hm(cfgs.size());
    cudaMemcpy(hm.data(), dm, cfgs.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < cfgs.size(); i += 10) {
      printf("[probe] cblk=%d row=%dB group_rows=%2d base_offset_policy=%d  mismatches for shift 0..9:", cblk, cblk * 2,
             cfgs[i].group_rows, cfgs[i].policy);
      for (int s = 0; s < 10; ++s) printf(" %4d", hm[i + s]);
      printf("\n");
    }
    if (hm[0] != 0) ++total_unexpected;     // shift 0, contiguous groups must always work
    cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dm);
  }
  (void)verbose;
  return total_unexpected;
}

}  // namespace jvae

extern "C" int jvae_probe_descriptors(int verbose) { return jvae::probe_descriptors(verbose); }

// ------------------------------------------------------------------------------------------------
// MMA pacing probe: cycles per tcgen05.mma (M=128, K=16, bf16) for N in {16,32,64,128,256}, issued back to back into
// ONE accumulator versus round-robin over several accumulators, operands in shared memory (contents irrelevant).
namespace jvae {

__global__ void __launch_bounds__(128, 1) probe_mma_rate_kernel(int N, int nacc, int iters, int cblk, long long* out, int M = 128) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    const uint32_t rb = (uint32_t)cblk * 2u;
    const uint32_t swz = (cblk == 64) ? SWZ_128B : (cblk == 32 ? SWZ_64B : SWZ_32B);
    const uint32_t idesc = make_idesc_bf16(M, N, false, false);
    const uint64_t a_desc = make_smem_desc(smem_u32(smem), 16, 8 * rb, swz);
    const uint64_t b_desc = make_smem_desc(smem_u32(smem) + 16384, 16, 8 * rb, swz);
    const long long t0 = clock64();
    const uint32_t mask = (uint32_t)nacc - 1u;      // nacc is a power of two
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
      const uint32_t a = (uint32_t)i & mask;
      umma_bf16(tmem + a * (uint32_t)N, a_desc + (uint64_t)(2 * (i & 1)), b_desc + (uint64_t)(2 * (i & 1)), idesc, 1);
    }
    umma_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int probe_mma_rate() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2048;
  for (int cblk : {64, 32}) {
    for (int N : {16, 32, 64, 128, 256}) {
      for (int nacc : {1, 2, 4, 8}) {
        if (nacc * N > 512) continue;
        probe_mma_rate_kernel<<<1, 128, 64 * 1024>>>(N, nacc, iters, cblk, d);
        probe_mma_rate_kernel<<<1, 128, 64 * 1024>>>(N, nacc, iters, cblk, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("[probe] mma rate: CUDA error\n"); return -1; }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("[probe] mma M=128 N=%3d K=16 rows=%dB accumulators=%d: issue %.1f clk/mma, complete %.1f clk/mma (ideal %d)\n", N,
               cblk * 2, nacc, (double)h[0] / iters, (double)h[1] / iters, 128 * N / 256);
      }
    }
  }
  // N that is not a power of two, and M = 64 (tap stacking produces both)
  for (int M : {128, 64})
    for (int N : {32, 48, 64, 96, 128, 160, 192, 224, 256}) {
      probe_mma_rate_kernel<<<1, 128, 64 * 1024>>>(N, 1, iters, 32, d, M);
      probe_mma_rate_kernel<<<1, 128, 64 * 1024>>>(N, 1, iters, 32, d, M);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("[probe] mma rate: CUDA error\n"); return -1; }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("[probe] mma M=%3d N=%3d K=16 rows=64B: complete %.1f clk/mma (M*N/256 = %d)\n", M, N, (double)h[1] / iters, M * N / 256);
    }
  cudaFree(d);
  return 0;
}
}  // namespace jvae

extern "C" int jvae_probe_mma_rate(void) { return jvae::probe_mma_rate(); }

// ------------------------------------------------------------------------------------------------
// Poison probe: fills the shared memory (220 KB per SM) and all 512 TMEM columns of every SM with a bit pattern, so that
// a kernel whose results depend on whatever a previous kernel left there shows up as a difference between two patterns.
namespace jvae {
__global__ void __launch_bounds__(128, 1) probe_poison_kernel(uint32_t pattern, int nwords) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw);
  for (int i = threadIdx.x; i < nwords; i += 128) w[i] = pattern;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  for (int c = 0; c < 512; c += 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};"
                 ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c), "r"(pattern) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace jvae

extern "C" int jvae_probe_poison(unsigned pattern, void* stream) {
  static bool attr = false;
  const int bytes = 220 * 1024;
  if (!attr) { cudaFuncSetAttribute(jvae::probe_poison_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); attr = true; }
  jvae::probe_poison_kernel<<<148 * 4, 128, bytes, reinterpret_cast<cudaStream_t>(stream)>>>(pattern, bytes / 4);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
