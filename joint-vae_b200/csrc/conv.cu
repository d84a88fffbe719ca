// Implicit-GEMM convolution kernels on tcgen05 (placeholder until the kernels land).
#include "tc_common.cuh"
namespace jvae {
int conv_selftest(int verbose) { (void)verbose; return 0; }
}  // namespace jvae
