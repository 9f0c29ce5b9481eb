// tcgen05 implicit-GEMM convolution (BC_PREC_BF16 / BC_PREC_BF16X3) -- placeholder until the
// UMMA kernel lands; reports "unsupported" rather than silently falling back.
#include "common.cuh"
namespace bc {
int conv1d_tc_fwd(const float*, const float*, const float*, const float*, const float*, const float*, float*, int, int,
                  int, int, int, int, int, int, int, int, int, int, int, int precision, cudaStream_t) {
  return fail(BC_EUNSUPPORTED, "conv1d: precision mode %d (tensor-core path) is not built yet", precision);
}
}  // namespace bc
