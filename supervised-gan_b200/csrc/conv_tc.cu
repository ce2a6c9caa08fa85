// tcgen05 / TMEM implicit-GEMM convolution (placeholder until the tensor-core path lands).
#include "conv_plan.h"
namespace sgk {
int conv_fwd_tc(const SgkConvDesc*, const GatherPlan&, const float*, const float*, const float*, float*, int, float,
                cudaStream_t) {
  return SGK_EUNSUPPORTED;
}
int conv_wgrad_tc(const SgkConvDesc*, const float*, const float*, float*, void*, size_t, cudaStream_t) {
  return SGK_EUNSUPPORTED;
}
}  // namespace sgk
