// Instantiations of tail_mma_bwd_kernel (decoder tail on mma.sync, backward).
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t tail_mma_backward(int geo, const TallPlan& plan, const TailParams& P, cudaStream_t st) {
  return with_geo(geo, plan.cpl, P.H, 1, [&](auto g, auto c, auto h, auto) {
    return launch_smem(tail_mma_bwd_kernel<decltype(g)::value, decltype(c)::value, decltype(h)::value>, dim3(plan.grid), plan.threads,
                       plan.smem, P, st);
  });
}

}  // namespace launch
}  // namespace pit
