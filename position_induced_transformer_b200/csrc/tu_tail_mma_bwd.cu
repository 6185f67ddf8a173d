// Instantiations of tail_mma_bwd_kernel (decoder tail on mma.sync, backward).
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t tail_mma_backward(int geo, const TallPlan& plan, const TailParams& P, cudaStream_t st) {
  return with_geo(geo, plan.cpl, P.H, 1, [&](auto g, auto c, auto h, auto) {
    constexpr int G = decltype(g)::value, C = decltype(c)::value, H = decltype(h)::value;
    if (P.O == 1) return launch_smem(tail_mma_bwd_kernel<G, C, H, 1>, dim3(plan.grid), plan.threads, plan.smem, P, st);
    return launch_smem(tail_mma_bwd_kernel<G, C, H, TAIL_MAX_OUT>, dim3(plan.grid), plan.threads, plan.smem, P, st);
  });
}

}  // namespace launch
}  // namespace pit
