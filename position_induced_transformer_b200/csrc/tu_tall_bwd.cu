// Instantiations of tall_bwd_kernel.
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t tall_backward(int geo, const TallPlan& plan, const TallParams& P, bool with_values, cudaStream_t st) {
  return with_geo(geo, plan.cpl, P.H, plan.l4, [&](auto g, auto c, auto h, auto l) {
    constexpr int G = decltype(g)::value, C = decltype(c)::value, NH = decltype(h)::value, L = decltype(l)::value;
    const dim3 grid(plan.grid, plan.chunks);
    return with_values ? launch_smem(tall_bwd_kernel<G, C, NH, L, true>, grid, TALL_THREADS, plan.smem, P, st)
                       : launch_smem(tall_bwd_kernel<G, C, NH, L, false>, grid, TALL_THREADS, plan.smem, P, st);
  });
}

}  // namespace launch
}  // namespace pit
