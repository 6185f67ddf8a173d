// Instantiations of tall_fwd_kernel.
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t tall_forward(int geo, const TallPlan& plan, const TallParams& P, cudaStream_t st) {
  return with_geo(geo, plan.cpl, P.H, plan.l4, [&](auto g, auto c, auto h, auto l) {
    return launch_smem(tall_fwd_kernel<decltype(g)::value, decltype(c)::value, decltype(h)::value, decltype(l)::value>,
                       dim3(plan.grid, plan.chunks), TALL_THREADS, plan.smem, P, st);
  });
}

}  // namespace launch
}  // namespace pit
