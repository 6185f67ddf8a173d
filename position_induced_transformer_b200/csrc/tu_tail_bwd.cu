// Instantiations of tail_bwd_kernel.
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t tail_backward(int geo, const TallPlan& plan, const TailParams& P, cudaStream_t st) {
  return with_geo(geo, plan.cpl, P.H, plan.l4, [&](auto g, auto c, auto h, auto l) {
    return launch_smem(tail_bwd_kernel<decltype(g)::value, decltype(c)::value, decltype(h)::value, decltype(l)::value>,
                       dim3(plan.grid, plan.chunks), TALL_THREADS, plan.smem, P, st);
  });
}

}  // namespace launch
}  // namespace pit
