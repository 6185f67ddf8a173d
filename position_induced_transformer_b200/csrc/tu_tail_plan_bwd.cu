// Instantiations of tail_plan_bwd_kernel (decoder tail driven by a cached tile plan, backward).
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t tail_plan_backward(int geo, const TallPlan& plan, const TailParams& P, const TailPlanDev& V, cudaStream_t st) {
  {
    (void)geo;  // the plan carries the squared distances: the kernels do not depend on the distance variant
    auto go = [&](auto kernel) {
      if (plan.smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
        if (e != cudaSuccess) return e;
      }
      kernel<<<plan.grid, plan.threads, plan.smem, st>>>(P, V);
      return cudaGetLastError();
    };
    if (P.H == 1) return P.O == 1 ? go(tail_plan_bwd_kernel<1, 1>) : go(tail_plan_bwd_kernel<1, TAIL_MAX_OUT>);
    return P.O == 1 ? go(tail_plan_bwd_kernel<2, 1>) : go(tail_plan_bwd_kernel<2, TAIL_MAX_OUT>);
  }
}

}  // namespace launch
}  // namespace pit
