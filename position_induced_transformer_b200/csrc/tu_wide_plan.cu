// Instantiations of the plan-driven encoder kernels (encoder_plan.cuh).
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

namespace {
template <typename F>
cudaError_t with_heads_pad(int nh, int wpad, F&& f) {
  auto pad = [&](auto h) {
    if (wpad == 8) return f(h, Int<8>{});
    if (wpad == 16) return f(h, Int<16>{});
    if (wpad == 24) return f(h, Int<24>{});
    return f(h, Int<32>{});
  };
  return nh == 1 ? pad(Int<1>{}) : pad(Int<2>{});
}
}  // namespace

cudaError_t wide_plan_forward(int grid, const WideParams& P, const TailPlanDev& V, cudaStream_t st) {
  const size_t smem = ((size_t)P.N * 2 * P.H + P.width) * sizeof(float);
  return with_heads_pad(P.H, wide_pad(P.width), [&](auto h, auto wp) {
    auto kernel = wide_plan_fwd_kernel<decltype(h)::value, decltype(wp)::value>;
    kernel<<<grid, WP_THREADS, smem, st>>>(P, V);
    return cudaGetLastError();
  });
}

cudaError_t wide_plan_dscale(int grid, const WideParams& P, const TailPlanDev& V, cudaStream_t st) {
  const size_t smem = ((size_t)P.N * 2 * P.H + P.width + (size_t)P.N * P.H * wide_pad(P.width)) * sizeof(float);
  return with_heads_pad(P.H, wide_pad(P.width), [&](auto h, auto wp) {
    auto kernel = wide_plan_dscale_kernel<decltype(h)::value, decltype(wp)::value>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kernel<<<grid, WP_THREADS, smem, st>>>(P, V);
    return cudaGetLastError();
  });
}

}  // namespace launch
}  // namespace pit
