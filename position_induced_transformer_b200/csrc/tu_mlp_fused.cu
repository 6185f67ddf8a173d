// Instantiations of the fused narrow-input MLP kernels (mlp_fused.cuh).
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t mlp_fused(bool backward, int D, bool lin3, const MlpFusedParams& P, cudaStream_t st) {
  const int grid = (int)((P.R + MF_ROWS - 1) / MF_ROWS);
  auto go = [&](auto fwd, auto bwd) {
    if (backward) bwd<<<grid, PB_THREADS, 0, st>>>(P);
    else fwd<<<grid, PB_THREADS, 0, st>>>(P);
    return cudaGetLastError();
  };
  if (D == 32) return lin3 ? go(mlp_fused_fwd_kernel<32, true>, mlp_fused_bwd_kernel<32, true>) : go(mlp_fused_fwd_kernel<32, false>, mlp_fused_bwd_kernel<32, false>);
  return lin3 ? go(mlp_fused_fwd_kernel<64, true>, mlp_fused_bwd_kernel<64, true>) : go(mlp_fused_fwd_kernel<64, false>, mlp_fused_bwd_kernel<64, false>);
}

}  // namespace launch
}  // namespace pit
