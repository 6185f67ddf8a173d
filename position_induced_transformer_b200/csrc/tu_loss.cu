// Launchers of the relative-Lp loss kernels.
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t rel_lp(int stage, const LossParams& P, int grid_x, cudaStream_t st) {
  const dim3 grid(grid_x, P.B);
  if (P.p == 2) {
    if (stage == 0) rel_lp_partial_kernel<2><<<grid, LOSS_THREADS, 0, st>>>(P);
    if (stage == 1) rel_lp_finalize_kernel<2><<<1, 32, 0, st>>>(P);
    if (stage == 2) rel_lp_backward_kernel<2><<<grid, LOSS_THREADS, 0, st>>>(P);
  } else {
    if (stage == 0) rel_lp_partial_kernel<1><<<grid, LOSS_THREADS, 0, st>>>(P);
    if (stage == 1) rel_lp_finalize_kernel<1><<<1, 32, 0, st>>>(P);
    if (stage == 2) rel_lp_backward_kernel<1><<<grid, LOSS_THREADS, 0, st>>>(P);
  }
  return cudaGetLastError();
}

}  // namespace launch
}  // namespace pit
