// Launchers of the MLP epilogue kernels.
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t bias_act(bool backward, const EpiParams& P, int grid, cudaStream_t st) {
  if (backward) {
    if (P.gelu)
      bias_act_bwd_kernel<true><<<grid, EPI_THREADS, 0, st>>>(P);
    else
      bias_act_bwd_kernel<false><<<grid, EPI_THREADS, 0, st>>>(P);
  } else {
    if (P.gelu)
      bias_act_fwd_kernel<true><<<grid, EPI_THREADS, 0, st>>>(P);
    else
      bias_act_fwd_kernel<false><<<grid, EPI_THREADS, 0, st>>>(P);
  }
  return cudaGetLastError();
}

}  // namespace launch
}  // namespace pit
