// Launch of the fused gradient all-reduce + Adam kernel (allreduce_adam.cuh).
#include "launchers.h"

namespace pit {
namespace launch {

int allreduce_adam_grid(int64_t total, int sms) {
  int64_t g = (total / 4 + ARA_THREADS - 1) / ARA_THREADS;   // one group of four elements per thread: one NVLink round trip
  const int cap = 2 * sms;                  // every CTA must be resident at once (they wait for each other): two small CTAs per SM
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

cudaError_t allreduce_adam(const AllReduceAdamParams& P, int grid, cudaStream_t st) {
  if (P.world > 1) allreduce_adam_kernel<true><<<grid, ARA_THREADS, 0, st>>>(P);
  else allreduce_adam_kernel<false><<<grid, ARA_THREADS, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace launch
}  // namespace pit
