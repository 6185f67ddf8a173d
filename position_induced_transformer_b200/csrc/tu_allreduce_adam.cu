// Launch of the fused gradient all-reduce + Adam kernel (allreduce_adam.cuh).
#include "launchers.h"

namespace pit {
namespace launch {

int allreduce_adam_grid(int64_t total, int sms, bool multi) {
  int64_t g = (total / 4 + ARA_THREADS - 1) / ARA_THREADS;   // one group of four elements per thread: one NVLink round trip
  // every CTA must be resident at once (they wait for each other): ask the runtime how many fit, never assume
  int per_sm = 0;
  cudaError_t e = multi ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, allreduce_adam_kernel<true>, ARA_THREADS, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, allreduce_adam_kernel<false>, ARA_THREADS, 0);
  if (e != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 1;
  }
  if (per_sm > 2) per_sm = 2;
  const int64_t cap = (int64_t)per_sm * sms;
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
