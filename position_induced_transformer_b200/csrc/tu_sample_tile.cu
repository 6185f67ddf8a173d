// Instantiations of the per-sample tiled kernels (sample_tile.cuh).
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

int sample_tile_slots(int M, int D, bool backward, int smem_optin) {
  // The kernels are latency-bound (shuffle -> shared load -> FMA chains per list entry), so resident warps matter more than slot
  // capacity: four CTAs per SM if at least 40 slots fit into a quarter of the shared memory, else two, else one (<= 128 slots).
  const long fixed = (long)sample_tile_smem_bytes(M, D, 0, backward);
  const long per_slot = (long)D * 4 * (backward ? 2 : 1) + 2;
  const int share[3] = {4, 2, 1};
  for (int i = 0; i < 3; ++i) {
    const long budget = smem_optin / share[i] - 1536;
    long ns = (budget - fixed) / per_slot;
    if (ns > 128) ns = 128;
    if (ns >= 40 || share[i] == 1) return ns < 8 ? 0 : (int)(ns / 8 * 8);
  }
  return 0;
}

cudaError_t sample_tile(bool backward, const SampleTileParams& P, cudaStream_t st) {
  const dim3 grid((P.N + ST_ROWS - 1) / ST_ROWS, P.B);
  const size_t smem = sample_tile_smem_bytes(P.M, P.D, P.n_slots, backward);
  auto go = [&](auto kernel) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kernel<<<grid, ST_THREADS, smem, st>>>(P);
    return cudaGetLastError();
  };
  const bool two = P.H == 2, wide = P.D > 128;
  if (!backward) {
    if (two) return wide ? go(sample_tile_kernel<2, 2, false>) : go(sample_tile_kernel<2, 1, false>);
    return wide ? go(sample_tile_kernel<1, 2, false>) : go(sample_tile_kernel<1, 1, false>);
  }
  if (two) return wide ? go(sample_tile_kernel<2, 2, true>) : go(sample_tile_kernel<2, 1, true>);
  return wide ? go(sample_tile_kernel<1, 2, true>) : go(sample_tile_kernel<1, 1, true>);
}

}  // namespace launch
}  // namespace pit
