// Instantiations of the coordinate-gradient kernel (coord_gradient.cuh).
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

cudaError_t coord_gradient(int geo, const CoordGradParams& P, cudaStream_t st) {
  const int64_t items = (int64_t)P.B * P.N;
  const int grid = (int)((items + CG_WARPS - 1) / CG_WARPS);
  return with_geo_only(geo, [&](auto g) {
    coord_gradient_kernel<decltype(g)::value><<<grid, CG_WARPS * 32, 0, st>>>(P);
    return cudaGetLastError();
  });
}

}  // namespace launch
}  // namespace pit
