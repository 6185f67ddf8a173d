// Instantiations of the tcgen05 dense kernel.
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

namespace {
template <int GEO, int MODE, int NV>
cudaError_t one(const DenseParams& P, dim3 grid, cudaStream_t st) {
  return launch_smem(dense_attention_kernel<GEO, MODE, NV>, grid, DENSE_THREADS, (size_t)DenseSmem<MODE, NV>::TOTAL, P, st);
}
template <int GEO, int MODE>
cudaError_t by_nv(int nv, const DenseParams& P, dim3 grid, cudaStream_t st) {
  if (nv == 64) return one<GEO, MODE, 64>(P, grid, st);
  if (nv == 128) return one<GEO, MODE, 128>(P, grid, st);
  if constexpr (MODE != DENSE_DSCALE) return one<GEO, MODE, 256>(P, grid, st);
  return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t dense(int mode, int geo, int nv, dim3 grid, const DenseParams& P, cudaStream_t st) {
  return with_geo_only(geo, [&](auto g) -> cudaError_t {
    constexpr int G = decltype(g)::value;
    if (mode == DENSE_FWD) return by_nv<G, DENSE_FWD>(nv, P, grid, st);
    if (mode == DENSE_DSCALE) return by_nv<G, DENSE_DSCALE>(nv, P, grid, st);
    return by_nv<G, DENSE_DVALUES>(nv, P, grid, st);
  });
}

cudaError_t dense_bwd_pair(int geo, const DenseParams& Ps, dim3 gs, const DenseParams& Pv, dim3 gv, cudaStream_t st) {
  return with_geo_only(geo, [&](auto g) -> cudaError_t {
    constexpr int G = decltype(g)::value;
    auto kernel = dense_bwd_pair_kernel<G, 64>;
    constexpr size_t smem = DensePairSmem<64>::TOTAL;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int n_scale = gs.x * gs.y * gs.z, n_values = gv.x * gv.y * gv.z;
    kernel<<<n_scale + n_values, DENSE_THREADS, smem, st>>>(Ps, Pv, n_scale, gs, gv);
    return cudaGetLastError();
  });
}

}  // namespace launch
}  // namespace pit
