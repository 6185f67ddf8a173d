// Runtime -> compile-time dispatch helpers shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

#include "geometry.cuh"

namespace pit {
namespace launch {

template <int V>
using Int = std::integral_constant<int, V>;

// Calls f(Int<GEO>, Int<CPL>, Int<NH>, Int<L4>) for the runtime values.  CPL is 8 (M <= 256) or 32 (M <= 1024).
template <typename G, typename C, typename H, typename F>
cudaError_t with_l4(G g, C c, H h, int l4, F&& f) {
  if (l4 == 1) return f(g, c, h, Int<1>{});
  if (l4 == 2) return f(g, c, h, Int<2>{});
  return f(g, c, h, Int<4>{});
}
template <typename G, typename C, typename F>
cudaError_t with_heads(G g, C c, int nh, int l4, F&& f) {
  if (nh == 1) return with_l4(g, c, Int<1>{}, l4, f);
  return with_l4(g, c, Int<2>{}, l4, f);
}
template <typename G, typename F>
cudaError_t with_cpl(G g, int cpl, int nh, int l4, F&& f) {
  if (cpl == 8) return with_heads(g, Int<8>{}, nh, l4, f);
  return with_heads(g, Int<32>{}, nh, l4, f);
}
template <typename F>
cudaError_t with_geo(int geo, int cpl, int nh, int l4, F&& f) {
  if (geo == GEO_EUCLID1) return with_cpl(Int<GEO_EUCLID1>{}, cpl, nh, l4, f);
  if (geo == GEO_EUCLID2) return with_cpl(Int<GEO_EUCLID2>{}, cpl, nh, l4, f);
  if (geo == GEO_PERIODIC1) return with_cpl(Int<GEO_PERIODIC1>{}, cpl, nh, l4, f);
  return with_cpl(Int<GEO_PERIODIC2>{}, cpl, nh, l4, f);
}

template <typename F>
cudaError_t with_geo_only(int geo, F&& f) {
  if (geo == GEO_EUCLID1) return f(Int<GEO_EUCLID1>{});
  if (geo == GEO_EUCLID2) return f(Int<GEO_EUCLID2>{});
  if (geo == GEO_PERIODIC1) return f(Int<GEO_PERIODIC1>{});
  return f(Int<GEO_PERIODIC2>{});
}

// Launch with the dynamic shared-memory opt-in when more than 48 KB is requested.
template <typename K, typename P>
cudaError_t launch_smem(K kernel, dim3 grid, int threads, size_t smem, const P& params, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kernel<<<grid, threads, smem, st>>>(params);
  return cudaGetLastError();
}

}  // namespace launch
}  // namespace pit
