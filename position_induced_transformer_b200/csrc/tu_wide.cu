// Instantiations of the wide (encoder) kernels.
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

int wide_pad(int width) { return width <= 8 ? 8 : (width <= 16 ? 16 : (width <= 24 ? 24 : 32)); }

namespace {
template <typename F>
cudaError_t with_geo_heads_pad(int geo, int nh, int wpad, F&& f) {
  auto pad = [&](auto g, auto h) {
    if (wpad == 8) return f(g, h, Int<8>{});
    if (wpad == 16) return f(g, h, Int<16>{});
    if (wpad == 24) return f(g, h, Int<24>{});
    return f(g, h, Int<32>{});
  };
  return with_geo_only(geo, [&](auto g) { return nh == 1 ? pad(g, Int<1>{}) : pad(g, Int<2>{}); });
}
}  // namespace

cudaError_t wide_forward(int geo, const WidePlan& w, const WideParams& P, cudaStream_t st) {
  return with_geo_heads_pad(geo, P.H, wide_pad(P.width), [&](auto g, auto h, auto wp) {
    return launch_smem(wide_fwd_kernel<decltype(g)::value, decltype(h)::value, decltype(wp)::value>, dim3(w.grid), WIDE_THREADS,
                       w.smem, P, st);
  });
}

cudaError_t wide_dscale(int geo, const WidePlan& w, const WideParams& P, cudaStream_t st) {
  // the backward also keeps the upstream-gradient rows in shared memory
  const size_t smem = w.smem + (size_t)P.N * P.H * wide_pad(P.width) * sizeof(float);
  return with_geo_heads_pad(geo, P.H, wide_pad(P.width), [&](auto g, auto h, auto wp) {
    return launch_smem(wide_dscale_kernel<decltype(g)::value, decltype(h)::value, decltype(wp)::value>, dim3(w.grid),
                       WIDE_THREADS, smem, P, st);
  });
}

}  // namespace launch
}  // namespace pit
