// Instantiations of the generic warp-per-row kernels and of the row-statistics kernels.
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

namespace {
template <typename F>
cudaError_t with_geo_vec_a(int geo, int vec, int a, F&& f) {
  auto pick_a = [&](auto g, auto v) {
    if (a == 1) return f(g, v, Int<1>{});
    if (a == 2) return f(g, v, Int<2>{});
    return f(g, v, Int<4>{});
  };
  return with_geo_only(geo, [&](auto g) { return vec == 4 ? pick_a(g, Int<4>{}) : pick_a(g, Int<1>{}); });
}

__global__ void reduce_scale_rows_kernel(const float* __restrict__ rows, int64_t n_rows, int H, float* __restrict__ d_scale) {
  __shared__ float red[32];
  const int h = blockIdx.x;
  float acc = 0.f;
  for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) acc += rows[r * H + h];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    d_scale[h] = t;
  }
}
}  // namespace

cudaError_t local_forward(int geo, int vec, int a, dim3 grid, const AttnParams& P, cudaStream_t st) {
  return with_geo_vec_a(geo, vec, a, [&](auto g, auto v, auto aa) {
    posatt_fwd_kernel<decltype(g)::value, decltype(v)::value, decltype(aa)::value><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(P);
    return cudaGetLastError();
  });
}
cudaError_t local_dscale(int geo, int vec, int a, dim3 grid, const AttnParams& P, cudaStream_t st) {
  return with_geo_vec_a(geo, vec, a, [&](auto g, auto v, auto aa) {
    posatt_dscale_kernel<decltype(g)::value, decltype(v)::value, decltype(aa)::value><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(P);
    return cudaGetLastError();
  });
}
cudaError_t local_dvalues(int geo, int vec, int a, dim3 grid, const AttnParams& P, cudaStream_t st) {
  return with_geo_vec_a(geo, vec, a, [&](auto g, auto v, auto aa) {
    posatt_dvalues_kernel<decltype(g)::value, decltype(v)::value, decltype(aa)::value><<<grid, WARPS_PER_BLOCK * 32, 0, st>>>(P);
    return cudaGetLastError();
  });
}
cudaError_t local_forward_finalize(int64_t total, const AttnParams& P, cudaStream_t st) {
  posatt_fwd_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P);
  return cudaGetLastError();
}
cudaError_t local_dscale_finalize(int64_t items, const AttnParams& P, float* rows, cudaStream_t st) {
  posatt_dscale_finalize_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(P, rows);
  return cudaGetLastError();
}
cudaError_t reduce_scale_rows(const float* rows, int64_t n_rows, int H, float* d_scale, cudaStream_t st) {
  reduce_scale_rows_kernel<<<H, 512, 0, st>>>(rows, n_rows, H, d_scale);
  return cudaGetLastError();
}

cudaError_t rowstat(int geo, const RowstatParams& R, cudaStream_t st) {
  return with_geo_only(geo, [&](auto g) {
    constexpr int G = decltype(g)::value;
    const unsigned warp_grid = (unsigned)((R.rows_total + 3) / 4);
    if (R.M <= 128) rowstat_warp_kernel<G, 4><<<warp_grid, 128, 0, st>>>(R);
    else if (R.M <= 256) rowstat_warp_kernel<G, 8><<<warp_grid, 128, 0, st>>>(R);
    else if (R.M <= 512) rowstat_warp_kernel<G, 16><<<warp_grid, 128, 0, st>>>(R);
    else if (R.M <= 1024) rowstat_warp_kernel<G, 32><<<warp_grid, 128, 0, st>>>(R);
    else rowstat_block_kernel<G><<<R.rows_total, ROWSTAT_BLOCK, 0, st>>>(R);
    return cudaGetLastError();
  });
}

}  // namespace launch
}  // namespace pit
