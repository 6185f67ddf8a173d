// Construction of the decoder-tail tile plan (decoder_tail_plan.cuh): candidate-set keys, a radix sort of the rows
// (CUB, device-wide, runs once per mesh pair), per-tile candidate counts, their exclusive scan, and the fill pass.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

namespace {
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct PlanScratch {
  unsigned long long *keys_in, *keys_out;
  int32_t *rows_in, *rows_out, *tile_pad;
  void* cub_temp;
  size_t cub_bytes, total;
};

PlanScratch carve(void* ws, int N, int n_tiles) {
  PlanScratch s{};
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, N);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, n_tiles + 1);
  s.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  unsigned char* p = static_cast<unsigned char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* q = p ? p + off : nullptr;
    off += align256(bytes);
    return q;
  };
  s.keys_in = reinterpret_cast<unsigned long long*>(take((size_t)N * 8));
  s.keys_out = reinterpret_cast<unsigned long long*>(take((size_t)N * 8));
  s.rows_in = reinterpret_cast<int32_t*>(take((size_t)N * 4));
  s.rows_out = reinterpret_cast<int32_t*>(take((size_t)N * 4));
  s.tile_pad = reinterpret_cast<int32_t*>(take((size_t)(n_tiles + 1) * 4));
  s.cub_temp = take(s.cub_bytes);
  s.total = off;
  return s;
}

template <typename F>
cudaError_t with_geo_cpl(int geo, int cpl, F&& f) {
  return with_geo_only(geo, [&](auto g) { return cpl == 8 ? f(g, Int<8>{}) : f(g, Int<32>{}); });
}
}  // namespace

size_t tail_plan_workspace_bytes(int N) { return carve(nullptr, N, (N + TP_ROWS - 1) / TP_ROWS).total; }

// Stage 1: sort the rows by candidate set, count the candidates of every 32-row tile (tile_cnt), scan the counts rounded up
// to multiples of 8 into tile_off ([n_tiles+1]; its last entry is the length the caller sizes `cand` and `d2` with).
cudaError_t tail_plan_rows(int geo, int cpl, PlanBuildParams B, int32_t* tile_off, int32_t* tile_cnt, void* ws, cudaStream_t st) {
  const PlanScratch s = carve(ws, B.N, B.n_tiles);
  B.keys = s.keys_in;
  B.rows = s.rows_in;
  cudaError_t e = with_geo_cpl(geo, cpl, [&](auto g, auto c) {
    plan_key_kernel<decltype(g)::value, decltype(c)::value><<<(B.N + 3) / 4, 128, 0, st>>>(B);
    return cudaGetLastError();
  });
  if (e != cudaSuccess) return e;
  size_t bytes = s.cub_bytes;
  e = cub::DeviceRadixSort::SortPairs(s.cub_temp, bytes, s.keys_in, s.keys_out, s.rows_in, s.rows_out, B.N, 0, 64, st);
  if (e != cudaSuccess) return e;
  B.perm = s.rows_out;
  B.tile_cnt = tile_cnt;
  B.tile_pad = s.tile_pad;
  e = with_geo_cpl(geo, cpl, [&](auto g, auto c) {
    plan_tile_kernel<decltype(g)::value, decltype(c)::value, false><<<(B.n_tiles + 1 + 3) / 4, 128, 0, st>>>(B);
    return cudaGetLastError();
  });
  if (e != cudaSuccess) return e;
  bytes = s.cub_bytes;
  return cub::DeviceScan::ExclusiveSum(s.cub_temp, bytes, s.tile_pad, tile_off, B.n_tiles + 1, st);
}

// Stage 2: candidate lists and row records (same workspace as stage 1: it still holds the sorted rows).
cudaError_t tail_plan_fill(int geo, int cpl, PlanBuildParams B, void* ws, cudaStream_t st) {
  const PlanScratch s = carve(ws, B.N, B.n_tiles);
  B.perm = s.rows_out;
  return with_geo_cpl(geo, cpl, [&](auto g, auto c) {
    plan_tile_kernel<decltype(g)::value, decltype(c)::value, true><<<(B.n_tiles + 3) / 4, 128, 0, st>>>(B);
    return cudaGetLastError();
  });
}

}  // namespace launch
}  // namespace pit
