// K1p: the local encoder (few rows, huge column set, narrow values -- wide_attention.cuh) driven by a cached tile plan.
//
// wide_fwd_kernel / wide_dscale_kernel re-discover in every launch which of the 256 latent rows can see which of the
// 177 241 mesh columns (bounding-box tests plus ~13 exact row visits per 64 columns, 90.8 M pairs at Darcy-421).  Which
// (row, column) pairs CAN be kept does not depend on lmda (d2 <= v_hi(row)(1 + 1e-6) is a superset of every head's kept
// set), so the same tile plan as the decoder's is built once per mesh pair with the roles transposed
// (decoder_tail_plan.cuh, PlanBuildParams::transposed): the COLUMNS are sorted by the set of rows that can see them and cut
// into tiles of 32; a tile lists its candidate rows (~6 at Darcy-421: the encoder keeps 2 % of 177 241 columns per row, i.e.
// every column is seen by ~5 rows) and the bit-exact squared distances of its 32 columns to each of them.
//
//   wide_plan_fwd_kernel     lane = column: its value row (B*D <= 32 scalars) in registers, one weight per candidate row and
//                            head from the cached d2 (exact per-head cut as everywhere), the 32 lanes' contributions combined
//                            with a reduce-scatter butterfly and added to partial[(row, h), :] / rowsum[h, row] with REDs
//   wide_plan_dscale_kernel  the three per-(row, head) sums of the scale gradient (see wide_dscale_kernel) from the same walk
// The generic finalize kernels of local_attention.cuh turn the sums into the output / the gradient, as for the wide kernels.
#pragma once
#include "decoder_tail_plan.cuh"
#include "wide_attention.cuh"

namespace pit {

constexpr int WP_THREADS = 256;
constexpr int WP_WARPS = WP_THREADS / 32;

// Per-row constants (soft-max shift and cut of every head) in shared memory: [N][2*NH].
template <int NH>
__device__ __forceinline__ void wp_build_rows(const WideParams& P, float* rowtab, int* val_off) {
  for (int r = threadIdx.x; r < P.N; r += WP_THREADS) {
    const float vmin = __ldg(P.v_min + r);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float s = __ldg(P.scale + h);
      rowtab[(size_t)r * 2 * NH + 2 * h] = __fmul_rn(vmin, s);
      rowtab[(size_t)r * 2 * NH + 2 * h + 1] = P.masked ? head_threshold(__ldg(P.v_lo + r), __ldg(P.v_hi + r), s, P.weight) : INFINITY;
    }
  }
  for (int e = threadIdx.x; e < P.width; e += WP_THREADS) {
    const int b = e / P.D, d = e - b * P.D;
    val_off[e] = b * P.M * P.D + d;  // host guarantees B*M*D < 2^31
  }
}

template <int NH, int WPAD>
__global__ void __launch_bounds__(WP_THREADS) wide_plan_fwd_kernel(const WideParams P, const TailPlanDev V) {
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  float* rowtab = reinterpret_cast<float*>(wide_smem_raw);
  int* val_off = reinterpret_cast<int*>(rowtab + (size_t)P.N * 2 * NH);
  wp_build_rows<NH>(P, rowtab, val_off);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  for (int tile = blockIdx.x * WP_WARPS + warp; tile < V.n_tiles; tile += gridDim.x * WP_WARPS) {
    const int off = __ldg(V.tile_off + tile), cnt = __ldg(V.tile_cnt + tile);
    const int j = __float_as_int(__ldg(V.rec + (size_t)tile * TP_ROWS + lane).w);  // this lane's column (-1: padding)
    float u[WPAD];
#pragma unroll
    for (int e = 0; e < WPAD; ++e) u[e] = (e < P.width && j >= 0) ? __ldg(P.values + val_off[e] + (int64_t)j * P.D) : 0.f;
    for (int k = 0; k < cnt; ++k) {
      const int r = (int)__ldg(V.cand + off + k);
      const float d2 = __ldg(V.d2 + (size_t)(off + k) * TP_ROWS + lane);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float p = 0.f;
        if (j >= 0) {
          const float sc = __fmul_rn(d2, s[h]);
          if (sc <= rowtab[(size_t)r * 2 * NH + 2 * h + 1]) p = expf(__fsub_rn(rowtab[(size_t)r * 2 * NH + 2 * h], sc));
        }
        const float lsum = warp_sum(p);
        if (lsum > 0.f) {  // warp-uniform
          // lane e ends up with element e of the tile's contribution: a reduce-scatter butterfly (31 shuffles) instead of one
          // full warp reduction per element
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = e < WPAD ? p * u[e] : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; ++i) {
              const float keep = upper ? v[i + o] : v[i], send = upper ? v[i] : v[i + o];
              v[i] = keep + __shfl_xor_sync(FULL, send, o);
            }
          }
          if (lane < P.width) atomicAdd(P.partial + ((int64_t)r * NH + h) * P.width + lane, v[0]);
          if (lane == 0) atomicAdd(P.rowsum + (int64_t)h * P.N + r, lsum);
        }
      }
    }
  }
}

template <int NH, int WPAD>
__global__ void __launch_bounds__(WP_THREADS) wide_plan_dscale_kernel(const WideParams P, const TailPlanDev V) {
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  float* rowtab = reinterpret_cast<float*>(wide_smem_raw);
  int* val_off = reinterpret_cast<int*>(rowtab + (size_t)P.N * 2 * NH);
  float* gtab = reinterpret_cast<float*>(val_off + P.width);  // [N][NH][WPAD] upstream gradient rows
  wp_build_rows<NH>(P, rowtab, val_off);
  for (int rh = threadIdx.x; rh < P.N * NH; rh += WP_THREADS) {
    const int r = rh / NH, h = rh - r * NH;
    const float* src = P.d_out + (int64_t)r * P.ld_out + P.col_off + (int64_t)h * P.D;
    float* g = gtab + (size_t)rh * WPAD;
    int b = 0, d = 0;
#pragma unroll
    for (int e = 0; e < WPAD; ++e) {
      g[e] = e < P.width ? __ldg(src + (int64_t)b * P.N * P.ld_out + d) : 0.f;
      if (++d == P.D) {
        d = 0;
        ++b;
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  for (int tile = blockIdx.x * WP_WARPS + warp; tile < V.n_tiles; tile += gridDim.x * WP_WARPS) {
    const int off = __ldg(V.tile_off + tile), cnt = __ldg(V.tile_cnt + tile);
    const int j = __float_as_int(__ldg(V.rec + (size_t)tile * TP_ROWS + lane).w);
    float u[WPAD];
#pragma unroll
    for (int e = 0; e < WPAD; ++e) u[e] = (e < P.width && j >= 0) ? __ldg(P.values + val_off[e] + (int64_t)j * P.D) : 0.f;
    for (int k = 0; k < cnt; ++k) {
      const int r = (int)__ldg(V.cand + off + k);
      const float d2 = __ldg(V.d2 + (size_t)(off + k) * TP_ROWS + lane);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        float p = 0.f;
        if (j >= 0) {
          const float sc = __fmul_rn(d2, s[h]);
          if (sc <= rowtab[(size_t)r * 2 * NH + 2 * h + 1]) p = expf(__fsub_rn(rowtab[(size_t)r * 2 * NH + 2 * h], sc));
        }
        if (!__any_sync(FULL, p > 0.f)) continue;
        const float* g = gtab + ((size_t)r * NH + h) * WPAD;
        float dp = 0.f;  // <dO[row,h,:], U[j,:]>
#pragma unroll
        for (int e = 0; e < WPAD; ++e) dp = fmaf(g[e], u[e], dp);
        const float pd = p * d2;
        const float a_sum = warp_sum(pd * dp), b_sum = warp_sum(p * dp), m_sum = warp_sum(pd);
        if (lane == 0) {
          float* dst = P.dscale_terms + ((int64_t)r * NH + h) * 3;
          atomicAdd(dst + 0, a_sum);
          atomicAdd(dst + 1, b_sum);
          atomicAdd(dst + 2, m_sum);
        }
      }
    }
  }
}

}  // namespace pit
