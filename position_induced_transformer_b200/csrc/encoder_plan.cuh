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
// Per tile the work is a small dense product, run on mma.sync.m16n8k8 (3xTF32, fp32 accumulate) by one warp:
//   wide_plan_fwd_kernel     out^T[e x (row, h)] += U^T[e x 32 columns] . P[32 columns x (row, h)]   (e = the B*D <= 32 value
//                            scalars of a column): the weights are evaluated directly in B-fragment order from the cached d2
//                            (exact per-head cut as everywhere), 8 candidate rows x 2 heads per group; the accumulators go to
//                            partial[(row, h), :] with REDs, the row sums to rowsum[h, row]
//   wide_plan_dscale_kernel  dP[32 columns x (row, h)] = U . dO[row, h, :]^T the same way, then the three per-(row, head) sums
//                            of the scale gradient (see wide_dscale_kernel) from the accumulator fragments
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

// Weight of (tile column c, candidate k) for head h; rowc = {top, cut} of the candidate's row.
__device__ __forceinline__ float wp_weight(float d2, float s, float top, float cut, bool live) {
  if (!live) return 0.f;
  const float sc = __fmul_rn(d2, s);
  return sc <= cut ? expf(__fsub_rn(top, sc)) : 0.f;
}

template <int NH, int WPAD>
__global__ void __launch_bounds__(WP_THREADS) wide_plan_fwd_kernel(const WideParams P, const TailPlanDev V) {
  constexpr int MT = (WPAD + 15) / 16;  // m16 tiles over the value scalars
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  float* rowtab = reinterpret_cast<float*>(wide_smem_raw);
  int* val_off = reinterpret_cast<int*>(rowtab + (size_t)P.N * 2 * NH);
  wp_build_rows<NH>(P, rowtab, val_off);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  for (int tile = blockIdx.x * WP_WARPS + warp; tile < V.n_tiles; tile += gridDim.x * WP_WARPS) {
    const int off = __ldg(V.tile_off + tile), cnt = __ldg(V.tile_cnt + tile);
    const int jc = __float_as_int(__ldg(V.rec + (size_t)tile * TP_ROWS + lane).w);  // column of tile slot `lane` (-1: padding)
    // A = U^T: a[ks][mt] = {U[col 8ks+t][e], U[col 8ks+t][e+8], U[col 8ks+t+4][e], U[col 8ks+t+4][e+8]}, e = 16mt + g
    float a[4][MT][4];
    bool clive[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = __shfl_sync(FULL, jc, 8 * ks + t + 4 * half);
        clive[ks][half] = j >= 0;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int up = 0; up < 2; ++up) {
            const int e = 16 * mt + g + 8 * up;
            a[ks][mt][2 * half + up] = (j >= 0 && e < P.width) ? __ldg(P.values + val_off[e] + (int64_t)j * P.D) : 0.f;
          }
      }
    }
    for (int cg = 0; cg < cnt; cg += 8) {  // groups of 8 candidate rows: n = candidate cg + g
      const bool klive = cg + g < cnt;
      const int r = klive ? (int)__ldg(V.cand + off + cg + g) : 0;
      float top[NH], cut[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) top[h] = rowtab[(size_t)r * 2 * NH + 2 * h], cut[h] = rowtab[(size_t)r * 2 * NH + 2 * h + 1];
      float acc[MT][NH][4], lsum[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        lsum[h] = 0.f;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mt][h][i] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        // B = P: b = {P[col 8ks+t][cand], P[col 8ks+t+4][cand]} per head
        float d2[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) d2[half] = klive ? __ldg(V.d2 + (size_t)(off + cg + g) * TP_ROWS + 8 * ks + t + 4 * half) : 0.f;
        uint32_t al[MT][4], ah[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) ah[mt][i] = __float_as_uint(a[ks][mt][i]), al[mt][i] = tm_trunc_lo(a[ks][mt][i]);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float p0 = wp_weight(d2[0], s[h], top[h], cut[h], klive && clive[ks][0]);
          const float p1 = wp_weight(d2[1], s[h], top[h], cut[h], klive && clive[ks][1]);
          lsum[h] += p0 + p1;
          const float h0 = tm_round_hi(p0), h1 = tm_round_hi(p1);
          const uint32_t bh[2] = {__float_as_uint(h0), __float_as_uint(h1)}, bl[2] = {__float_as_uint(p0 - h0), __float_as_uint(p1 - h1)};
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_tf32_16x8x8(acc[mt][h], al[mt], bh);
            mma_tf32_16x8x8(acc[mt][h], ah[mt], bl);
            mma_tf32_16x8x8(acc[mt][h], ah[mt], bh);
          }
        }
      }
      // row sums: this lane covered 8 of the 32 columns of candidate g; the other 24 sit in the lanes sharing g
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        lsum[h] += __shfl_xor_sync(FULL, lsum[h], 1);
        lsum[h] += __shfl_xor_sync(FULL, lsum[h], 2);
        if (t == 0 && klive && lsum[h] > 0.f) atomicAdd(P.rowsum + (int64_t)h * P.N + r, lsum[h]);
      }
      // accumulators: value scalar e = 16mt + g (+8), candidate cg + 2t (+1)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = cg + 2 * t + q;
        if (k >= cnt) continue;
        const int rk = (int)__ldg(V.cand + off + k);
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int up = 0; up < 2; ++up) {
              const int e = 16 * mt + g + 8 * up;
              const float v = acc[mt][h][2 * up + q];
              if (e < P.width && v != 0.f) atomicAdd(P.partial + ((int64_t)rk * NH + h) * P.width + e, v);
            }
      }
    }
  }
}

template <int NH, int WPAD>
__global__ void __launch_bounds__(WP_THREADS) wide_plan_dscale_kernel(const WideParams P, const TailPlanDev V) {
  constexpr int KS = (WPAD + 7) / 8;  // k-steps over the value scalars
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  float* rowtab = reinterpret_cast<float*>(wide_smem_raw);
  int* val_off = reinterpret_cast<int*>(rowtab + (size_t)P.N * 2 * NH);
  float* gtab = reinterpret_cast<float*>(val_off + P.width);  // [N][NH][WPAD] upstream gradient rows
  wp_build_rows<NH>(P, rowtab, val_off);
  for (int rh = threadIdx.x; rh < P.N * NH; rh += WP_THREADS) {
    const int r = rh / NH, h = rh - r * NH;
    const float* src = P.d_out + (int64_t)r * P.ld_out + P.col_off + (int64_t)h * P.D;
    float* gr = gtab + (size_t)rh * WPAD;
    int b = 0, d = 0;
#pragma unroll
    for (int e = 0; e < WPAD; ++e) {
      gr[e] = e < P.width ? __ldg(src + (int64_t)b * P.N * P.ld_out + d) : 0.f;
      if (++d == P.D) {
        d = 0;
        ++b;
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  for (int tile = blockIdx.x * WP_WARPS + warp; tile < V.n_tiles; tile += gridDim.x * WP_WARPS) {
    const int off = __ldg(V.tile_off + tile), cnt = __ldg(V.tile_cnt + tile);
    const int jc = __float_as_int(__ldg(V.rec + (size_t)tile * TP_ROWS + lane).w);
    // A = U: a[mt][ks] = {U[col 16mt+g][e], U[col 16mt+g+8][e], U[col 16mt+g][e+4], U[col 16mt+g+8][e+4]}, e = 8ks + t
    float a[2][KS][4];
    bool clive[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int up = 0; up < 2; ++up) {
        const int j = __shfl_sync(FULL, jc, 16 * mt + g + 8 * up);
        clive[mt][up] = j >= 0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int e = 8 * ks + t + 4 * half;
            a[mt][ks][2 * half + up] = (j >= 0 && e < P.width) ? __ldg(P.values + val_off[e] + (int64_t)j * P.D) : 0.f;
          }
      }
    for (int cg = 0; cg < cnt; cg += 8) {
      // B = dO^T: n = candidate cg + g; b = {G[row][h][8ks+t], G[row][h][8ks+t+4]}
      const bool nlive = cg + g < cnt;
      const int rn = nlive ? (int)__ldg(V.cand + off + cg + g) : 0;
      float dp[2][NH][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i) dp[mt][h][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t al[2][4], ah[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) ah[mt][i] = __float_as_uint(a[mt][ks][i]), al[mt][i] = tm_trunc_lo(a[mt][ks][i]);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float* gr = gtab + ((size_t)rn * NH + h) * WPAD + 8 * ks + t;
          const float g0 = (nlive && 8 * ks + t < WPAD) ? gr[0] : 0.f, g1 = (nlive && 8 * ks + t + 4 < WPAD) ? gr[4] : 0.f;
          const float h0 = tm_round_hi(g0), h1 = tm_round_hi(g1);
          const uint32_t bh[2] = {__float_as_uint(h0), __float_as_uint(h1)}, bl[2] = {__float_as_uint(g0 - h0), __float_as_uint(g1 - h1)};
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_tf32_16x8x8(dp[mt][h], al[mt], bh);
            mma_tf32_16x8x8(dp[mt][h], ah[mt], bl);
            mma_tf32_16x8x8(dp[mt][h], ah[mt], bh);
          }
        }
      }
      // dp[mt][h][2*up + q]: column 16mt + g + 8up, candidate cg + 2t + q.  Weights in the same layout, then the three sums.
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = cg + 2 * t + q;
        const bool klive = k < cnt;
        const int rk = klive ? (int)__ldg(V.cand + off + k) : 0;
        float sums[NH][3];
#pragma unroll
        for (int h = 0; h < NH; ++h) sums[h][0] = sums[h][1] = sums[h][2] = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int up = 0; up < 2; ++up) {
            const bool live = klive && clive[mt][up];
            const float d2 = live ? __ldg(V.d2 + (size_t)(off + k) * TP_ROWS + 16 * mt + g + 8 * up) : 0.f;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const float p = wp_weight(d2, s[h], rowtab[(size_t)rk * 2 * NH + 2 * h], rowtab[(size_t)rk * 2 * NH + 2 * h + 1], live);
              const float pd = p * d2, dpv = dp[mt][h][2 * up + q];
              sums[h][0] = fmaf(pd, dpv, sums[h][0]);
              sums[h][1] = fmaf(p, dpv, sums[h][1]);
              sums[h][2] += pd;
            }
          }
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            float v = sums[h][i];
            v += __shfl_xor_sync(FULL, v, 4);
            v += __shfl_xor_sync(FULL, v, 8);
            v += __shfl_xor_sync(FULL, v, 16);
            if (g == 0 && klive && v != 0.f) atomicAdd(P.dscale_terms + ((int64_t)rk * NH + h) * 3 + i, v);
          }
      }
    }
  }
}

}  // namespace pit
