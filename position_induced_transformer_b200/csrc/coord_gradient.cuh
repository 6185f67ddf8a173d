// K6: gradient of position-attention with respect to the mesh coordinates (`dX`).
//
// The reference's dist2att is a differentiable function of mesh_out / mesh_in (pit.py:46-52); no script uses that (meshes never
// require grad), but a learnable latent mesh does.  With P^ the normalised weights of a row, dP_ij = <dO_i, U_j> and
// delta_i = sum_j P^_ij dP_ij, the gradient with respect to the logit -s_h d2_ij is dS_ij = P^_ij (dP_ij - delta_i), hence
//     dL/d d2_ij = -s_h dS_ij,      d d2 / d xo = 2 (xo - xi) = -d d2 / d xi                                 (Euclidean)
// and, for the periodic variants, per axis with m = |xo - xi| and m' = min(m, l - m) (pit.py:193-195, 251-253):
//     d d2 / d xo = 2 m' (+-1) sign(xo - xi)   (+ if m < l - m, - if m > l - m, 0 at a tie),    d d2 / d l = 2 m' [m > l - m]  (1/2 at a tie)
// exactly what autograd does with abs / minimum; the wrap length l is itself a function of mesh_in (pit.py:191-192, 248-250), its
// gradient is returned as d_period and chained through those torch ops by the caller.  No gradient flows through the quantile
// mask (it only enters a comparison), masked entries have weight exactly 0.
//
// One warp per (sample, row): lanes sweep the columns twice (delta first, then the scatter), every kept pair costs a D-long dot
// product.  This is a correctness path (rarely requested), not a tuned one.
#pragma once
#include "geometry.cuh"

namespace pit {

constexpr int CG_WARPS = 4;

struct CoordGradParams {
  const float* mesh_out;  // [(B),N,sd]
  const float* mesh_in;   // [(B),M,sd]
  const float* period;    // device scalar or null
  const float* values;    // [B,M,D]
  const float* scale;     // [H]
  const float* v_min;     // [(B),N]
  const float* v_lo;
  const float* v_hi;
  const float* rowsum;    // [(B),H,N]
  const float* d_out;     // rows addressed with ld_out / col_off
  float weight;
  int masked, B, H, N, M, D, sd, mesh_batched;
  int64_t ld_out, col_off;
  float* d_mesh_out;  // [(B),N,sd] zero-initialised
  float* d_mesh_in;   // [(B),M,sd] zero-initialised
  float* d_period;    // [1] zero-initialised (periodic variants) or null
};

// per-axis factor of d d2 / d xo and of d d2 / d l for one coordinate pair
template <int GEO>
__device__ __forceinline__ void cg_axis(float xo, float xi, float period, float& g_xo, float& g_l) {
  const float raw = __fsub_rn(xo, xi);
  if (GEO == GEO_EUCLID1 || GEO == GEO_EUCLID2) {
    g_xo = 2.f * raw;
    g_l = 0.f;
  } else {
    const float m = fabsf(raw), alt = __fsub_rn(period, m);
    const float sgn = raw > 0.f ? 1.f : (raw < 0.f ? -1.f : 0.f);
    const float mp = fminf(m, alt);
    const float fm = m < alt ? 1.f : (m > alt ? -1.f : 0.f);     // torch.minimum splits the gradient evenly at a tie
    const float fl = m < alt ? 0.f : (m > alt ? 1.f : 0.5f);
    g_xo = 2.f * mp * fm * sgn;
    g_l = 2.f * mp * fl;
  }
}

template <int GEO>
__global__ void __launch_bounds__(CG_WARPS * 32) coord_gradient_kernel(const CoordGradParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * CG_WARPS + (threadIdx.x >> 5);
  if (item >= (int64_t)P.B * P.N) return;
  const int b = (int)(item / P.N), i = (int)(item - (int64_t)b * P.N);
  const int bm = P.mesh_batched ? b : 0;
  constexpr bool TWO = (GEO == GEO_EUCLID2 || GEO == GEO_PERIODIC2);
  const float period = P.period ? __ldg(P.period) : 0.f;
  const float* mo = P.mesh_out + (int64_t)bm * P.N * P.sd;
  const float* mi = P.mesh_in + (int64_t)bm * P.M * P.sd;
  const Point<GEO> o = load_point<GEO>(mo, i, P.sd);
  const int64_t srow = (int64_t)bm * P.N + i;
  const float v_min = __ldg(P.v_min + srow);
  const float v_lo = P.masked ? __ldg(P.v_lo + srow) : 0.f, v_hi = P.masked ? __ldg(P.v_hi + srow) : 0.f;
  float gx = 0.f, gy = 0.f, gl = 0.f;
  for (int h = 0; h < P.H; ++h) {
    const float s = __ldg(P.scale + h);
    const float top = __fmul_rn(v_min, s);
    const float cut = P.masked ? head_threshold(v_lo, v_hi, s, P.weight) : 0.f;
    const float inv_l = 1.f / __ldg(P.rowsum + ((int64_t)bm * P.H + h) * P.N + i);
    const float* go = P.d_out + ((int64_t)b * P.N + i) * P.ld_out + P.col_off + (int64_t)h * P.D;
    float delta = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = lane; j < P.M; j += 32) {
        const Point<GEO> q = load_point<GEO>(mi, j, P.sd);
        const float sc = __fmul_rn(dist2<GEO>(o, q, period), s);
        if (P.masked && !(sc <= cut)) continue;
        const float p = expf(__fsub_rn(top, sc)) * inv_l;
        if (p == 0.f) continue;
        const float* u = P.values + ((int64_t)b * P.M + j) * P.D;
        float dp = 0.f;
        for (int d = 0; d < P.D; ++d) dp = fmaf(__ldg(go + d), __ldg(u + d), dp);
        if (pass == 0) {
          delta = fmaf(p, dp, delta);
        } else {
          const float coef = -s * p * (dp - delta);   // dL / d d2_ij
          float ax, al, ay = 0.f, bl = 0.f;
          cg_axis<GEO>(o.x, q.x, period, ax, al);
          if (TWO) cg_axis<GEO>(o.y, q.y, period, ay, bl);
          gx = fmaf(coef, ax, gx);
          gl = fmaf(coef, al + bl, gl);
          atomicAdd(P.d_mesh_in + ((int64_t)bm * P.M + j) * P.sd, -coef * ax);
          if (TWO) {
            gy = fmaf(coef, ay, gy);
            atomicAdd(P.d_mesh_in + ((int64_t)bm * P.M + j) * P.sd + 1, -coef * ay);
          }
        }
      }
      if (pass == 0) delta = warp_sum(delta);
    }
  }
  gx = warp_sum(gx);
  gy = warp_sum(gy);
  gl = warp_sum(gl);
  if (lane == 0) {
    atomicAdd(P.d_mesh_out + srow * P.sd, gx);
    if (TWO) atomicAdd(P.d_mesh_out + srow * P.sd + 1, gy);
    if (P.d_period && gl != 0.f) atomicAdd(P.d_period, gl);
  }
}

}  // namespace pit
