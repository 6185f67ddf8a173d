// K1: exact SIMT position-attention (forward, scale-gradient pass, value-gradient pass).
//
// One warp owns one output row of one head.  Its lanes sweep the reduction index 32 entries at
// a time: each lane recomputes one squared distance from the coordinates (bit-exact, see
// geometry.cuh), applies the per-head scale, the quantile cut and the soft-max shift that is
// known in advance (the row minimum), and a ballot compacts the few kept entries (2 % of the
// row in the local encoder/decoder stages).  For every kept entry the whole warp gathers the
// matching value row with 128-bit loads and accumulates into registers.  The N x M weight
// matrix, its mask and its soft-max never exist in memory (pit.py:46-57 materialises all three).
//
// The same sweep serves three kernels:
//   posatt_fwd_kernel   out[b,i,h,:]   = sum_j P_ij U[b,j,:] / l_i                (pit.py:52-57)
//   posatt_dscale_kernel dL/ds_h rows  = -(1/l) sum_e dO_e (W_e - (m/l) O_e),  W = sum_j P d2 U,  m = sum_j P d2
//   posatt_dvalues_kernel dU[b,j,:]    = sum_h sum_i P_ij/l_i dO[b,i,h,:]   (roles of rows/columns swapped)
// Fixed meshes share P across the batch, so a row's value vector is the B*D-wide concatenation
// over samples; per-sample meshes use D-wide vectors.  Vectors wider than one register tile
// (32 lanes x VEC x A floats) are split over blockIdx.y; rows that are too few to fill the
// GPU are split along the reduction index over blockIdx.z and combined with fp32 REDs (the
// soft-max shift is known, so partial sums need no rescaling).
#pragma once
#include "geometry.cuh"

namespace pit {

constexpr int WARPS_PER_BLOCK = 4;
constexpr int SWEEP = 4;  // 32-entry groups each lane evaluates per trip (instruction-level parallelism)

struct AttnParams {
  const float* mesh_out;  // [(B),N,sd]
  const float* mesh_in;   // [(B),M,sd]
  const float* period;    // device scalar or null
  const float* values;    // [B,M,D]
  const float* scale;     // [H]
  const float* v_min;     // [(B),N]
  const float* v_lo;      // [(B),N] or null
  const float* v_hi;      // [(B),N] or null
  float weight;
  int masked;
  int B, H, N, M, D, sd, mesh_batched;
  int width;              // value-vector width per row: mesh_batched ? D : B*D
  int split_len;          // reduction entries per blockIdx.z slice (multiple of 32)
  // forward
  float* out;
  int64_t ld_out, col_off;
  float* rowsum;          // [(B),H,N]
  float* partial;         // split accumulation buffer [items, width] (forward, n_split > 1)
  // backward
  const float* d_out;
  float* d_values;        // [B,M,D]
  float* dscale_terms;    // [items,3]: A = sum dO.W, Bq = sum dO.O(unnormalised), m = sum P d2
  int add_concat;
};

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
  float v[1];
};
template <>
struct Vec<4> {
  float v[4];
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> load_vec(const float* p) {
  Vec<VEC> r;
  if (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x;
    r.v[VEC > 1 ? 1 : 0] = t.y;
    r.v[VEC > 1 ? 2 : 0] = t.z;
    r.v[VEC > 1 ? 3 : 0] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const Vec<VEC>& r) {
  if (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[VEC > 1 ? 1 : 0], r.v[VEC > 1 ? 2 : 0], r.v[VEC > 1 ? 3 : 0]);
  } else {
    *p = r.v[0];
  }
}

template <int VEC>
__device__ __forceinline__ void red_vec(float* p, const Vec<VEC>& r) {
#pragma unroll
  for (int v = 0; v < VEC; ++v) atomicAdd(p + v, r.v[v]);
}

// Element layout of one warp's register tile: lane owns A groups of VEC consecutive floats,
// group a starts at element  e = chunk*32*VEC*A + (a*32 + lane)*VEC  of the row's value vector.
template <int VEC, int A>
struct Tile {
  bool ok[A];
  int64_t val_off[A];  // offset of the group inside `values` for reduction entry 0
  int64_t out_off[A];  // offset of the group inside out / d_out for this row and head
};

// Fills the tile offsets for a row of the (row i, head h) kind used by forward and dscale.
template <int VEC, int A>
__device__ __forceinline__ void make_row_tile(Tile<VEC, A>& t, const AttnParams& P, int bm, int i, int h, int chunk, int lane) {
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const int e = (chunk * A + a) * 32 * VEC + lane * VEC;
    t.ok[a] = e < P.width;
    const int b = P.mesh_batched ? bm : e / P.D;
    const int d = P.mesh_batched ? e : e % P.D;
    t.val_off[a] = (int64_t)b * P.M * P.D + d;
    t.out_off[a] = ((int64_t)b * P.N + i) * P.ld_out + P.col_off + (int64_t)h * P.D + d;
  }
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <int GEO, int VEC, int A>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) posatt_fwd_kernel(const AttnParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int rows_total = (P.mesh_batched ? P.B : 1) * P.N;
  if (item >= (int64_t)rows_total * P.H) return;
  const int h = (int)(item % P.H);
  const int row = (int)(item / P.H);
  const int bm = P.mesh_batched ? row / P.N : 0;
  const int i = row - bm * P.N;
  const int chunk = blockIdx.y;
  const bool split = gridDim.z > 1;

  const float* mesh_in = P.mesh_in + (int64_t)bm * P.M * P.sd;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const Point<GEO> o = load_point<GEO>(P.mesh_out, row, P.sd);
  const float s = __ldg(P.scale + h);
  const float top = __fmul_rn(__ldg(P.v_min + row), s);  // -max logit = fl(s*v_min)
  const float cut = P.masked ? head_threshold(__ldg(P.v_lo + row), __ldg(P.v_hi + row), s, P.weight) : INFINITY;

  Tile<VEC, A> t;
  make_row_tile(t, P, bm, i, h, chunk, lane);
  Vec<VEC> acc[A];
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[a].v[v] = 0.f;
  float lsum = 0.f;

  const int j_begin = blockIdx.z * P.split_len;
  const int j_end = min(P.M, j_begin + P.split_len);
  for (int j0 = j_begin; j0 < j_end; j0 += 32 * SWEEP) {
    float p[SWEEP];
#pragma unroll
    for (int u = 0; u < SWEEP; ++u) {
      const int j = j0 + u * 32 + lane;
      p[u] = 0.f;
      if (j < j_end) {
        const float sc = __fmul_rn(dist2<GEO>(o, load_point<GEO>(mesh_in, j, P.sd), period), s);
        if (sc <= cut) p[u] = expf(__fsub_rn(top, sc));
      }
      lsum += p[u];
    }
#pragma unroll
    for (int u = 0; u < SWEEP; ++u) {
      // underflowed weights are exact zeros in the reference too: skip them
      unsigned todo = __ballot_sync(FULL, p[u] > 0.f);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const float pj = __shfl_sync(FULL, p[u], src);
        const float* vrow = P.values + (int64_t)(j0 + u * 32 + src) * P.D;
#pragma unroll
        for (int a = 0; a < A; ++a) {
          if (t.ok[a]) {
            const Vec<VEC> uv = load_vec<VEC>(vrow + t.val_off[a]);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[a].v[v] = fmaf(pj, uv.v[v], acc[a].v[v]);
          }
        }
      }
    }
  }
  lsum = warp_sum(lsum);
  const int64_t stat_idx = ((int64_t)bm * P.H + h) * P.N + i;
  if (!split) {
    if (chunk == 0 && lane == 0) P.rowsum[stat_idx] = lsum;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      if (t.ok[a]) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[a].v[v] = acc[a].v[v] / lsum;
        store_vec<VEC>(P.out + t.out_off[a], acc[a]);
      }
    }
  } else {
    if (chunk == 0 && lane == 0) atomicAdd(P.rowsum + stat_idx, lsum);
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const int e = (chunk * A + a) * 32 * VEC + lane * VEC;
      if (t.ok[a]) red_vec<VEC>(P.partial + item * P.width + e, acc[a]);
    }
  }
}

// Split forward: out = partial / rowsum.  One thread per (item, element).
static __global__ void posatt_fwd_finalize_kernel(const AttnParams P) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows_total = (P.mesh_batched ? P.B : 1) * P.N;
  const int64_t total = (int64_t)rows_total * P.H * P.width;
  if (idx >= total) return;
  const int e = (int)(idx % P.width);
  const int64_t item = idx / P.width;
  const int h = (int)(item % P.H);
  const int row = (int)(item / P.H);
  const int bm = P.mesh_batched ? row / P.N : 0;
  const int i = row - bm * P.N;
  const int b = P.mesh_batched ? bm : e / P.D;
  const int d = P.mesh_batched ? e : e % P.D;
  const float l = P.rowsum[((int64_t)bm * P.H + h) * P.N + i];
  P.out[((int64_t)b * P.N + i) * P.ld_out + P.col_off + (int64_t)h * P.D + d] = P.partial[idx] / l;
}

// ---------------------------------------------------------------------------------------
// backward, scale gradient:  per row  A = sum_e dO_e W_e,  Bq = sum_e dO_e O_e,  m = sum_j P d2
// ---------------------------------------------------------------------------------------
template <int GEO, int VEC, int A>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) posatt_dscale_kernel(const AttnParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int rows_total = (P.mesh_batched ? P.B : 1) * P.N;
  if (item >= (int64_t)rows_total * P.H) return;
  const int h = (int)(item % P.H);
  const int row = (int)(item / P.H);
  const int bm = P.mesh_batched ? row / P.N : 0;
  const int i = row - bm * P.N;
  const int chunk = blockIdx.y;

  const float* mesh_in = P.mesh_in + (int64_t)bm * P.M * P.sd;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const Point<GEO> o = load_point<GEO>(P.mesh_out, row, P.sd);
  const float s = __ldg(P.scale + h);
  const float top = __fmul_rn(__ldg(P.v_min + row), s);
  const float cut = P.masked ? head_threshold(__ldg(P.v_lo + row), __ldg(P.v_hi + row), s, P.weight) : INFINITY;

  Tile<VEC, A> t;
  make_row_tile(t, P, bm, i, h, chunk, lane);
  Vec<VEC> acc[A], accw[A];
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[a].v[v] = accw[a].v[v] = 0.f;
  float msum = 0.f;

  const int j_begin = blockIdx.z * P.split_len;
  const int j_end = min(P.M, j_begin + P.split_len);
  for (int j0 = j_begin; j0 < j_end; j0 += 32 * SWEEP) {
    float p[SWEEP], pd[SWEEP];
#pragma unroll
    for (int u = 0; u < SWEEP; ++u) {
      const int j = j0 + u * 32 + lane;
      p[u] = 0.f;
      pd[u] = 0.f;
      if (j < j_end) {
        const float d2 = dist2<GEO>(o, load_point<GEO>(mesh_in, j, P.sd), period);
        const float sc = __fmul_rn(d2, s);
        if (sc <= cut) p[u] = expf(__fsub_rn(top, sc));
        pd[u] = p[u] * d2;
      }
      msum += pd[u];
    }
#pragma unroll
    for (int u = 0; u < SWEEP; ++u) {
      unsigned todo = __ballot_sync(FULL, p[u] > 0.f);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const float pj = __shfl_sync(FULL, p[u], src);
        const float pdj = __shfl_sync(FULL, pd[u], src);
        const float* vrow = P.values + (int64_t)(j0 + u * 32 + src) * P.D;
#pragma unroll
        for (int a = 0; a < A; ++a) {
          if (t.ok[a]) {
            const Vec<VEC> uv = load_vec<VEC>(vrow + t.val_off[a]);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              acc[a].v[v] = fmaf(pj, uv.v[v], acc[a].v[v]);
              accw[a].v[v] = fmaf(pdj, uv.v[v], accw[a].v[v]);
            }
          }
        }
      }
    }
  }
  float dot_w = 0.f, dot_o = 0.f;
#pragma unroll
  for (int a = 0; a < A; ++a) {
    if (t.ok[a]) {
      const Vec<VEC> g = load_vec<VEC>(P.d_out + t.out_off[a]);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        dot_w = fmaf(g.v[v], accw[a].v[v], dot_w);
        dot_o = fmaf(g.v[v], acc[a].v[v], dot_o);
      }
    }
  }
  dot_w = warp_sum(dot_w);
  dot_o = warp_sum(dot_o);
  msum = warp_sum(msum);
  if (lane == 0) {
    float* dst = P.dscale_terms + item * 3;
    atomicAdd(dst + 0, dot_w);
    atomicAdd(dst + 1, dot_o);
    if (chunk == 0) atomicAdd(dst + 2, msum);
  }
}

// rows[item] = -(A - (m/l) Bq) / l   with item = row * H + h (summed per head by the host-side reduction)
static __global__ void posatt_dscale_finalize_kernel(const AttnParams P, float* __restrict__ rows) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows_total = (P.mesh_batched ? P.B : 1) * P.N;
  if (item >= (int64_t)rows_total * P.H) return;
  const int h = (int)(item % P.H);
  const int row = (int)(item / P.H);
  const int bm = P.mesh_batched ? row / P.N : 0;
  const int i = row - bm * P.N;
  const float l = P.rowsum[((int64_t)bm * P.H + h) * P.N + i];
  const float* term = P.dscale_terms + item * 3;
  rows[item] = -(term[0] - (term[2] / l) * term[1]) / l;
}

// ---------------------------------------------------------------------------------------
// backward, value gradient: the warp owns column j; lanes sweep rows i of every head.
// ---------------------------------------------------------------------------------------
template <int GEO, int VEC, int A>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) posatt_dvalues_kernel(const AttnParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int cols_total = (P.mesh_batched ? P.B : 1) * P.M;
  if (item >= cols_total) return;
  const int bm = P.mesh_batched ? (int)(item / P.M) : 0;
  const int j = (int)(item - (int64_t)bm * P.M);
  const int chunk = blockIdx.y;
  const bool split = gridDim.z > 1;

  const float* mesh_out = P.mesh_out + (int64_t)bm * P.N * P.sd;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const Point<GEO> in = load_point<GEO>(P.mesh_in, item, P.sd);

  bool ok[A];
  int64_t g_off[A];   // offset inside d_out for row 0, head 0
  int64_t dv_off[A];  // offset inside d_values for this column
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const int e = (chunk * A + a) * 32 * VEC + lane * VEC;
    ok[a] = e < P.width;
    const int b = P.mesh_batched ? bm : e / P.D;
    const int d = P.mesh_batched ? e : e % P.D;
    g_off[a] = (int64_t)b * P.N * P.ld_out + P.col_off + d;
    dv_off[a] = ((int64_t)b * P.M + j) * P.D + d;
  }
  Vec<VEC> acc[A];
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[a].v[v] = 0.f;

  const int i_begin = blockIdx.z * P.split_len;
  const int i_end = min(P.N, i_begin + P.split_len);
  for (int h = 0; h < P.H; ++h) {
    const float s = __ldg(P.scale + h);
    const float* rowsum = P.rowsum + ((int64_t)bm * P.H + h) * P.N;
    for (int i0 = i_begin; i0 < i_end; i0 += 32 * SWEEP) {
      float p[SWEEP];
#pragma unroll
      for (int u = 0; u < SWEEP; ++u) {
        const int i = i0 + u * 32 + lane;
        p[u] = 0.f;
        if (i < i_end) {
          const int64_t row = (int64_t)bm * P.N + i;
          const Point<GEO> o = load_point<GEO>(mesh_out, i, P.sd);
          const float sc = __fmul_rn(dist2<GEO>(o, in, period), s);
          const float top = __fmul_rn(__ldg(P.v_min + row), s);
          const float cut = P.masked ? head_threshold(__ldg(P.v_lo + row), __ldg(P.v_hi + row), s, P.weight) : INFINITY;
          if (sc <= cut) p[u] = expf(__fsub_rn(top, sc)) / __ldg(rowsum + i);
        }
      }
#pragma unroll
      for (int u = 0; u < SWEEP; ++u) {
        unsigned todo = __ballot_sync(FULL, p[u] > 0.f);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const float pi = __shfl_sync(FULL, p[u], src);
          const float* grow = P.d_out + (int64_t)(i0 + u * 32 + src) * P.ld_out + (int64_t)h * P.D;
#pragma unroll
          for (int a = 0; a < A; ++a) {
            if (ok[a]) {
              const Vec<VEC> g = load_vec<VEC>(grow + g_off[a]);
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[a].v[v] = fmaf(pi, g.v[v], acc[a].v[v]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < A; ++a) {
    if (!ok[a]) continue;
    if (P.add_concat && blockIdx.z == 0) {  // concat pass-through: d_out[b, j, 0:D]  (N == M)
      const Vec<VEC> g = load_vec<VEC>(P.d_out + g_off[a] - P.col_off + (int64_t)j * P.ld_out);
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[a].v[v] += g.v[v];
    }
    if (split)
      red_vec<VEC>(P.d_values + dv_off[a], acc[a]);
    else
      store_vec<VEC>(P.d_values + dv_off[a], acc[a]);
  }
}

}  // namespace pit
