// K3p: fused decoder tail (pit.decoder, pit.py:124-127) driven by a cached TILE PLAN of the mesh pair.
//
// The reference re-derives the locality mask of the decoder from scratch every step -- a full row sort inside
// torch.quantile (pit.py:136) -- although the meshes never change.  tail_mma_*_kernel (decoder_tail_mma.cuh) already
// replaced the sort by order statistics, but still re-scanned all M latent points for every 16-row tile of every
// launch.  Everything about that scan that does not depend on lmda is moved here into a plan that is built ONCE per
// (mesh_out, mesh_in, locality) and cached next to the row statistics:
//
//   * a row is a CANDIDATE partner of latent point j iff d2(row, j) <= v_hi(row) * (1 + 1e-6): a superset of what any
//     head can keep (the per-head test fl(d2*s) <= T(s) is monotone in d2 and T <= fl(v_hi*s)); it depends on the
//     meshes and the locality only;
//   * the rows are SORTED by their candidate set, so 32 consecutive rows of the sorted order share (nearly) one set:
//     Darcy-421 has ~2 300 distinct sets for 177 241 rows.  A tile of 32 sorted rows therefore needs 7-9 candidates
//     (one k-step of 8) where a tile of 16 mesh-ordered rows needed 8.9 on average and two k-steps;
//   * per tile the plan stores the candidate list (int16) and the squared distances of its 32 rows to every candidate
//     (bit-exact d2, one 128-byte line per candidate); per row a 16-byte record {v_min, v_lo, v_hi, row}.
//
// With the plan, preparing a tile is a coalesced load plus the exact per-head weight evaluation (lane = row), and the
// products run on 32-row tiles:  pre^T[(b,c) x 32 rows] = Y^T[(b,c) x cand] . P^T[cand x 32 rows]  with
// mma.sync.m16n8k8 (3xTF32 as in decoder_tail_mma.cuh: hi*hi + lo*hi + hi*lo, fp32 accumulate).  A warp owns one
// sample (C hidden channels) and walks it in 32-column chunks: a thread holds 4 contiguous channels, so the A operand
// (Y^T) is ONE 128-bit load per candidate and head, and the B operand (weights, pre-split into hi/lo by the preparing
// warp) is ONE 128-bit shared-memory load per n-tile.  The mask decision itself is unchanged: bit-exact d2, per-head
// cut, fp32 weights.  Output rows leave through the row index of the record (a permutation of the mesh order).
#pragma once
#include "decoder_tail_mma.cuh"

namespace pit {

constexpr int TP_ROWS = 32;   // rows per tile = four n8 MMA tiles
constexpr int TP_KB = 16;     // candidates per weight block held in shared memory (two k-steps of 8)
constexpr int TP_MT = 2;      // m16 MMA tiles per warp pass
constexpr int TP_CHUNK = 16 * TP_MT;  // hidden-vector columns per warp pass
constexpr int TP_TPC = 2 * TP_MT;     // contiguous columns a thread owns inside a chunk (one float4)
constexpr int TP_MAX_WARPS = 8;       // forward: one warp per sample
constexpr int TP_BWD_MAX_WARPS = 16;  // backward: one warp per (sample, 32-column chunk), one CTA per SM
constexpr int TP_BWD_ROUND = 2;   // tiles prepared per round in the backward
constexpr int TP_FWD_CTAS = 2;    // resident forward CTAs per SM the register budget is set for (2: 128 registers, 3: 80)

// Device view of a plan (all pointers into caller-owned buffers, see pit_tail_plan_t in include/pit_posatt.h).
struct TailPlanDev {
  const float4* rec;        // [n_tiles*32] {v_min, v_lo, v_hi, row index as int bits} of the sorted rows; row = -1: padding
  const int32_t* tile_off;  // [n_tiles+1] offsets into cand, multiples of 8 (every list starts on a 16-byte boundary)
  const int32_t* tile_cnt;  // [n_tiles] candidates of the tile
  const int16_t* cand;      // candidate columns of every tile, ascending inside a tile
  const float* d2;          // [tile_off[n_tiles]][32] squared distance of tile row r to candidate k at d2[(tile_off[t] + k)*32 + r]
  int n_tiles;
  int tiles_per_cta;
  int round;  // forward: tiles prepared per round (<= 8)
};

// ---------------------------------------------------------------------------------------
// plan construction (once per mesh pair)
// ---------------------------------------------------------------------------------------
// The plan tiles the LARGE mesh of a stage against the small one: `mesh_out` below is the tiled side (N points, sorted and
// cut into tiles of 32), `mesh_in` the small side (M <= 1024 points, the candidates).  For a decoder these are the stage's
// own mesh_out / mesh_in and the pre-filter radius belongs to the tiled point (its row statistics); for an encoder
// (`transposed`) the tiled side is the stage's mesh_in (the attention's columns), the candidates are its rows, and the
// radius -- like all statistics -- belongs to the candidate.
struct PlanBuildParams {
  const float* mesh_out;  // [N,sd] tiled side
  const float* mesh_in;   // [M,sd] candidate side
  const float* period;
  const float* v_min;
  const float* v_lo;
  const float* v_hi;
  int masked;
  int transposed;
  int N, M, sd;
  // stage 1
  unsigned long long* keys;  // [N]
  int32_t* rows;             // [N] identity
  // stage 2 / 3
  const int32_t* perm;  // [N] rows sorted by key
  int32_t* tile_cnt;    // [n_tiles] candidate count per tile (count pass)
  int32_t* tile_pad;    // [n_tiles+1] the count rounded up to a multiple of 8 (scanned into tile_off); entry n_tiles is zeroed
  const int32_t* tile_off;
  float4* rec;
  int16_t* cand;
  float* d2;
  int n_tiles;
};

template <int GEO>
__device__ __forceinline__ float plan_vcap(const PlanBuildParams& P, int row) {
  return P.masked ? __ldg(P.v_hi + row) * 1.000001f : INFINITY;  // same pre-filter radius as tm_row / tall_scan_row
}
// d2 in the argument order of the stage (mesh_out point first), whichever side is tiled
template <int GEO>
__device__ __forceinline__ float plan_dist2(const PlanBuildParams& P, const Point<GEO>& tiled, const Point<GEO>& cand, float period) {
  return P.transposed ? dist2<GEO>(cand, tiled, period) : dist2<GEO>(tiled, cand, period);
}

// Sort key of a row: its four smallest candidate columns (10 bits each) and a hash of the whole set.  Equal sets get
// equal keys; sets that share their leading columns stay close, which keeps the tiles that straddle two sets small.
template <int GEO, int CPL>
__global__ void __launch_bounds__(128) plan_key_kernel(const PlanBuildParams P) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= P.N) return;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const Point<GEO> o = load_point<GEO>(P.mesh_out, row, P.sd);
  const float vcap = P.transposed ? 0.f : plan_vcap<GEO>(P, row);
  unsigned long long lead = 0;
  int found = 0;
  uint32_t hash = 2166136261u;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    const int jc = j < P.M ? j : 0;
    const Point<GEO> q = load_point<GEO>(P.mesh_in, jc, P.sd);
    const float cap = P.transposed ? plan_vcap<GEO>(P, jc) : vcap;
    unsigned m = __ballot_sync(FULL, j < P.M && plan_dist2<GEO>(P, o, q, period) <= cap);
    hash = (hash ^ m) * 16777619u;
    hash ^= hash >> 15;
    while (m && found < 4) {
      const int b = __ffs(m) - 1;
      lead = (lead << 10) | (unsigned)(c * 32 + b);
      m &= m - 1;
      ++found;
    }
  }
  for (; found < 4; ++found) lead = (lead << 10) | 1023u;
  if (lane == 0) {
    P.keys[row] = (lead << 24) | (hash & 0xffffffu);
    P.rows[row] = row;
  }
}

// One warp per tile of 32 sorted rows: union of the rows' candidate columns.  FILL = false counts, FILL = true writes the
// list at tile_off[tile] and the row records.
template <int GEO, int CPL, bool FILL>
__global__ void __launch_bounds__(128) plan_tile_kernel(const PlanBuildParams P) {
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile >= P.n_tiles) {
    if (!FILL && tile == P.n_tiles && lane == 0) P.tile_pad[tile] = 0;
    return;
  }
  const float period = P.period ? __ldg(P.period) : 0.f;
  const int slot = tile * TP_ROWS + lane;
  const int row = slot < P.N ? __ldg(P.perm + slot) : -1;
  const int r = row >= 0 ? row : 0;
  const Point<GEO> o = load_point<GEO>(P.mesh_out, r, P.sd);
  const float vcap = row < 0 ? -1.f : (P.transposed ? INFINITY : plan_vcap<GEO>(P, r));
  if (FILL) {
    float4 rec = make_float4(0.f, 0.f, 0.f, __int_as_float(row));
    if (!P.transposed) {
      rec.x = __ldg(P.v_min + r);
      rec.y = P.masked ? __ldg(P.v_lo + r) : 0.f;
      rec.z = P.masked ? __ldg(P.v_hi + r) : 0.f;
    }
    P.rec[slot] = rec;
  }
  Point<GEO> col[CPL];
  float ccap[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    const int jc = j < P.M ? j : 0;
    col[c] = load_point<GEO>(P.mesh_in, jc, P.sd);
    ccap[c] = P.transposed ? plan_vcap<GEO>(P, jc) : INFINITY;
  }
  uint32_t flags = 0;
  for (int i = 0; i < TP_ROWS; ++i) {
    Point<GEO> oi;
    oi.x = __shfl_sync(FULL, o.x, i);
    oi.y = __shfl_sync(FULL, o.y, i);
    const float vc = __shfl_sync(FULL, vcap, i);
#pragma unroll
    for (int c = 0; c < CPL; ++c) flags |= (plan_dist2<GEO>(P, oi, col[c], period) <= fminf(vc, ccap[c])) ? (1u << c) : 0u;
  }
  int n = 0;
  const unsigned lt = (1u << lane) - 1u;
  const int off = FILL ? __ldg(P.tile_off + tile) : 0;
  int16_t* out = FILL ? P.cand + off : nullptr;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    const bool f = ((flags >> c) & 1u) && j < P.M;
    const unsigned m = __ballot_sync(FULL, f);
    if (FILL && f) out[n + __popc(m & lt)] = (int16_t)j;
    n += __popc(m);
  }
  if (!FILL && lane == 0) {
    P.tile_cnt[tile] = n;
    P.tile_pad[tile] = (n + 7) & ~7;
  }
  if (FILL) {
    if (lane < ((n + 7) & ~7) - n) out[n + lane] = 0;  // padding entries (never read as candidates: the count bounds every loop)
    // the tile's squared distances, bit-exact as the kernels would evaluate them: lane = row, one 128-byte line per candidate
    __syncwarp();
    for (int k = 0; k < n; ++k) {
      const Point<GEO> q = load_point<GEO>(P.mesh_in, (int)out[k], P.sd);
      P.d2[(size_t)(off + k) * TP_ROWS + lane] = plan_dist2<GEO>(P, o, q, period);
    }
  }
}

// ---------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 of sm_100: two fp32 lanes per instruction and issue slot) and the
// exact-GELU epilogue built on it
// ---------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 dup2(float a) { return pk2(a, a); }

// Exact (erf) GELU with ONE special-function instruction per element.  With a = |x|, z = a / sqrt(2):
//   erfc(z) = exp(-z^2) erfcx(z),   erfcx(z) ~ Q(a): degree-10 polynomial in a, fitted (weighted minimax, weight
//   exp(-z^2)) so that |exp(-z^2) (Q - erfcx)| <= 1.4e-7 for every a >= 0  (Abramowitz-Stegun 7.1.26, used by the other
//   tail kernels, has 1.5e-7 and needs a reciprocal on top of the exponential);
//   Phi(x) = 1 - erfc(z)/2 (x > 0),  erfc(z)/2 (x < 0)      =>  gelu(x) = max(x, 0) - (a Q(a) / 2) exp(-x^2 / 2)
//   gelu'(x) = Phi(x) + x exp(-x^2/2) / sqrt(2 pi).
// Measured in fp32 against erf in fp64 on [-10, 10]: |gelu error| <= 1.5e-7 max(1, |x|), |gelu' error| <= 1.9e-7
// (scripts/gelu_fit.py reproduces the fit and these figures).  TPG_COEFFS[k] = -Q_k / 2.
#define TPG_COEFFS                                                                                                              \
  {-0.49999993418f, 0.39893651669f, -0.24991680755f, 0.13251384598f, -0.06115497274f, 0.024314022073f, -0.007939450696f,       \
   0.0019746516026f, -0.0003397443039f, 3.5208560222e-05f, -1.6317331320e-06f}

// u[i] = -(Q(a)/2) exp(-x^2/2) = -erfc(|x|/sqrt 2)/2 for the two lanes of each x[i]; e[i] = exp(-x^2/2); a[i] = |x|.
// N independent Horner chains advance together (coefficient loop outermost), which is what hides the FFMA2 latency.
template <int N>
__device__ __forceinline__ void tpg_core(const f32x2 (&x)[N], f32x2 (&u)[N], f32x2 (&e)[N], f32x2 (&a)[N]) {
  constexpr float c[11] = TPG_COEFFS;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float x0, x1, g0, g1, e0, e1;
    unpk2(x[i], x0, x1);
    const f32x2 arg = mul2(mul2(x[i], x[i]), dup2(-0.72134752044448170368f));  // -x^2 / 2 * log2(e)
    unpk2(arg, g0, g1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
    e[i] = pk2(e0, e1);
    a[i] = pk2(fabsf(x0), fabsf(x1));
    u[i] = fma2(dup2(c[10]), a[i], dup2(c[9]));
  }
#pragma unroll
  for (int k = 8; k >= 0; --k)
#pragma unroll
    for (int i = 0; i < N; ++i) u[i] = fma2(u[i], a[i], dup2(c[k]));
#pragma unroll
  for (int i = 0; i < N; ++i) u[i] = mul2(u[i], e[i]);
}
// gelu of N pairs
template <int N>
__device__ __forceinline__ void tpg_gelu2(const f32x2 (&x)[N], f32x2 (&g)[N]) {
  f32x2 u[N], e[N], a[N];
  tpg_core<N>(x, u, e, a);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float x0, x1;
    unpk2(x[i], x0, x1);
    g[i] = fma2(u[i], a[i], pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  }
}
// gelu and gelu' of N pairs
template <int N>
__device__ __forceinline__ void tpg_gelu_pair2(const f32x2 (&x)[N], f32x2 (&g)[N], f32x2 (&dg)[N]) {
  f32x2 u[N], e[N], a[N];
  tpg_core<N>(x, u, e, a);  // u = -erfc/2 in [-0.5, 0)
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float x0, x1, w0, w1;
    unpk2(x[i], x0, x1);
    unpk2(add2(u[i], dup2(0.5f)), w0, w1);  // 1/2 - erfc/2 >= 0
    const f32x2 phi = add2(pk2(copysignf(w0, x0), copysignf(w1, x1)), dup2(0.5f));
    g[i] = mul2(x[i], phi);
    dg[i] = fma2(mul2(x[i], dup2(0.3989422804014327f)), e[i], phi);
  }
}

// ---------------------------------------------------------------------------------------
// shared pieces of the forward and backward kernels
// ---------------------------------------------------------------------------------------
// Weight block in shared memory, MMA-fragment order: for head h, k-step ks (8 candidates) and tile row r one 64-byte
// line of four float4 {hi(k=t), hi(k=t+4), lo(k=t), lo(k=t+4)}, t = 0..3 -- the B fragments of both split terms of lane
// (g = r & 7, t) for n-tile r >> 3 as two aligned register pairs, fetched with one LDS.128.  The float4 index inside a
// line is XOR-ed with (r >> 1) & 3 so that the 32 lanes of the preparing warp (lane = row) spread over the banks.
__device__ __forceinline__ int tp_p4_index(int h, int ks, int r, int t) { return ((h * (TP_KB / 8) + ks) * TP_ROWS + r) * 4 + (t ^ ((r >> 1) & 3)); }

// Backward: P^ and Z = P^ (d2 - m), both split, transposed for the dY / dZ products (k = tile row, n = candidate): for
// head h and candidate k one line of 16 row pairs x 8 floats {ph(r0), ph(r1), pl(r0), pl(r1), zh(r0), zh(r1), zl(r0), zl(r1)},
// r0 = 2 rp, r1 = r0 + 1 -- lane (g, t) fetches row pair 4nt + t of candidate g with two LDS.128.  132-float pitch:
// the 8 lanes of a quarter warp (two candidates x four row pairs) hit distinct 16-byte bank groups.
constexpr int TP_ZP = 132;
__device__ __forceinline__ int tp_pz_index(int h, int k, int r) { return (h * TP_KB + k) * TP_ZP + (r >> 1) * 8 + (r & 1); }

template <int NH, bool BWD>
struct TpTile {
  float4 p4[NH * (TP_KB / 8) * TP_ROWS * 4];
  float pz[BWD ? NH * TP_KB * TP_ZP : 4];
  float inv_l[NH][TP_ROWS];
  float m[BWD ? NH : 1][BWD ? TP_ROWS : 4];  // backward: sum_j P^ d2 of each row
  int row[TP_ROWS];                         // mesh row of each tile row (-1: padding)
  int cnt, off;
  int pad[2];
};

// Per-row constants of the preparing warp (lane = tile row).
template <int NH>
struct TpRow {
  float top[NH], cut[NH];
  int row;
};

template <int NH>
__device__ __forceinline__ TpRow<NH> tp_row_of(const TailParams& P, const float4 a, const float (&s)[NH]) {
  TpRow<NH> R;
  R.row = __float_as_int(a.w);
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    R.top[h] = __fmul_rn(a.x, s[h]);
    R.cut[h] = R.row < 0 ? -INFINITY : (P.masked ? head_threshold(a.y, a.z, s[h], P.weight) : INFINITY);
  }
  return R;
}
template <int NH>
__device__ __forceinline__ TpRow<NH> tp_row(const TailParams& P, const TailPlanDev& V, int tile, int lane, const float (&s)[NH]) {
  return tp_row_of<NH>(P, __ldg(V.rec + (size_t)tile * TP_ROWS + lane), s);
}

// Inputs of a tile that is prepared in the NEXT round, fetched with cp.async while the current round computes: phase 1
// then starts from shared memory instead of paying three dependent trips to L2 (offsets -> lists -> distances).
template <int NH, bool BWD>
struct TpStage {
  float4 rec[TP_ROWS];
  float d2[TP_KB][TP_ROWS];               // first 16 candidates (longer lists: the rest is read from global memory)
  float rs[BWD ? 2 * NH : 1][TP_ROWS];    // backward: row sums l, then m = sum_j P^ d2 (saved by the forward in tile order)
  int16_t cand[TP_KB];
  int cnt, off;
  int pad[2];
};

__device__ __forceinline__ void tp_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void tp_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tp_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One warp queues the inputs of `tile` (rows_sorted = n_tiles * 32: pitch of the row-sum planes).
template <int NH, bool BWD>
__device__ __forceinline__ void tp_stage_issue(TpStage<NH, BWD>* G, const TailParams& P, const TailPlanDev& V, int tile, int lane) {
  const int off = __ldg(V.tile_off + tile), cnt = __ldg(V.tile_cnt + tile);
  if (lane == 0) {
    G->cnt = cnt;
    G->off = off;
  }
  tp_cp16(&G->rec[lane], V.rec + (size_t)tile * TP_ROWS + lane);
  const int n = min(cnt, TP_KB);
  for (int c = lane; c < n * 8; c += 32) tp_cp16(&G->d2[0][0] + c * 4, V.d2 + (size_t)off * TP_ROWS + c * 4);
  if (lane < 2 && lane * 8 < n) tp_cp16(&G->cand[lane * 8], V.cand + off + lane * 8);  // every list is padded to a multiple of 8
  if (BWD) {
    const size_t plane = (size_t)V.n_tiles * TP_ROWS;
    for (int c = lane; c < 2 * NH * 8; c += 32) tp_cp16(&G->rs[c >> 3][(c & 7) * 4], P.rowsum + (c >> 3) * plane + (size_t)tile * TP_ROWS + (c & 7) * 4);
  }
  tp_cp_commit();
}

// Unnormalised weights of a candidate at squared distance d2 for this lane's row: exp(s*v_min - s*d2) if kept by head h,
// else exactly 0 (the per-head test on the rounded product, as everywhere).
template <int NH>
__device__ __forceinline__ void tp_weights(const TpRow<NH>& R, const float (&s)[NH], float d2, float (&p)[NH]) {
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float sc = __fmul_rn(d2, s[h]);
    p[h] = (sc <= R.cut[h]) ? __expf(__fsub_rn(R.top[h], sc)) : 0.f;
  }
}

// Candidates ks*8 + t and ks*8 + t + 4 of a weight block (d2a / d2b: their squared distances to this lane's row, or < 0
// for the zero fill past the end of the list): P^ = p * post in fragment order and, backward, the transposed split copies
// of P^ and P^ (d2 - m).  Lane = tile row.
template <int NH, bool BWD>
__device__ __forceinline__ void tp_store_pair(const TpRow<NH>& R, const float (&s)[NH], float d2a, float d2b, int ks, int t, int lane,
                                              TpTile<NH, BWD>* T, const float (&post)[NH], const float (&m)[NH]) {
  float pn[2][NH];
  const float d2[2] = {d2a, d2b};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    tp_weights<NH>(R, s, d2[e], pn[e]);
#pragma unroll
    for (int h = 0; h < NH; ++h) pn[e][h] = d2[e] < 0.f ? 0.f : pn[e][h] * post[h];
  }
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float h0 = tm_round_hi(pn[0][h]), h1 = tm_round_hi(pn[1][h]);
    T->p4[tp_p4_index(h, ks, lane, t)] = make_float4(h0, h1, pn[0][h] - h0, pn[1][h] - h1);
    if (BWD) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float z = pn[e][h] * (d2[e] - m[h]);
        const float ph = e ? h1 : h0, zh = tm_round_hi(z);
        float* q = &T->pz[tp_pz_index(h, ks * 8 + t + 4 * e, lane)];
        q[0] = ph, q[2] = pn[e][h] - ph, q[4] = zh, q[6] = z - zh;
      }
    }
  }
}

// Writes one weight block (up to 16 candidates, zero-filled to a multiple of 8).  d2blk: squared distance of this lane's
// row to candidate i of the block at d2blk[i * 32] (shared-memory stage or the plan's array); n = candidates in the block.
template <int NH, bool BWD>
__device__ __forceinline__ void tp_store_block(const TpRow<NH>& R, const float (&s)[NH], const float* d2blk, int n, int lane,
                                               TpTile<NH, BWD>* T, const float (&post)[NH], const float (&m)[NH]) {
  const int ksteps = (n + 7) >> 3;
  for (int ks = 0; ks < ksteps; ++ks) {
    float d2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d2[i] = ks * 8 + i < n ? d2blk[(ks * 8 + i) * TP_ROWS] : -1.f;  // warp-uniform guard
#pragma unroll
    for (int t = 0; t < 4; ++t) tp_store_pair<NH, BWD>(R, s, d2[t], d2[t + 4], ks, t, lane, T, post, m);
  }
}

// Sums of the unnormalised weights (and of weight*d2) of this lane's row over the whole candidate list, eight independent
// loads at a time: the first 16 candidates from the stage, the rest from the plan.
template <int NH>
__device__ __forceinline__ void tp_row_sums(const TpRow<NH>& R, const float (&s)[NH], const float* d2stage, const float* d2plan, int cnt,
                                            float (&psum)[NH], float (&pdsum)[NH]) {
#pragma unroll
  for (int h = 0; h < NH; ++h) psum[h] = pdsum[h] = 0.f;
  for (int k0 = 0; k0 < cnt; k0 += 8) {
    const float* src = k0 < TP_KB ? d2stage + k0 * TP_ROWS : d2plan + (size_t)k0 * TP_ROWS;
    float d2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d2[i] = k0 + i < cnt ? src[i * TP_ROWS] : -1.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float p[NH];
      tp_weights<NH>(R, s, d2[i], p);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        p[h] = d2[i] < 0.f ? 0.f : p[h];
        psum[h] += p[h];
        pdsum[h] = fmaf(p[h], d2[i], pdsum[h]);
      }
    }
  }
}

// pre^T += Y_h^T . P^_h^T for one weight block of one tile and one 32-column chunk.
//   acc[mt][nt][e]: chunk column 4g + 2mt + (e >> 1), tile row 8nt + 2t + (e & 1).
// y_chunk points at Y[b, 0, 0, chunk column 4g]; row j of head h sits (NH*j + h)*C floats further.  p4t = the tile's
// p4 array advanced to this lane's float4 (row g of n-tile 0, swizzled t).
template <int NH>
__device__ __forceinline__ void tp_mma_block(float (&acc)[TP_MT][4][4], const float4* p4t, const int16_t* cand, int cnt, int kb,
                                             const float* y_chunk, int C, int t) {
  const int base = kb * TP_KB;
  const int ksteps = (min(cnt - base, TP_KB) + 7) >> 3;
  for (int ks = 0; ks < ksteps; ++ks) {
    const int ka = base + ks * 8 + t, kc = ka + 4;
    const int ja = ka < cnt ? (int)cand[ka] : 0, jb = kc < cnt ? (int)cand[kc] : 0;
    // A operands: two 64-bit loads per m16 tile (rows g, g+8 of tile mt = chunk columns 4g+2mt, 4g+2mt+1; k-slots t, t+4)
    float2 ya[NH][TP_MT], yb[NH][TP_MT];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float2* ra = reinterpret_cast<const float2*>(y_chunk + ((size_t)ja * NH + h) * C);
      const float2* rb = reinterpret_cast<const float2*>(y_chunk + ((size_t)jb * NH + h) * C);
#pragma unroll
      for (int mt = 0; mt < TP_MT; ++mt) {
        ya[h][mt] = __ldg(ra + mt);
        yb[h][mt] = __ldg(rb + mt);
      }
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      uint32_t ah[TP_MT][4], al[TP_MT][4];
#pragma unroll
      for (int mt = 0; mt < TP_MT; ++mt) {
        ah[mt][0] = __float_as_uint(ya[h][mt].x), ah[mt][1] = __float_as_uint(ya[h][mt].y);
        ah[mt][2] = __float_as_uint(yb[h][mt].x), ah[mt][3] = __float_as_uint(yb[h][mt].y);
        al[mt][0] = tm_trunc_lo(ya[h][mt].x), al[mt][1] = tm_trunc_lo(ya[h][mt].y);
        al[mt][2] = tm_trunc_lo(yb[h][mt].x), al[mt][3] = tm_trunc_lo(yb[h][mt].y);
      }
      const float4* line = p4t + (h * (TP_KB / 8) + ks) * TP_ROWS * 4;
      float4 b4[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b4[nt] = line[nt * 32];
      // split terms outermost: the three MMAs on one accumulator are eight instructions apart (HMMA latency is ~3 issue slots)
#pragma unroll
      for (int term = 0; term < 3; ++term) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t bh[2] = {__float_as_uint(b4[nt].x), __float_as_uint(b4[nt].y)};
          const uint32_t bl[2] = {__float_as_uint(b4[nt].z), __float_as_uint(b4[nt].w)};
#pragma unroll
          for (int mt = 0; mt < TP_MT; ++mt) mma_tf32_16x8x8(acc[mt][nt], term == 0 ? al[mt] : ah[mt], term == 1 ? bl : bh);
        }
      }
    }
  }
}

// Sum of v[q] over the 8 lanes that share t (lane bits 2..4); on return lane (g, t) holds in v[0] the total of
// q = 4*bit4 + 2*bit3 + bit2 of its lane index (a reduce-scatter: 7 shuffles instead of 24).
__device__ __forceinline__ int tp_reduce_scatter8(float (&v)[8], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h4 ? v[i] : v[i + 4], keep = h4 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(FULL, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h3 ? v[i] : v[i + 2], keep = h3 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(FULL, send, 8);
  }
  {
    const float send = h2 ? v[0] : v[1], keep = h2 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(FULL, send, 4);
  }
  return (h4 ? 4 : 0) + (h3 ? 2 : 0) + (h2 ? 1 : 0);
}

__host__ __device__ inline size_t tp_tile_bytes(int nh, bool bwd) {
  return nh == 1 ? (bwd ? tm_align(sizeof(TpTile<1, true>)) : tm_align(sizeof(TpTile<1, false>)))
                 : (bwd ? tm_align(sizeof(TpTile<2, true>)) : tm_align(sizeof(TpTile<2, false>)));
}
__host__ __device__ inline size_t tp_stage_bytes(int nh, bool bwd) {
  return nh == 1 ? (bwd ? tm_align(sizeof(TpStage<1, true>)) : tm_align(sizeof(TpStage<1, false>)))
                 : (bwd ? tm_align(sizeof(TpStage<2, true>)) : tm_align(sizeof(TpStage<2, false>)));
}
__host__ __device__ inline size_t tp_fwd_smem_bytes(int nh, int M, int C, int O, int round) {
  return (size_t)round * (tp_tile_bytes(nh, false) + tp_stage_bytes(nh, false) + tm_cand_bytes(M)) + tm_align((size_t)C * (1 + O) * 4);
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
// A CTA walks its tiles in rounds of `round` (<= nwarps): warp w prepares tile w of the round (phase 1), then every warp
// contracts its sample(s) against every prepared tile (phase 2).  NO as in tail_mma_fwd_kernel.
template <int NH, int NO>
__global__ void __launch_bounds__(32 * TP_MAX_WARPS, TP_FWD_CTAS) tail_plan_fwd_kernel(const TailParams P, const TailPlanDev V) {
  const int n_out = NO == 1 ? 1 : P.O;
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  using Tile = TpTile<NH, false>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int round = V.round;
  const size_t tile_stride = tm_align(sizeof(Tile));
  using Stage = TpStage<NH, false>;
  const size_t stage_stride = tm_align(sizeof(Stage));
  unsigned char* stage_base = tall_smem_raw + round * tile_stride;
  unsigned char* cand_base = stage_base + round * stage_stride;
  const size_t cand_stride = tm_cand_bytes(P.M);
  float* par = reinterpret_cast<float*>(cand_base + round * cand_stride);  // [b1 (C) | W2 (O x C)]
  const size_t plane = (size_t)V.n_tiles * TP_ROWS;  // pitch of the saved row-sum planes (tile order)

  const int g = lane >> 2, t = lane & 3;
  const int p4_lane = g * 4 + (t ^ ((g >> 1) & 3));
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  for (int i = tid; i < P.C * (1 + n_out); i += blockDim.x) par[i] = i < P.C ? __ldg(P.b1 + i) : __ldg(P.w2 + (i - P.C));
  float b2r[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) b2r[o] = o < n_out ? __ldg(P.b2 + o) : 0.f;

  const int cps = P.C / TP_CHUNK;  // chunks per sample
  const int tile_begin = blockIdx.x * V.tiles_per_cta;
  const int tile_end = min(V.n_tiles, tile_begin + V.tiles_per_cta);
  for (int pv = warp; pv < min(round, tile_end - tile_begin); pv += nwarps)
    tp_stage_issue<NH, false>(reinterpret_cast<Stage*>(stage_base + pv * stage_stride), P, V, tile_begin + pv, lane);
  for (int tb = tile_begin; tb < tile_end; tb += round) {
    const int in_round = min(round, tile_end - tb);
    __syncthreads();  // the previous round's tiles are no longer read (also orders the `par` fill before its first use)
    // ---- phase 1: one tile per warp, lane = row; its inputs were staged during the previous round ----
    tp_cp_wait();
    __syncwarp();
    for (int pv = warp; pv < in_round; pv += nwarps) {
      Tile* T = reinterpret_cast<Tile*>(tall_smem_raw + pv * tile_stride);
      const Stage* G = reinterpret_cast<const Stage*>(stage_base + pv * stage_stride);
      int16_t* cand = reinterpret_cast<int16_t*>(cand_base + pv * cand_stride);
      const int tile = tb + pv;
      const int off = G->off, cnt = G->cnt;
      for (int k = lane; k < cnt; k += 32) cand[k] = k < TP_KB ? G->cand[k] : V.cand[off + k];
      const TpRow<NH> R = tp_row_of<NH>(P, G->rec[lane], s);
      float psum[NH], pdsum[NH], inv[NH];
      tp_row_sums<NH>(R, s, &G->d2[0][lane], V.d2 + (size_t)off * TP_ROWS + lane, cnt, psum, pdsum);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        inv[h] = R.row >= 0 ? 1.f / psum[h] : 0.f;
        T->inv_l[h][lane] = inv[h];
        // saved for the backward in tile order: the row sum l and m = sum_j P^ d2
        P.rowsum[h * plane + (size_t)tile * TP_ROWS + lane] = R.row >= 0 ? psum[h] : 1.f;
        P.rowsum[(NH + h) * plane + (size_t)tile * TP_ROWS + lane] = pdsum[h] * inv[h];
      }
      tp_store_block<NH, false>(R, s, &G->d2[0][lane], min(cnt, TP_KB), lane, T, inv, inv);
      T->row[lane] = R.row;
      if (lane == 0) {
        T->cnt = cnt;
        T->off = off;
      }
    }
    __syncthreads();
    // the stages are free again: queue the next round's tiles, which land while this round computes
    for (int pv = warp; pv < min(round, tile_end - tb - round); pv += nwarps)
      tp_stage_issue<NH, false>(reinterpret_cast<Stage*>(stage_base + pv * stage_stride), P, V, tb + round + pv, lane);
    // ---- phase 2: every warp contracts its sample(s) against every prepared tile ----
    for (int v = 0; v < in_round; ++v) {
      Tile* T = reinterpret_cast<Tile*>(tall_smem_raw + v * tile_stride);
      const int16_t* cand = reinterpret_cast<const int16_t*>(cand_base + v * cand_stride);
      const int cnt = T->cnt;
      const int nkb = (cnt + TP_KB - 1) / TP_KB;
      for (int b0 = 0; b0 < P.B; b0 += nwarps) {
        const int b = b0 + warp;
        const bool active = b < P.B;
        f32x2 part[NO][4];  // part[o][nt] = rows (8nt + 2t, 8nt + 2t + 1)
#pragma unroll
        for (int o = 0; o < NO; ++o)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) part[o][nt] = 0ull;
        for (int ch = 0; ch < cps; ++ch) {
          const int c0 = ch * TP_CHUNK + TP_TPC * g;  // this thread's four hidden channels
          const float* y_chunk = P.y + (size_t)(active ? b : 0) * P.M * NH * P.C + c0;
          float acc[TP_MT][4][4];
#pragma unroll
          for (int mt = 0; mt < TP_MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
          for (int kb = 0; kb < nkb; ++kb) {
            if (nkb > 1) {  // rare: more than 16 candidates -> rebuild block kb in place (CTA-uniform branch)
              __syncthreads();
              if (warp == 0) {
                const TpRow<NH> R = tp_row<NH>(P, V, tb + v, lane, s);
                float inv[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) inv[h] = T->inv_l[h][lane];
                tp_store_block<NH, false>(R, s, V.d2 + (size_t)(T->off + kb * TP_KB) * TP_ROWS + lane, min(cnt - kb * TP_KB, TP_KB), lane, T, inv, inv);
              }
              __syncthreads();
            }
            if (active) tp_mma_block<NH>(acc, T->p4 + p4_lane, cand, cnt, kb, y_chunk, P.C, t);
          }
          if (!active) continue;
          // epilogue of the chunk: part[o][rows] += W2[o, c] gelu(b1[c] + pre[c]) over the thread's four channels, two rows per op
          const float4 b1v = *reinterpret_cast<const float4*>(par + c0);
          const float4 w2v0 = *reinterpret_cast<const float4*>(par + P.C + c0);
#pragma unroll
          for (int mt = 0; mt < TP_MT; ++mt) {
            f32x2 x[8], hid[8];  // [half][nt]: eight independent GELU chains
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int i = 2 * mt + half;
              const f32x2 bias = dup2(i == 0 ? b1v.x : i == 1 ? b1v.y : i == 2 ? b1v.z : b1v.w);
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) x[half * 4 + nt] = add2(pk2(acc[mt][nt][half * 2], acc[mt][nt][half * 2 + 1]), bias);
            }
            tpg_gelu2<8>(x, hid);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int i = 2 * mt + half;
#pragma unroll
              for (int o = 0; o < NO; ++o) {
                if (o < n_out) {
                  const float wv = o == 0 ? (i == 0 ? w2v0.x : i == 1 ? w2v0.y : i == 2 ? w2v0.z : w2v0.w) : par[(1 + o) * P.C + c0 + i];
                  const f32x2 wv2 = dup2(wv);
#pragma unroll
                  for (int nt = 0; nt < 4; ++nt) part[o][nt] = fma2(wv2, hid[half * 4 + nt], part[o][nt]);
                }
              }
            }
          }
        }
        if (!active) continue;
        // out[b, row, o] = b2[o] + sum over the 8 lanes sharing t; the reduce-scatter leaves one tile row per lane
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          if (o >= n_out) break;
          float pv8[8];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) unpk2(part[o][nt], pv8[2 * nt], pv8[2 * nt + 1]);
          const int q = tp_reduce_scatter8(pv8, lane);
          const int row = T->row[8 * (q >> 1) + 2 * t + (q & 1)];
          if (row >= 0) P.out[((int64_t)b * P.N + row) * n_out + o] = pv8[0] + b2r[o];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
struct TpBwdSmem {
  unsigned char* tiles;  // [TP_BWD_ROUND] TpTile<NH, true>
  unsigned char* stage;  // [TP_BWD_ROUND] TpStage<NH, true>
  unsigned char* cand;   // [TP_BWD_ROUND][M] int16
  float* slot_acc;       // [n_slots][NH][W], 16-byte groups rotated inside 128-byte windows by the slot index
  float* par;            // [C + O*C]: b1 then W2
  float* gpar;           // [C + O*C]: CTA-level reduction of d_b1, d_w2
  float* red;            // [TP_BWD_MAX_WARPS]
  int16_t* slot_j;       // [n_slots] slot -> column (-1: free)
  int16_t* bind;         // [TP_BWD_ROUND][M] position in the tile's candidate list -> slot (-1: unbound)
  int16_t* evict;        // [n_slots] column to flush before the slot is reused this round (-1: none)
};

__host__ __device__ inline size_t tp_bwd_smem_bytes(int nh, int M, int W, int C, int O, int n_slots) {
  return TP_BWD_ROUND * (tp_tile_bytes(nh, true) + tp_stage_bytes(nh, true) + 2 * tm_cand_bytes(M)) + tm_align((size_t)n_slots * nh * W * 4) +
         2 * tm_align((size_t)C * (1 + O) * 4) + 64 + 2 * tm_align((size_t)n_slots * 2);
}

__device__ inline TpBwdSmem tp_bwd_carve(unsigned char* p, int nh, int M, int W, int C, int O, int n_slots) {
  TpBwdSmem s{};
  s.tiles = p;
  p += TP_BWD_ROUND * tp_tile_bytes(nh, true);
  s.stage = p;
  p += TP_BWD_ROUND * tp_stage_bytes(nh, true);
  s.cand = p;
  p += TP_BWD_ROUND * tm_cand_bytes(M);
  s.bind = reinterpret_cast<int16_t*>(p);
  p += TP_BWD_ROUND * tm_cand_bytes(M);
  s.slot_acc = reinterpret_cast<float*>(p);
  p += tm_align((size_t)n_slots * nh * W * 4);
  s.par = reinterpret_cast<float*>(p);
  p += tm_align((size_t)C * (1 + O) * 4);
  s.gpar = reinterpret_cast<float*>(p);
  p += tm_align((size_t)C * (1 + O) * 4);
  s.red = reinterpret_cast<float*>(p);
  p += 64;
  s.slot_j = reinterpret_cast<int16_t*>(p);
  p += tm_align((size_t)n_slots * 2);
  s.evict = reinterpret_cast<int16_t*>(p);
  return s;
}

// Slot assignment of one round, by warp 0 between phase 1 and phase 2.  dY is accumulated in shared-memory slots bound
// to latent columns.  The cells of a slot that belong to (sample b, channels 4g..4g+3) are only ever touched by the four
// lanes of one quad of one warp, so recycling a slot needs no CTA-wide flush: the decision is published here
// (`evict[slot]` = the column whose partial sums still sit in the slot) and every quad flushes its own cells before its
// first use of the round.  Free slots are taken first, then slots that no tile of this round uses (`pinned`); a candidate
// that finds neither stays unbound (-1) and its contributions go to d_y with direct REDs.  n_slots <= 32: lane s mirrors
// slot_j[s] in a register.
__device__ __forceinline__ void tp_assign_slots(const TpBwdSmem& S, int n_slots, const int16_t* cand, int cnt, int16_t* bind, int lane,
                                                int& my_sj, uint32_t& pinned) {
  const uint32_t all = n_slots >= 32 ? 0xffffffffu : ((1u << n_slots) - 1u);
  for (int base = 0; base < cnt; base += 32) {
    const int k = base + lane;
    const int j = k < cnt ? (int)cand[k] : -1;
    int slot = -1;
    for (int sidx = 0; sidx < n_slots; ++sidx) {
      const int sj = __shfl_sync(FULL, my_sj, sidx);
      if (j >= 0 && sj == j) slot = sidx;
    }
    pinned |= __reduce_or_sync(FULL, slot >= 0 ? (1u << slot) : 0u);
    const bool want = j >= 0 && slot < 0;
    const unsigned need = __ballot_sync(FULL, want);
    if (need) {  // warp-uniform
      const uint32_t freem = __ballot_sync(FULL, lane < n_slots && my_sj < 0) & all;
      const uint32_t recyc = all & ~freem & ~pinned;
      const int rank = __popc(need & ((1u << lane) - 1u));
      const int nfree = __popc(freem);
      if (want) {
        if (rank < nfree)
          slot = (int)__fns(freem, 0, rank + 1);
        else if (rank - nfree < __popc(recyc))
          slot = (int)__fns(recyc, 0, rank - nfree + 1);
        if (slot >= 0) S.slot_j[slot] = (int16_t)j;
      }
      const uint32_t took = __reduce_or_sync(FULL, (want && slot >= 0) ? (1u << slot) : 0u);
      __syncwarp();
      if (lane < n_slots && ((took >> lane) & 1u)) {
        if (my_sj >= 0) S.evict[lane] = (int16_t)my_sj;  // a slot is taken at most once per round: it is pinned from here on
        my_sj = S.slot_j[lane];
      }
      pinned |= took;
    }
    if (k < cnt) bind[k] = (int16_t)slot;
  }
  __syncwarp();
}

// The quad's cells of slot `sidx` (sample b, hidden channels c0..c0+3 = column x of the B*C-wide vector, every head) go to
// d_y[b, j] with vector REDs and are cleared.
template <int NH>
__device__ __forceinline__ void tp_flush_cells(const TailParams& P, const TpBwdSmem& S, int W, int sidx, int j, int b, int c0, int x) {
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    float4* cell = reinterpret_cast<float4*>(S.slot_acc + ((size_t)sidx * NH + h) * W + tm_slot_pos(x, sidx));
    const float4 v = *cell;
    atomicAdd(reinterpret_cast<float4*>(P.d_y + (((size_t)b * P.M + j) * NH + h) * P.C + c0), v);
    *cell = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NH, int NO>
__global__ void __launch_bounds__(32 * TP_BWD_MAX_WARPS, 1) tail_plan_bwd_kernel(const TailParams P, const TailPlanDev V) {
  const int n_out = NO == 1 ? 1 : P.O;
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  using Tile = TpTile<NH, true>;
  const int W = P.B * P.C;
  const TpBwdSmem S = tp_bwd_carve(tall_smem_raw, NH, P.M, W, P.C, n_out, P.n_slots);
  const size_t tile_stride = tm_align(sizeof(Tile));
  const size_t cand_stride = tm_cand_bytes(P.M);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int p4_lane = g * 4 + (t ^ ((g >> 1) & 3));
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  const int n_par = P.C * (1 + n_out);
  for (int i = tid; i < n_par; i += blockDim.x) {
    S.par[i] = i < P.C ? __ldg(P.b1 + i) : __ldg(P.w2 + (i - P.C));
    S.gpar[i] = 0.f;
  }
  for (int i = tid; i < P.n_slots * NH * W; i += blockDim.x) S.slot_acc[i] = 0.f;
  for (int i = tid; i < P.n_slots; i += blockDim.x) {
    S.slot_j[i] = -1;
    S.evict[i] = -1;
  }
  int my_sj = -1;  // warp 0: register mirror of slot_j[lane]

  float ds_head[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) ds_head[h] = 0.f;
  float db2[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) db2[o] = 0.f;
  // d_b1 / d_w2 partial sums of this thread's hidden channels: channel (chunk ch, 4g + i) -> gacc[ch][i][0 (b1), 1 + o (W2 row o)]
  constexpr int TP_MAX_CPS = 2;  // C <= 64
  float gacc[TP_MAX_CPS][TP_TPC][1 + NO];
#pragma unroll
  for (int a = 0; a < TP_MAX_CPS; ++a)
#pragma unroll
    for (int i = 0; i < TP_TPC; ++i)
#pragma unroll
      for (int o = 0; o <= NO; ++o) gacc[a][i][o] = 0.f;

  const int cps = P.C / TP_CHUNK;
  const int units = P.B * cps;  // (sample, 32-column chunk) pairs; warp w takes units w, w + nwarps, ...
  const int tile_begin = blockIdx.x * V.tiles_per_cta;
  const int tile_end = min(V.n_tiles, tile_begin + V.tiles_per_cta);
  using Stage = TpStage<NH, true>;
  const size_t stage_stride = tm_align(sizeof(Stage));
  for (int pv = warp; pv < min(TP_BWD_ROUND, tile_end - tile_begin); pv += nwarps)
    tp_stage_issue<NH, true>(reinterpret_cast<Stage*>(S.stage + pv * stage_stride), P, V, tile_begin + pv, lane);
  for (int tb = tile_begin; tb < tile_end; tb += TP_BWD_ROUND) {
    const int in_round = min(TP_BWD_ROUND, tile_end - tb);
    tp_cp_wait();     // the copies this warp queued for this round have landed ...
    __syncthreads();  // ... and are visible to every warp; the previous round is finished (first round: the initialisation above)
    // ---- phase 1, spread over all warps: item (tile v, t) = candidates t, t+4 (and 8+t, 12+t) of tile v; lane = row.
    //      With l and m saved by the forward every (row, candidate) weight is independent of the others. ----
    for (int item = warp; item < in_round * 4; item += nwarps) {
      const int pv = item >> 2, tt = item & 3;
      Tile* T = reinterpret_cast<Tile*>(S.tiles + pv * tile_stride);
      const Stage* G = reinterpret_cast<const Stage*>(S.stage + pv * stage_stride);
      const int off = G->off, cnt = G->cnt;
      const TpRow<NH> R = tp_row_of<NH>(P, G->rec[lane], s);
      float inv[NH], m[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        inv[h] = R.row >= 0 ? 1.f / G->rs[h][lane] : 0.f;
        m[h] = G->rs[NH + h][lane];
      }
      const int n = min(cnt, TP_KB);
      for (int ks = 0; ks < (n + 7) >> 3; ++ks) {
        const int ka = ks * 8 + tt, kc = ka + 4;
        tp_store_pair<NH, true>(R, s, ka < n ? G->d2[ka][lane] : -1.f, kc < n ? G->d2[kc][lane] : -1.f, ks, tt, lane, T, inv, m);
      }
      if (tt == 0) {
        int16_t* cand = reinterpret_cast<int16_t*>(S.cand + pv * cand_stride);
        for (int k = lane; k < cnt; k += 32) cand[k] = k < TP_KB ? G->cand[k] : V.cand[off + k];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          T->inv_l[h][lane] = inv[h];
          T->m[h][lane] = m[h];
        }
        T->row[lane] = R.row;
        if (lane == 0) {
          T->cnt = cnt;
          T->off = off;
        }
      }
    }
    __syncthreads();
    // the stages are free again: queue the next round's tiles
    for (int pv = warp; pv < min(TP_BWD_ROUND, tile_end - tb - TP_BWD_ROUND); pv += nwarps)
      tp_stage_issue<NH, true>(reinterpret_cast<Stage*>(S.stage + pv * stage_stride), P, V, tb + TP_BWD_ROUND + pv, lane);
    if (warp == 0) {
      if (lane < P.n_slots) S.evict[lane] = -1;
      __syncwarp();
      uint32_t pinned = 0;
      for (int v = 0; v < in_round; ++v) {
        const Tile* T = reinterpret_cast<const Tile*>(S.tiles + v * tile_stride);
        tp_assign_slots(S, P.n_slots, reinterpret_cast<const int16_t*>(S.cand + v * cand_stride), T->cnt,
                        reinterpret_cast<int16_t*>(S.bind + v * (cand_stride / 2)), lane, my_sj, pinned);
      }
    }
    __syncthreads();
    // every quad first flushes its own cells of the slots that change hands in this round (lane t takes slots s = t mod 4)
    for (int sidx = t; sidx < P.n_slots; sidx += 4) {
      const int j = S.evict[sidx];
      if (j < 0) continue;
      for (int u = warp; u < units; u += nwarps) {
        const int b = u / cps, ch = u - b * cps;
        tp_flush_cells<NH>(P, S, W, sidx, j, b, ch * TP_CHUNK + TP_TPC * g, b * P.C + ch * TP_CHUNK + TP_TPC * g);
      }
    }
    __syncwarp();
    // ---- phase 2 ----
    for (int v = 0; v < in_round; ++v) {
      Tile* T = reinterpret_cast<Tile*>(S.tiles + v * tile_stride);
      const int16_t* cand = reinterpret_cast<const int16_t*>(S.cand + v * cand_stride);
      const int16_t* bind = S.bind + v * (cand_stride / 2);
      const int cnt = T->cnt;
      const int nkb = (cnt + TP_KB - 1) / TP_KB;
      for (int u0 = 0; u0 < units; u0 += nwarps) {
        const int u = u0 + warp;
        const bool active = u < units;
        const int b = active ? u / cps : 0, ch = active ? u - b * cps : 0;
        // upstream gradient of this thread's eight tile rows, as pairs (8nt + 2t, 8nt + 2t + 1)
        f32x2 go[4][NO];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r0 = T->row[8 * nt + 2 * t], r1 = T->row[8 * nt + 2 * t + 1];
#pragma unroll
          for (int o = 0; o < NO; ++o) {
            const float g0 = (active && o < n_out && r0 >= 0) ? __ldg(P.d_out + ((int64_t)b * P.N + r0) * n_out + o) : 0.f;
            const float g1 = (active && o < n_out && r1 >= 0) ? __ldg(P.d_out + ((int64_t)b * P.N + r1) * n_out + o) : 0.f;
            go[nt][o] = pk2(g0, g1);
            if (g == 0 && ch == 0) db2[o] += g0 + g1;
          }
        }
        {
          const int c0 = ch * TP_CHUNK + TP_TPC * g;
          const int xcol = b * P.C + c0;
          const float* y_chunk = P.y + (size_t)(active ? b : 0) * P.M * NH * P.C + c0;
          float acc[TP_MT][4][4];
#pragma unroll
          for (int mt = 0; mt < TP_MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
          // (a) hidden pre-activation (without the bias), as in the forward
          for (int kb = 0; kb < nkb; ++kb) {
            if (nkb > 1) {
              __syncthreads();
              if (warp == 0) {
                const TpRow<NH> R = tp_row<NH>(P, V, tb + v, lane, s);
                float inv[NH], m[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) inv[h] = T->inv_l[h][lane], m[h] = T->m[h][lane];
                tp_store_block<NH, true>(R, s, V.d2 + (size_t)(T->off + kb * TP_KB) * TP_ROWS + lane, min(cnt - kb * TP_KB, TP_KB), lane, T, inv, m);
              }
              __syncthreads();
            }
            if (active) tp_mma_block<NH>(acc, T->p4 + p4_lane, cand, cnt, kb, y_chunk, P.C, t);
          }
          // (b) g1 = gelu'(b1 + pre) * (W2^T dOut[b, row, :]); parameter-gradient partials (two rows per packed op).
          //     ga[mt][nt] receives g1 directly in the A-fragment order of the products below: accumulator entries
          //     (e0, e2, e1, e3) = (column g | row 2t, column g+8 | row 2t, column g | row 2t+1, column g+8 | row 2t+1)
          float ga[TP_MT][4][4];
          if (active) {
#pragma unroll
            for (int mt = 0; mt < TP_MT; ++mt) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const int i = 2 * mt + half;
                const int c = c0 + i;
                const f32x2 bias = dup2(S.par[c]);
                f32x2 wv[NO], dw[NO];
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                  wv[o] = dup2(o < n_out ? S.par[(1 + o) * P.C + c] : 0.f);
                  dw[o] = 0ull;
                }
                f32x2 x[4], hid[4], dhid[4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) x[nt] = add2(pk2(acc[mt][nt][half * 2], acc[mt][nt][half * 2 + 1]), bias);
                tpg_gelu_pair2<4>(x, hid, dhid);
                float gsum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                  f32x2 up = 0ull;
#pragma unroll
                  for (int o = 0; o < NO; ++o) {
                    if (o < n_out) {
                      up = fma2(go[nt][o], wv[o], up);
                      dw[o] = fma2(go[nt][o], hid[nt], dw[o]);
                    }
                  }
                  float u0, u1, d0, d1;
                  unpk2(up, u0, u1);
                  unpk2(dhid[nt], d0, d1);
                  ga[mt][nt][half] = u0 * d0;      // row 2t   (k-slot t)
                  ga[mt][nt][2 + half] = u1 * d1;  // row 2t+1 (k-slot t+4)
                  gsum += ga[mt][nt][half] + ga[mt][nt][2 + half];
                }
#pragma unroll
                for (int a = 0; a < TP_MAX_CPS; ++a) {
                  if (a == ch) {
                    gacc[a][i][0] += gsum;
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                      float s0, s1;
                      unpk2(dw[o], s0, s1);
                      gacc[a][i][1 + o] += s0 + s1;
                    }
                  }
                }
              }
            }
          }
          // (c) dY^T += g1^T . P^  and  dZ^T = g1^T . (P^ (d2 - m)) per group of 8 candidates and head (k = the 32 tile rows)
          for (int kb = 0; kb < nkb; ++kb) {
            if (nkb > 1) {
              __syncthreads();
              if (warp == 0) {
                const TpRow<NH> R = tp_row<NH>(P, V, tb + v, lane, s);
                float inv[NH], m[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) inv[h] = T->inv_l[h][lane], m[h] = T->m[h][lane];
                tp_store_block<NH, true>(R, s, V.d2 + (size_t)(T->off + kb * TP_KB) * TP_ROWS + lane, min(cnt - kb * TP_KB, TP_KB), lane, T, inv, m);
              }
              __syncthreads();
            }
            if (!active) continue;
            const int base = kb * TP_KB;
            const int groups8 = (min(cnt - base, TP_KB) + 7) >> 3;
            const float4* pz_lane = reinterpret_cast<const float4*>(T->pz + g * TP_ZP + t * 8);
            for (int ct = 0; ct < groups8; ++ct) {
              // this thread's two candidates of the group: 8ct + 2t + e
              bool live[2];
              int sidx[2], jj[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int ci = base + ct * 8 + 2 * t + e;
                live[e] = ci < cnt;
                jj[e] = live[e] ? (int)cand[ci] : 0;
                sidx[e] = live[e] ? (int)bind[ci] : -1;
              }
              // Y rows of the two candidates for the scale gradient: requested before the products so that the trip to L2
              // overlaps the 48 MMAs of each head (dead candidates read row 0 and are skipped below)
              float4 yv[2][NH];
#pragma unroll
              for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int h = 0; h < NH; ++h)
                  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                               : "=f"(yv[e][h].x), "=f"(yv[e][h].y), "=f"(yv[e][h].z), "=f"(yv[e][h].w)
                               : "l"(y_chunk + ((size_t)jj[e] * NH + h) * P.C));
#pragma unroll
              for (int h = 0; h < NH; ++h) {
                float dy[TP_MT][4], dz[TP_MT][4];
#pragma unroll
                for (int mt = 0; mt < TP_MT; ++mt)
#pragma unroll
                  for (int e = 0; e < 4; ++e) dy[mt][e] = dz[mt][e] = 0.f;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                  // B fragments: k-slot t <-> tile row 8nt+2t, k-slot t+4 <-> row 8nt+2t+1; n = candidate 8ct+g
                  const float4* zp = pz_lane + ((h * TP_KB + ct * 8) * (TP_ZP / 4) + nt * 8);
                  const float4 pq = zp[0], zq = zp[1];
                  const uint32_t ph[2] = {__float_as_uint(pq.x), __float_as_uint(pq.y)};
                  const uint32_t pl[2] = {__float_as_uint(pq.z), __float_as_uint(pq.w)};
                  const uint32_t zh[2] = {__float_as_uint(zq.x), __float_as_uint(zq.y)};
                  const uint32_t zl[2] = {__float_as_uint(zq.z), __float_as_uint(zq.w)};
                  // A fragments: g1^T in fragment order (m = column, k = row), hi = the raw value, lo = its truncation residual
                  uint32_t al[TP_MT][4], ah[TP_MT][4];
#pragma unroll
                  for (int mt = 0; mt < TP_MT; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      ah[mt][i] = __float_as_uint(ga[mt][nt][i]);
                      al[mt][i] = tm_trunc_lo(ga[mt][nt][i]);
                    }
                  // split terms outermost: consecutive MMAs on one accumulator are four instructions apart
#pragma unroll
                  for (int term = 0; term < 3; ++term) {
#pragma unroll
                    for (int mt = 0; mt < TP_MT; ++mt) {
                      mma_tf32_16x8x8(dy[mt], term == 0 ? al[mt] : ah[mt], term == 1 ? pl : ph);
                      mma_tf32_16x8x8(dz[mt], term == 0 ? al[mt] : ah[mt], term == 1 ? zl : zh);
                    }
                  }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  if (!live[e]) continue;
                  // scale gradient: -sum_col Y_h[j, col] dZ[col, j]
                  ds_head[h] += yv[e][h].x * dz[0][e] + yv[e][h].y * dz[0][2 + e] + yv[e][h].z * dz[1][e] + yv[e][h].w * dz[1][2 + e];
                  // value gradient: the quad owns these cells of the slot (or adds to d_y itself if the column is unbound)
                  const float4 add = make_float4(dy[0][e], dy[0][2 + e], dy[1][e], dy[1][2 + e]);
                  if (sidx[e] >= 0) {
                    float4* c4 = reinterpret_cast<float4*>(S.slot_acc + ((size_t)sidx[e] * NH + h) * W + tm_slot_pos(xcol, sidx[e]));
                    float4 cur = *c4;
                    cur.x += add.x, cur.y += add.y, cur.z += add.z, cur.w += add.w;
                    *c4 = cur;
                  } else {
                    atomicAdd(reinterpret_cast<float4*>(P.d_y + (((size_t)b * P.M + jj[e]) * NH + h) * P.C + c0), add);
                  }
                }
              }
              __syncwarp();  // another lane of the quad may reach the same slot cell in the next group / tile
            }
          }
        }
      }
    }
  }
  // flush what is still bound (each quad its own cells)
  __syncwarp();
  for (int sidx = t; sidx < P.n_slots; sidx += 4) {
    const int j = S.slot_j[sidx];
    if (j < 0) continue;
    for (int u = warp; u < units; u += nwarps) {
      const int b = u / cps, ch = u - b * cps;
      tp_flush_cells<NH>(P, S, W, sidx, j, b, ch * TP_CHUNK + TP_TPC * g, b * P.C + ch * TP_CHUNK + TP_TPC * g);
    }
  }

  // ---- parameter gradients: one reduction per CTA ----
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float v = warp_sum(ds_head[h]);
    __syncthreads();
    if (lane == 0) S.red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
      for (int w = 0; w < nwarps; ++w) a += S.red[w];
      atomicAdd(P.d_scale + h, -a);
    }
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    if (o >= n_out) continue;  // uniform
    const float v = warp_sum(db2[o]);
    __syncthreads();
    if (lane == 0) S.red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
      for (int w = 0; w < nwarps; ++w) a += S.red[w];
      atomicAdd(P.d_b2 + o, a);
    }
  }
  // b1 and W2: per-thread partials (channel ch*32 + 4g + i) -> CTA sums in shared memory -> one RED per address
  __syncthreads();
#pragma unroll
  for (int a = 0; a < TP_MAX_CPS; ++a) {
    if (a >= cps) continue;
#pragma unroll
    for (int i = 0; i < TP_TPC; ++i) {
      const int c = a * TP_CHUNK + TP_TPC * g + i;
      // the four lanes of a quad hold partials of the same channels: fold them first
      float v = gacc[a][i][0];
      v += __shfl_xor_sync(FULL, v, 1);
      v += __shfl_xor_sync(FULL, v, 2);
      if (t == 0) atomicAdd(&S.gpar[c], v);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        if (o >= n_out) continue;
        float u = gacc[a][i][1 + o];
        u += __shfl_xor_sync(FULL, u, 1);
        u += __shfl_xor_sync(FULL, u, 2);
        if (t == 0) atomicAdd(&S.gpar[(1 + o) * P.C + c], u);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n_par; i += blockDim.x) atomicAdd(i < P.C ? P.d_b1 + i : P.d_w2 + (i - P.C), S.gpar[i]);
}

}  // namespace pit
