// K3m: fused decoder tail on the warp-level tensor-core path (mma.sync m16n8k8, 3xTF32) for shared meshes.
//
// Same contract as decoder_tail.cuh (pit.decoder, pit.py:124-127: cross position-attention `up` + `de` MLP), but
// the contraction of a row's 6..40 kept weights with the B*C-wide latent rows is no longer a per-row gather on the
// CUDA cores.  Rows are taken in tiles of 16 consecutive points.  On a spatially coherent mesh the union of the
// columns kept by ANY row of a tile is barely larger than one row's kept set (7.1 -> 8.9 columns at Darcy-421), so
//
//     pre^T [(b,c) x 16 rows] = Y^T [(b,c) x candidates] . P^T [candidates x 16 rows]        (per head, summed)
//
// is a small dense product at ~50 % density -- worth moving to the tensor pipe, which has several times the FMA rate
// of the fp32 pipe even after the 3xTF32 split.  The mask itself is still decided exactly as in the SIMT kernels
// (bit-exact d2, per-head cut, fp32 weights); only the products are split:  hi*hi + lo*hi + hi*lo with fp32
// accumulation (error ~2^-21, inside the 1e-5 parity budget).
//
// Work split.  A CTA (4 or 8 warps) walks its tiles in rounds (forward: one tile per warp, backward: 4):
//   phase 1  warp w "prepares" tile w of the round: scans the M columns held in registers against the 16 rows
//            (head-independent d2-space pre-filter, behind a bounding-box test per 32 columns), ballot-compacts the
//            candidate list, evaluates the normalised weights P^[h][row][cand] (backward: also P^ (d2 - m)) into a
//            shared-memory block, and the row sums;
//   phase 2  for every prepared tile, warp w owns the 64-column chunk(s) w, w+nwarps, ... of the B*C-wide hidden
//            vector: operand A = Y^T straight from global/L1 (a thread reads 8 contiguous floats of two candidate
//            rows; the column order inside an m16 tile is permuted so that a fragment is two 64-bit loads),
//            operand B = P^T from the shared block (conflict-free with the 20-float pitch), accumulators
//            [64 cols x 16 rows] in 32 registers; epilogue = bias + exact GELU + C->O projection on the fragments.
// The transposed orientation (columns on the MMA m axis, rows on n) is what makes the backward cheap: the
// accumulator fragment of g1^T is, register for register, the A fragment of dY^T = g1^T . P (reduction over rows),
// and the scale gradient follows from the same fragments with P (d2 - m) as the second B operand:
//     ds_h = -sum_{j,col} Y_h[j,col] (sum_i P^_ij (d2_ij - m_i) g1[i,col]).
// dY is accumulated without atomics in shared-memory slots bound to latent columns (one owner thread per cell), as
// in tall_bwd_kernel, and flushed with vector REDs.
//
// Tiles whose candidate list exceeds one 16-column block (incoherent meshes, unmasked decoders) are handled by
// recomputing further blocks on the fly -- slower, same results.
#pragma once
#include "decoder_tail.cuh"

namespace pit {

constexpr int TM_ROWS = 16;   // rows per tile (two n8 MMA tiles)
constexpr int TM_KT = 16;     // candidates per weight block (99.8 % of the Darcy-421 tiles have <= 16)
constexpr int TM_LD = 20;     // pitch of a block row in floats: fragment reads hit 32 distinct banks
constexpr int TM_MT = 4;      // m16 MMA tiles per warp pass
constexpr int TM_CHUNK = 16 * TM_MT;  // hidden-vector columns per warp pass
constexpr int TM_TPC = 2 * TM_MT;     // contiguous columns a thread owns inside a chunk
constexpr int TM_MAX_WARPS = 8;
constexpr int TM_BWD_ROUND = 4;  // tiles prepared per round in the backward (warps 0..3): keeps shared memory for dY slots

// (not volatile: the scheduler may interleave independent accumulator chains)
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// 3xTF32 split.  The tensor pipe reads the upper 19 bits of an fp32 register, so the raw value serves as the
// (truncated) high part and only the residual needs arithmetic; weights use a rounded high part (half the residual).
__device__ __forceinline__ uint32_t tm_trunc_lo(float x) { return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u)); }
__device__ __forceinline__ float tm_round_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// Exact-GELU pair as gelu_pair (A&S 7.1.26, |err| <= 1.5e-7) with flush-to-zero approximations, which drop the
// denormal-range fix-up code around MUFU.EX2 / MUFU.RCP: 14 instructions instead of ~21.
__device__ __forceinline__ void tm_gelu_pair(float x, float& g, float& dg) {
  const float ax = fabsf(x);
  const float u = ax * 0.84932180028801904f;  // sqrt(log2(e) / 2): exp(-x^2/2) = 2^(-u^2)
  float e, t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u * u));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.3275911f * 0.70710678118654752f, 1.f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, e, 1.f);
  const float phi = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  g = x * phi;
  dg = fmaf(x * 0.3989422804014327f, e, phi);
}
__device__ __forceinline__ float tm_gelu(float x) {
  float g, dg;
  tm_gelu_pair(x, g, dg);
  return g;
}

// Per-row constants of a tile row, held by lanes i and i+16 of the preparing warp.
template <int GEO, int NH>
struct TmRow {
  Point<GEO> o;
  float top[NH], cut[NH];
  float vcap;
  bool valid;
};

template <int GEO, int NH>
__device__ __forceinline__ TmRow<GEO, NH> tm_row(const TailParams& P, int row, const float (&s)[NH]) {
  TmRow<GEO, NH> R;
  R.valid = row < P.N;
  const int r = R.valid ? row : P.N - 1;
  R.o = load_point<GEO>(P.mesh_out, r, P.sd);
  const float vmin = __ldg(P.v_min + r);
  const float vlo = P.masked ? __ldg(P.v_lo + r) : 0.f, vhi = P.masked ? __ldg(P.v_hi + r) : 0.f;
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    R.top[h] = __fmul_rn(vmin, s[h]);
    R.cut[h] = !R.valid ? -INFINITY : (P.masked ? head_threshold(vlo, vhi, s[h], P.weight) : INFINITY);
  }
  R.vcap = !R.valid ? -1.f : (P.masked ? vhi * 1.000001f : INFINITY);  // see tall_scan_row: d2 above it can never be kept
  return R;
}

__device__ __forceinline__ float tm_warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ float tm_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// Candidate list of a tile: every column within the pre-filter radius of at least one of its rows, in column order.
// Euclidean geometries first test each column against the tile's bounding box inflated by the largest radius; a
// group of 32 columns (one per lane) in which no lane passes skips the 16 exact row tests.
template <int GEO, int CPL, int NH>
__device__ __forceinline__ int tm_candidates(const TmRow<GEO, NH>& R, const Point<GEO> (&col)[CPL], int M, int lane, float period,
                                             int16_t* cand) {
  uint32_t groups = 0xffffffffu;
  if (GEO == GEO_EUCLID1 || GEO == GEO_EUCLID2) {
    const float x0 = tm_warp_min(R.o.x), x1 = tm_warp_max(R.o.x);
    const float y0 = GEO == GEO_EUCLID2 ? tm_warp_min(R.o.y) : 0.f, y1 = GEO == GEO_EUCLID2 ? tm_warp_max(R.o.y) : 0.f;
    const float reach = tm_warp_max(R.vcap) * 1.0001f;  // slack for the different rounding of the box distance
    groups = 0;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const float dx = fmaxf(fmaxf(x0 - col[c].x, col[c].x - x1), 0.f);
      const float dy = GEO == GEO_EUCLID2 ? fmaxf(fmaxf(y0 - col[c].y, col[c].y - y1), 0.f) : 0.f;
      if (__any_sync(FULL, fmaf(dx, dx, dy * dy) <= reach)) groups |= 1u << c;
    }
  }
  uint32_t flags = 0;
  for (int i = 0; i < TM_ROWS; ++i) {
    Point<GEO> o;
    o.x = __shfl_sync(FULL, R.o.x, i);
    o.y = __shfl_sync(FULL, R.o.y, i);
    const float vc = __shfl_sync(FULL, R.vcap, i);
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      if ((groups >> c) & 1u) flags |= (dist2<GEO>(o, col[c], period) <= vc) ? (1u << c) : 0u;  // warp-uniform branch
  }
  int n = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    if (!((groups >> c) & 1u)) continue;
    const int j = c * 32 + lane;
    const bool f = ((flags >> c) & 1u) && j < M;
    const unsigned m = __ballot_sync(FULL, f);
    if (f) cand[n + __popc(m & lt)] = (int16_t)j;
    n += __popc(m);
  }
  __syncwarp();
  return n;
}

// One block (<= 16 candidates, zero-filled to the next multiple of 8) of weights scaled by post[h]:
// lane -> (row lane&15, candidates of parity lane>>4).  pt = [NH][16][TM_LD]; zt (backward) = [NH][16][TM_LD], of which
// plane 0 receives d2 here and tm_block_finish turns all planes into P^ (d2 - m).
// psum / pdsum accumulate post*p and post*p*d2 of this lane's entries.
template <int GEO, int NH>
__device__ __forceinline__ void tm_block(const TailParams& P, const TmRow<GEO, NH>& R, const float (&s)[NH], float period,
                                         const int16_t* cand, int cnt, int kb, int lane, float* pt, float* zt, const float (&post)[NH],
                                         float (&psum)[NH], float (&pdsum)[NH]) {
  const int i = lane & 15;
  const int base = kb * TM_KT;
  const int kmax = min(TM_KT, (cnt - base + 7) & ~7);
  for (int k = lane >> 4; k < kmax; k += 2) {
    const bool live = base + k < cnt;
    const int j = live ? (int)cand[base + k] : 0;
    const Point<GEO> q = load_point<GEO>(P.mesh_in, j, P.sd);
    const float d2 = dist2<GEO>(R.o, q, period);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float p = 0.f;
      if (live) {
        const float sc = __fmul_rn(d2, s[h]);
        if (sc <= R.cut[h]) p = __expf(__fsub_rn(R.top[h], sc)) * post[h];
      }
      psum[h] += p;
      pdsum[h] = fmaf(p, d2, pdsum[h]);
      pt[(h * TM_ROWS + i) * TM_LD + k] = p;
    }
    if (zt) zt[i * TM_LD + k] = live ? d2 : 0.f;
  }
}
// Second pass over the entries this lane wrote: forward scales by 1/l (zt == null), backward forms P^ (d2 - m_h).
template <int NH>
__device__ __forceinline__ void tm_block_finish(int cnt, int kb, int lane, float* pt, float* zt, const float (&inv_l)[NH],
                                                const float (&m)[NH]) {
  const int i = lane & 15;
  const int kmax = min(TM_KT, (cnt - kb * TM_KT + 7) & ~7);
  for (int k = lane >> 4; k < kmax; k += 2) {
    if (zt) {
      const float d2 = zt[i * TM_LD + k];
#pragma unroll
      for (int h = NH - 1; h >= 0; --h) zt[(h * TM_ROWS + i) * TM_LD + k] = pt[(h * TM_ROWS + i) * TM_LD + k] * (d2 - m[h]);
    } else {
#pragma unroll
      for (int h = 0; h < NH; ++h) pt[(h * TM_ROWS + i) * TM_LD + k] *= inv_l[h];
    }
  }
}

// Shared-memory image of a prepared tile.
template <int NH, bool BWD>
struct TmTile {
  float p[NH][TM_ROWS][TM_LD];             // normalised weights P^ of block 0 (or of the block being processed)
  float z[BWD ? NH : 1][BWD ? TM_ROWS : 1][TM_LD];  // backward: P^ (d2 - m)
  float inv_l[NH][TM_ROWS];                // 1 / row sum (0 for rows past the end)
  float m[NH][TM_ROWS];                    // backward: sum_j P^ d2
  int cnt;
  int pad[3];
};

// Rebuild block kb of a prepared tile in place (candidate lists longer than one block); one warp.
template <int GEO, int NH, bool BWD>
__device__ __forceinline__ void tm_rebuild(const TailParams& P, TmTile<NH, BWD>* T, const int16_t* cand, int row0, int kb, int lane,
                                           const float (&s)[NH], float period) {
  const TmRow<GEO, NH> R = tm_row<GEO, NH>(P, row0 + (lane & 15), s);
  float inv[NH], m[NH], ps[NH], pd[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    inv[h] = T->inv_l[h][lane & 15];
    m[h] = T->m[h][lane & 15];
    ps[h] = pd[h] = 0.f;
  }
  tm_block<GEO, NH>(P, R, s, period, cand, T->cnt, kb, lane, &T->p[0][0][0], BWD ? &T->z[0][0][0] : nullptr, inv, ps, pd);
  if (BWD) tm_block_finish<NH>(T->cnt, kb, lane, &T->p[0][0][0], &T->z[0][0][0], inv, m);
}

// pre^T += Y_h^T . P^_h^T for one block of one tile and one 64-column chunk.
//   acc[mt][nt][e]: chunk column TPC*g + 2*mt + (e >> 1), tile row 8*nt + 2*t + (e & 1).
// y_chunk points at Y[b, 0, 0, c] for this thread's TPC columns; row j of head h sits (NH*j + h)*C floats further.
template <int NH>
__device__ __forceinline__ void tm_mma_block(float (&acc)[TM_MT][2][4], const float (*p)[TM_ROWS][TM_LD], const int16_t* cand, int cnt,
                                             int kb, const float* y_chunk, int C, int g, int t) {
  const int base = kb * TM_KT;
  const int ksteps = (min(cnt - base, TM_KT) + 7) >> 3;
  for (int ks = 0; ks < ksteps; ++ks) {
    const int ka = base + ks * 8 + t, kc = ka + 4;
    const int ja = ka < cnt ? (int)cand[ka] : 0, jb = kc < cnt ? (int)cand[kc] : 0;
    const float* ra0 = y_chunk + (size_t)ja * NH * C;
    const float* rb0 = y_chunk + (size_t)jb * NH * C;
    // the loads of every head are issued before the first MMA of the step
    float2 ya[NH][TM_MT], yb[NH][TM_MT];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float2* ra = reinterpret_cast<const float2*>(ra0 + h * C);
      const float2* rb = reinterpret_cast<const float2*>(rb0 + h * C);
#pragma unroll
      for (int mt = 0; mt < TM_MT; ++mt) {
        ya[h][mt] = __ldg(ra + mt);
        yb[h][mt] = __ldg(rb + mt);
      }
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float p0 = p[h][g + 8 * nt][ks * 8 + t], p1 = p[h][g + 8 * nt][ks * 8 + t + 4];
        const float h0 = tm_round_hi(p0), h1 = tm_round_hi(p1);
        bh[nt][0] = __float_as_uint(h0), bh[nt][1] = __float_as_uint(h1);
        bl[nt][0] = __float_as_uint(p0 - h0), bl[nt][1] = __float_as_uint(p1 - h1);
      }
      // split terms outermost: the dependent MMAs on one accumulator are eight instructions apart
#pragma unroll
      for (int mt = 0; mt < TM_MT; ++mt) {
        const uint32_t al[4] = {tm_trunc_lo(ya[h][mt].x), tm_trunc_lo(ya[h][mt].y), tm_trunc_lo(yb[h][mt].x), tm_trunc_lo(yb[h][mt].y)};
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_tf32_16x8x8(acc[mt][nt], al, bh[nt]);
      }
#pragma unroll
      for (int term = 0; term < 2; ++term) {
#pragma unroll
        for (int mt = 0; mt < TM_MT; ++mt) {
          const uint32_t ah[4] = {__float_as_uint(ya[h][mt].x), __float_as_uint(ya[h][mt].y), __float_as_uint(yb[h][mt].x), __float_as_uint(yb[h][mt].y)};
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) mma_tf32_16x8x8(acc[mt][nt], ah, term == 0 ? bl[nt] : bh[nt]);
        }
      }
    }
  }
}

__host__ __device__ inline size_t tm_align(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t tm_cand_bytes(int M) { return tm_align((size_t)M * 2); }
__host__ __device__ inline size_t tm_tile_bytes(int nh, bool bwd) {
  return nh == 1 ? (bwd ? tm_align(sizeof(TmTile<1, true>)) : tm_align(sizeof(TmTile<1, false>)))
                 : (bwd ? tm_align(sizeof(TmTile<2, true>)) : tm_align(sizeof(TmTile<2, false>)));
}
__host__ __device__ inline size_t tm_fwd_smem_bytes(int nh, int M, int C, int O, int threads) {
  return 2 * (threads / 32) * (tm_tile_bytes(nh, false) + tm_cand_bytes(M)) + tm_align((size_t)C * (1 + O) * 4);
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
// NO: compile-time bound on out_dim -- 1 (the scalar-output models) or TAIL_MAX_OUT (any out_dim, decided at run time).
template <int GEO, int CPL, int NH, int NO>
__global__ void __launch_bounds__(32 * TM_MAX_WARPS, 3) tail_mma_fwd_kernel(const TailParams P) {
  const int n_out = NO == 1 ? 1 : P.O;
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  using Tile = TmTile<NH, false>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const size_t tile_stride = tm_align(sizeof(Tile));
  unsigned char* cand_base = tall_smem_raw + 2 * nwarps * tile_stride;
  const size_t cand_stride = tm_cand_bytes(P.M);
  float* par = reinterpret_cast<float*>(cand_base + 2 * nwarps * cand_stride);  // [b1 (C) | W2 (O x C)]

  const int g = lane >> 2, t = lane & 3;
  const float period = P.period ? __ldg(P.period) : 0.f;
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  for (int i = tid; i < P.C * (1 + n_out); i += blockDim.x) par[i] = i < P.C ? __ldg(P.b1 + i) : __ldg(P.w2 + (i - P.C));
  float b2r[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) b2r[o] = o < n_out ? __ldg(P.b2 + o) : 0.f;

  const int chunks = P.B * P.C / TM_CHUNK;
  const int n_tiles = (P.N + TM_ROWS - 1) / TM_ROWS;
  const int tile_begin = blockIdx.x * P.rows_per_unit;  // rows_per_unit counts tiles here
  const int tile_end = min(n_tiles, tile_begin + P.rows_per_unit);
  const int c0 = (TM_TPC * g) % P.C;  // this thread's hidden channels (the same in every chunk: C divides the chunk width)
  const int group = P.C / TM_TPC;     // consecutive g sharing a sample
  const int log2c = 31 - __clz(P.C);
  int round = 0;
  for (int tb = tile_begin; tb < tile_end; tb += nwarps, ++round) {
    const int in_round = min(nwarps, tile_end - tb);
    const int set = (round & 1) * nwarps;
    // ---- phase 1: one tile per warp ----
    if (warp < in_round) {
      Point<GEO> col[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int j = c * 32 + lane;
        col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
      }
      Tile* T = reinterpret_cast<Tile*>(tall_smem_raw + (set + warp) * tile_stride);
      int16_t* cand = reinterpret_cast<int16_t*>(cand_base + (set + warp) * cand_stride);
      const int row = (tb + warp) * TM_ROWS + (lane & 15);
      const TmRow<GEO, NH> R = tm_row<GEO, NH>(P, row, s);
      const int cnt = tm_candidates<GEO, CPL, NH>(R, col, P.M, lane, period, cand);
      float psum[NH], pdsum[NH], one[NH], inv[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        psum[h] = pdsum[h] = 0.f;
        one[h] = 1.f;
      }
      const int nkb = (cnt + TM_KT - 1) / TM_KT;
      for (int kb = nkb - 1; kb >= 0; --kb)  // block 0 last: it is the one left in the tile for phase 2
        tm_block<GEO, NH>(P, R, s, period, cand, cnt, kb, lane, &T->p[0][0][0], nullptr, one, psum, pdsum);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float l = psum[h] + __shfl_xor_sync(FULL, psum[h], 16);
        inv[h] = R.valid ? 1.f / l : 0.f;
        if (lane < TM_ROWS) {
          T->inv_l[h][lane] = inv[h];
          T->m[h][lane] = 0.f;
          if (R.valid) P.rowsum[(int64_t)h * P.N + row] = l;
        }
      }
      if (nkb > 0) tm_block_finish<NH>(cnt, 0, lane, &T->p[0][0][0], nullptr, inv, inv);
      if (lane == 0) T->cnt = cnt;
    }
    __syncthreads();
    // ---- phase 2: every warp contracts its chunk(s) of every prepared tile ----
    for (int v = 0; v < in_round; ++v) {
      Tile* T = reinterpret_cast<Tile*>(tall_smem_raw + (set + v) * tile_stride);
      const int16_t* cand = reinterpret_cast<const int16_t*>(cand_base + (set + v) * cand_stride);
      const int cnt = T->cnt;
      const int nkb = (cnt + TM_KT - 1) / TM_KT;
      const int row0 = (tb + v) * TM_ROWS;
      for (int ch0 = 0; ch0 < chunks; ch0 += nwarps) {
        const int chunk = ch0 + warp;
        const bool active = chunk < chunks;
        const int b = active ? (chunk * TM_CHUNK + TM_TPC * g) >> log2c : 0;
        const float* y_chunk = P.y + (size_t)b * P.M * NH * P.C + c0;
        float acc[TM_MT][2][4];
#pragma unroll
        for (int mt = 0; mt < TM_MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = par[c0 + 2 * mt + (e >> 1)];  // the accumulators start from the bias b1
        for (int kb = 0; kb < nkb; ++kb) {
          if (nkb > 1) {  // rare: the candidate list spans several blocks -> rebuild block kb in place (CTA-uniform branch)
            __syncthreads();
            if (warp == 0) tm_rebuild<GEO, NH, false>(P, T, cand, row0, kb, lane, s, period);
            __syncthreads();
          }
          if (active) tm_mma_block<NH>(acc, T->p, cand, cnt, kb, y_chunk, P.C, g, t);
        }
        if (!active) continue;
        // epilogue: out[b, row, o] = b2[o] + sum_c W2[o, c] gelu(b1[c] + pre[c]); the thread holds channels c0..c0+TPC-1 of
        // rows 2t, 2t+1, 2t+8, 2t+9; the lanes of `group` consecutive g share the sample.
        float part[NO][4];
#pragma unroll
        for (int o = 0; o < NO; ++o)
#pragma unroll
          for (int q = 0; q < 4; ++q) part[o][q] = 0.f;
#pragma unroll
        for (int mt = 0; mt < TM_MT; ++mt) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c = c0 + 2 * mt + half;
            float hid[4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) hid[nt * 2 + e] = tm_gelu(acc[mt][nt][half * 2 + e]);
#pragma unroll
            for (int o = 0; o < NO; ++o) {
              if (o < n_out) {
                const float wv = par[(1 + o) * P.C + c];
#pragma unroll
                for (int q = 0; q < 4; ++q) part[o][q] = fmaf(wv, hid[q], part[o][q]);
              }
            }
          }
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) {
          if (o >= n_out) break;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float vsum = part[o][q];
            for (int off = 1; off < group; off <<= 1) vsum += __shfl_xor_sync(FULL, vsum, 4 * off);
            const int row = row0 + (q >> 1) * 8 + 2 * t + (q & 1);
            if ((g % group) == 0 && row < P.N) P.out[((int64_t)b * P.N + row) * n_out + o] = vsum + b2r[o];
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
struct TmBwdSmem {
  unsigned char* tiles;  // [TM_BWD_ROUND] TmTile<NH, true>
  unsigned char* cand;   // [TM_BWD_ROUND][M] int16
  float* slot_acc;       // [n_slots][NH][W], columns rotated inside 32-float windows by 4*(slot & 7)
  float* priv;           // [(1+O)*TPC][threads]: per-thread partial sums of d_b1 and d_w2 for its channels
  float* par;            // [C + O*C]: b1 then W2
  float* gpar;           // [C + O*C]: CTA-level reduction of d_b1, d_w2
  float* red;            // [TM_MAX_WARPS]
  int16_t* map;          // [M] column -> slot (-1: unbound)
  int16_t* slot_j;       // [n_slots] slot -> column (-1: free)
  uint32_t* touched;     // [M] columns that are candidates of a tile of the current round (set with atomicOr: several warps may mark one)
  uint8_t* evict;        // [n_slots]
  int* ctl;              // [0] columns needing a slot, [1] [2] free-slot bit masks, [3] bind counter
};

__host__ __device__ inline size_t tm_bwd_smem_bytes(int nh, int M, int W, int C, int O, int n_slots, int threads) {
  return TM_BWD_ROUND * (tm_tile_bytes(nh, true) + tm_cand_bytes(M)) + tm_align((size_t)n_slots * nh * W * 4) +
         tm_align((size_t)(1 + O) * TM_TPC * threads * 4) + 2 * tm_align((size_t)C * (1 + O) * 4) + 64 + tm_align((size_t)M * 2) +
         tm_align((size_t)n_slots * 2) + tm_align((size_t)M * 4) + tm_align(n_slots) + 16;
}

__device__ inline TmBwdSmem tm_bwd_carve(unsigned char* p, int nh, int M, int W, int C, int O, int n_slots, int threads) {
  TmBwdSmem s{};
  s.tiles = p;
  p += TM_BWD_ROUND * tm_tile_bytes(nh, true);
  s.cand = p;
  p += TM_BWD_ROUND * tm_cand_bytes(M);
  s.slot_acc = reinterpret_cast<float*>(p);
  p += tm_align((size_t)n_slots * nh * W * 4);
  s.priv = reinterpret_cast<float*>(p);
  p += tm_align((size_t)(1 + O) * TM_TPC * threads * 4);
  s.par = reinterpret_cast<float*>(p);
  p += tm_align((size_t)C * (1 + O) * 4);
  s.gpar = reinterpret_cast<float*>(p);
  p += tm_align((size_t)C * (1 + O) * 4);
  s.red = reinterpret_cast<float*>(p);
  p += 64;
  s.map = reinterpret_cast<int16_t*>(p);
  p += tm_align((size_t)M * 2);
  s.slot_j = reinterpret_cast<int16_t*>(p);
  p += tm_align((size_t)n_slots * 2);
  s.touched = reinterpret_cast<uint32_t*>(p);
  p += tm_align((size_t)M * 4);
  s.evict = reinterpret_cast<uint8_t*>(p);
  p += tm_align(n_slots);
  s.ctl = reinterpret_cast<int*>(p);
  return s;
}

// Position of column x inside the slot row of slot `sidx`: 16-byte groups rotate inside their 128-byte window so that
// the four lanes of a quad, which address four different slots at the same column, fall into different banks.
__device__ __forceinline__ int tm_slot_pos(int x, int sidx) { return (x & ~31) | ((x + 4 * (sidx & 7)) & 31); }

// Flush one slot into d_y with vector REDs and clear it; the cells are split over the CTA's threads.
template <int NH>
__device__ __forceinline__ void tm_flush_slot(const TailParams& P, const TmBwdSmem& S, int W, int log2c, int sidx) {
  const int j = S.slot_j[sidx];
  for (int h = 0; h < NH; ++h) {
    for (int x = 4 * threadIdx.x; x < W; x += 4 * blockDim.x) {
      float4* cell = reinterpret_cast<float4*>(S.slot_acc + ((size_t)sidx * NH + h) * W + tm_slot_pos(x, sidx));
      const int b = x >> log2c, c = x & (P.C - 1);
      atomicAdd(reinterpret_cast<float4*>(P.d_y + (((size_t)b * P.M + j) * NH + h) * P.C + c), *cell);
      *cell = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Bit masks of the free slots (slot_j < 0), computed by warp 0 into ctl[1], ctl[2] (n_slots <= 64).
__device__ __forceinline__ void tm_free_masks(const TailParams& P, const TmBwdSmem& S, int lane) {
  const unsigned m0 = __ballot_sync(FULL, lane < P.n_slots && S.slot_j[min(lane, P.n_slots - 1)] < 0);
  const unsigned m1 = __ballot_sync(FULL, lane + 32 < P.n_slots && S.slot_j[min(lane + 32, P.n_slots - 1)] < 0);
  if (lane == 0) {
    S.ctl[1] = (int)m0;
    S.ctl[2] = (int)m1;
  }
}

// Slot management of one round, called by the whole CTA after phase 1 has marked the candidate columns in
// S.touched: every touched column without a slot gets a free one; if there are not enough, the slots of columns
// NOT touched in this round are flushed and recycled first (the candidate window slides along the mesh, so these
// are the columns left behind); whatever is still unbound afterwards goes to d_y with direct REDs.
template <int NH>
__device__ __forceinline__ void tm_bind_slots(const TailParams& P, const TmBwdSmem& S, int W, int log2c) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) tm_free_masks(P, S, lane);
  for (int j = tid; j < P.M; j += blockDim.x)
    if (S.touched[j] && S.map[j] < 0) atomicAdd(&S.ctl[0], 1);
  __syncthreads();
  if (S.ctl[0] > __popc((unsigned)S.ctl[1]) + __popc((unsigned)S.ctl[2])) {  // CTA-uniform
    for (int sidx = tid; sidx < P.n_slots; sidx += blockDim.x) S.evict[sidx] = S.slot_j[sidx] >= 0 && !S.touched[S.slot_j[sidx]];
    __syncthreads();
    for (int sidx = 0; sidx < P.n_slots; ++sidx)
      if (S.evict[sidx]) tm_flush_slot<NH>(P, S, W, log2c, sidx);
    __syncthreads();
    for (int sidx = tid; sidx < P.n_slots; sidx += blockDim.x) {
      if (S.evict[sidx]) {
        S.map[S.slot_j[sidx]] = -1;
        S.slot_j[sidx] = -1;
      }
    }
    __syncthreads();
    if (warp == 0) tm_free_masks(P, S, lane);
    __syncthreads();
  }
  const unsigned m0 = (unsigned)S.ctl[1], m1 = (unsigned)S.ctl[2];
  const int n0 = __popc(m0), n1 = __popc(m1);
  for (int j = tid; j < P.M; j += blockDim.x) {
    if (S.touched[j] && S.map[j] < 0) {
      const int k = atomicAdd(&S.ctl[3], 1);
      int sidx = -1;
      if (k < n0)
        sidx = (int)__fns(m0, 0, k + 1);
      else if (k < n0 + n1)
        sidx = 32 + (int)__fns(m1, 0, k - n0 + 1);
      if (sidx >= 0) {
        S.map[j] = (int16_t)sidx;
        S.slot_j[sidx] = (int16_t)j;
      }
    }
  }
  __syncthreads();
  for (int j = tid; j < P.M; j += blockDim.x) S.touched[j] = 0;
  if (tid == 0) {
    S.ctl[0] = 0;
    S.ctl[3] = 0;
  }
  // the caller's next __syncthreads (end of the round) orders these resets before the next phase 1
}

template <int GEO, int CPL, int NH, int NO>
__global__ void __launch_bounds__(32 * TM_MAX_WARPS, 2) tail_mma_bwd_kernel(const TailParams P) {
  const int n_out = NO == 1 ? 1 : P.O;
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  using Tile = TmTile<NH, true>;
  const int W = P.B * P.C;
  const TmBwdSmem S = tm_bwd_carve(tall_smem_raw, NH, P.M, W, P.C, n_out, P.n_slots, blockDim.x);
  const size_t tile_stride = tm_align(sizeof(Tile));
  const size_t cand_stride = tm_cand_bytes(P.M);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float period = P.period ? __ldg(P.period) : 0.f;
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  const int n_par = P.C * (1 + n_out);
  for (int i = tid; i < n_par; i += blockDim.x) {
    S.par[i] = i < P.C ? __ldg(P.b1 + i) : __ldg(P.w2 + (i - P.C));
    S.gpar[i] = 0.f;
  }
  for (int j = tid; j < P.M; j += blockDim.x) {
    S.map[j] = -1;
    S.touched[j] = 0;
  }
  for (int i = tid; i < P.n_slots * NH * W; i += blockDim.x) S.slot_acc[i] = 0.f;
  for (int i = 0; i < (1 + n_out) * TM_TPC; ++i) S.priv[i * blockDim.x + tid] = 0.f;
  for (int i = tid; i < P.n_slots; i += blockDim.x) {
    S.slot_j[i] = -1;
    S.evict[i] = 0;
  }
  if (tid < 4) S.ctl[tid] = 0;
  const int log2c = 31 - __clz(P.C);
  __syncthreads();

  float ds_head[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) ds_head[h] = 0.f;
  float db2[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) db2[o] = 0.f;

  const int chunks = W / TM_CHUNK;
  const int n_tiles = (P.N + TM_ROWS - 1) / TM_ROWS;
  const int tile_begin = blockIdx.x * P.rows_per_unit;
  const int tile_end = min(n_tiles, tile_begin + P.rows_per_unit);
  const int c0 = (TM_TPC * g) % P.C;
  const int group = P.C / TM_TPC;
  for (int tb = tile_begin; tb < tile_end; tb += TM_BWD_ROUND) {
    const int in_round = min(TM_BWD_ROUND, tile_end - tb);
    // ---- phase 1: one tile per warp (warps 0..3): candidates, P^, P^ (d2 - m), 1/l, m ----
    if (warp < in_round) {
      Point<GEO> col[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int j = c * 32 + lane;
        col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
      }
      Tile* T = reinterpret_cast<Tile*>(S.tiles + warp * tile_stride);
      int16_t* cand = reinterpret_cast<int16_t*>(S.cand + warp * cand_stride);
      const int row = (tb + warp) * TM_ROWS + (lane & 15);
      const TmRow<GEO, NH> R = tm_row<GEO, NH>(P, row, s);
      const int cnt = tm_candidates<GEO, CPL, NH>(R, col, P.M, lane, period, cand);
      float psum[NH], pdsum[NH], inv[NH], m[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        psum[h] = pdsum[h] = 0.f;
        inv[h] = R.valid ? 1.f / __ldg(P.rowsum + (int64_t)h * P.N + row) : 0.f;
      }
      const int nkb = (cnt + TM_KT - 1) / TM_KT;
      for (int kb = nkb - 1; kb >= 0; --kb)
        tm_block<GEO, NH>(P, R, s, period, cand, cnt, kb, lane, &T->p[0][0][0], &T->z[0][0][0], inv, psum, pdsum);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        m[h] = pdsum[h] + __shfl_xor_sync(FULL, pdsum[h], 16);
        if (lane < TM_ROWS) {
          T->inv_l[h][lane] = inv[h];
          T->m[h][lane] = m[h];
        }
      }
      if (nkb > 0) tm_block_finish<NH>(cnt, 0, lane, &T->p[0][0][0], &T->z[0][0][0], inv, m);
      for (int k = lane; k < cnt; k += 32) atomicOr(&S.touched[cand[k]], 1u);
      if (lane == 0) T->cnt = cnt;
    }
    __syncthreads();
    tm_bind_slots<NH>(P, S, W, log2c);
    __syncthreads();
    // ---- phase 2 ----
    for (int v = 0; v < in_round; ++v) {
      Tile* T = reinterpret_cast<Tile*>(S.tiles + v * tile_stride);
      const int16_t* cand = reinterpret_cast<const int16_t*>(S.cand + v * cand_stride);
      const int cnt = T->cnt;
      const int nkb = (cnt + TM_KT - 1) / TM_KT;
      const int row0 = (tb + v) * TM_ROWS;
      for (int ch0 = 0; ch0 < chunks; ch0 += nwarps) {
        const int chunk = ch0 + warp;
        const bool active = chunk < chunks;
        const int xcol = chunk * TM_CHUNK + TM_TPC * g;  // first of this thread's columns of the B*C-wide vector
        const int b = active ? xcol >> log2c : 0;
        const float* y_chunk = P.y + (size_t)b * P.M * NH * P.C + c0;
        float acc[TM_MT][2][4];
#pragma unroll
        for (int mt = 0; mt < TM_MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = S.par[c0 + 2 * mt + (e >> 1)];  // start from the bias b1
        // (a) hidden pre-activation, as in the forward
        for (int kb = 0; kb < nkb; ++kb) {
          if (nkb > 1) {
            __syncthreads();
            if (warp == 0) tm_rebuild<GEO, NH, true>(P, T, cand, row0, kb, lane, s, period);
            __syncthreads();
          }
          if (active) tm_mma_block<NH>(acc, T->p, cand, cnt, kb, y_chunk, P.C, g, t);
        }
        // (b) g1 = gelu'(pre) * (W2^T dOut[b, row, :]) in place; parameter-gradient partials
        if (active) {
          float go[4][NO];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int row = row0 + (q >> 1) * 8 + 2 * t + (q & 1);
#pragma unroll
            for (int o = 0; o < NO; ++o) {
              go[q][o] = (o < n_out && row < P.N) ? __ldg(P.d_out + ((int64_t)b * P.N + row) * n_out + o) : 0.f;
              if ((g % group) == 0) db2[o] += go[q][o];
            }
          }
#pragma unroll
          for (int mt = 0; mt < TM_MT; ++mt) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int i = 2 * mt + half;
              const int c = c0 + i;
              float wv[NO], dw[NO];
#pragma unroll
              for (int o = 0; o < NO; ++o) {
                wv[o] = o < n_out ? S.par[(1 + o) * P.C + c] : 0.f;
                dw[o] = 0.f;
              }
              float gsum = 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float hid, dhid;
                tm_gelu_pair(acc[mt][q >> 1][half * 2 + (q & 1)], hid, dhid);
                float up = 0.f;
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                  if (o < n_out) {
                    up = fmaf(go[q][o], wv[o], up);
                    dw[o] = fmaf(go[q][o], hid, dw[o]);
                  }
                }
                const float g1 = up * dhid;
                acc[mt][q >> 1][half * 2 + (q & 1)] = g1;
                gsum += g1;
              }
              S.priv[i * blockDim.x + tid] += gsum;
#pragma unroll
              for (int o = 0; o < NO; ++o)
                if (o < n_out) S.priv[((1 + o) * TM_TPC + i) * blockDim.x + tid] += dw[o];
            }
          }
        }
        // (c) dY^T += g1^T . P^  and  dZ^T = g1^T . (P^ (d2 - m)) per group of 8 candidates and head
        for (int kb = 0; kb < nkb; ++kb) {
          if (nkb > 1) {
            __syncthreads();
            if (warp == 0) tm_rebuild<GEO, NH, true>(P, T, cand, row0, kb, lane, s, period);
            __syncthreads();
          }
          if (!active) continue;
          const int base = kb * TM_KT;
          const int groups8 = (min(cnt - base, TM_KT) + 7) >> 3;
          for (int ct = 0; ct < groups8; ++ct) {
            // this thread's two candidates of the group: 8ct + 2t + e  (dy/dz[mt][2*(col&1) + e])
            bool live[2], bound[2];
            int yoff[2];             // element offset of Y[b, j, 0, c0] / d_y[b, j, 0, c0]
            int soff[2][TM_TPC / 4];  // element offsets of the thread's slot cells (head 0)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ci = base + ct * 8 + 2 * t + e;
              live[e] = ci < cnt;
              const int j = live[e] ? (int)cand[ci] : 0;
              const int sidx = live[e] ? (int)S.map[j] : -1;
              bound[e] = sidx >= 0;
              yoff[e] = (b * P.M + j) * NH * P.C + c0;
#pragma unroll
              for (int v4 = 0; v4 < TM_TPC / 4; ++v4) soff[e][v4] = bound[e] ? sidx * NH * W + tm_slot_pos(xcol + 4 * v4, sidx) : 0;
            }
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              float dy[TM_MT][4], dz[TM_MT][4];
#pragma unroll
              for (int mt = 0; mt < TM_MT; ++mt)
#pragma unroll
                for (int e = 0; e < 4; ++e) dy[mt][e] = dz[mt][e] = 0.f;
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                // B fragments: k-slot t <-> tile row 8nt+2t, k-slot t+4 <-> row 8nt+2t+1; n = candidate 8ct+g
                uint32_t ph[2], pl[2], zh[2], zl[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int r = 8 * nt + 2 * t + e;
                  const float pn = T->p[h][r][ct * 8 + g], pz = T->z[h][r][ct * 8 + g];
                  const float h0 = tm_round_hi(pn), h1 = tm_round_hi(pz);
                  ph[e] = __float_as_uint(h0), pl[e] = __float_as_uint(pn - h0);
                  zh[e] = __float_as_uint(h1), zl[e] = __float_as_uint(pz - h1);
                }
                // A fragments: the accumulator registers of g1^T, reordered (m = column, k = row); split terms outermost
#pragma unroll
                for (int mt = 0; mt < TM_MT; ++mt) {
                  const float a0 = acc[mt][nt][0], a1 = acc[mt][nt][2], a2 = acc[mt][nt][1], a3 = acc[mt][nt][3];
                  const uint32_t al[4] = {tm_trunc_lo(a0), tm_trunc_lo(a1), tm_trunc_lo(a2), tm_trunc_lo(a3)};
                  mma_tf32_16x8x8(dy[mt], al, ph);
                  mma_tf32_16x8x8(dz[mt], al, zh);
                }
#pragma unroll
                for (int term = 0; term < 2; ++term) {
#pragma unroll
                  for (int mt = 0; mt < TM_MT; ++mt) {
                    const uint32_t ah[4] = {__float_as_uint(acc[mt][nt][0]), __float_as_uint(acc[mt][nt][2]), __float_as_uint(acc[mt][nt][1]),
                                            __float_as_uint(acc[mt][nt][3])};
                    mma_tf32_16x8x8(dy[mt], ah, term == 0 ? pl : ph);
                    mma_tf32_16x8x8(dz[mt], ah, term == 0 ? zl : zh);
                  }
                }
              }
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                if (!live[e]) continue;
                // scale gradient: -sum Y_h[j, col] dZ[col, j]
                const float2* yr = reinterpret_cast<const float2*>(P.y + yoff[e] + h * P.C);
                float dot = 0.f;
#pragma unroll
                for (int mt = 0; mt < TM_MT; ++mt) {
                  const float2 yv = __ldg(yr + mt);
                  dot = fmaf(yv.x, dz[mt][e], dot);
                  dot = fmaf(yv.y, dz[mt][2 + e], dot);
                }
                ds_head[h] += dot;
                // value gradient: the thread owns these TPC cells of the slot (or adds to d_y itself if the column is unbound)
#pragma unroll
                for (int v4 = 0; v4 < TM_TPC / 4; ++v4) {
                  const float4 add = make_float4(dy[2 * v4][e], dy[2 * v4][2 + e], dy[2 * v4 + 1][e], dy[2 * v4 + 1][2 + e]);
                  if (bound[e]) {
                    float4* c4 = reinterpret_cast<float4*>(S.slot_acc + soff[e][v4] + h * W);
                    float4 cur = *c4;
                    cur.x += add.x, cur.y += add.y, cur.z += add.z, cur.w += add.w;
                    *c4 = cur;
                  } else {
                    atomicAdd(reinterpret_cast<float4*>(P.d_y + yoff[e] + h * P.C + 4 * v4), add);
                  }
                }
              }
            }
            __syncwarp();  // the lanes of a quad may reach the same slot cell in the next group of candidates
          }
        }
      }
    }
    __syncthreads();
  }
  for (int sidx = 0; sidx < P.n_slots; ++sidx)
    if (S.slot_j[sidx] >= 0) tm_flush_slot<NH>(P, S, W, log2c, sidx);

  // ---- parameter gradients: one reduction per CTA ----
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float v = warp_sum(ds_head[h]);
    __syncthreads();
    if (lane == 0) S.red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      float acc = 0.f;
      for (int w = 0; w < nwarps; ++w) acc += S.red[w];
      atomicAdd(P.d_scale + h, -acc);
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
      float acc = 0.f;
      for (int w = 0; w < nwarps; ++w) acc += S.red[w];
      atomicAdd(P.d_b2 + o, acc);
    }
  }
  // b1 and W2: per-thread partials (channel c0 + i) -> CTA sums in shared memory -> one RED per address
  for (int i = 0; i < TM_TPC; ++i) {
    atomicAdd(&S.gpar[c0 + i], S.priv[i * blockDim.x + tid]);
    for (int o = 0; o < n_out; ++o) atomicAdd(&S.gpar[(1 + o) * P.C + c0 + i], S.priv[((1 + o) * TM_TPC + i) * blockDim.x + tid]);
  }
  __syncthreads();
  for (int i = tid; i < n_par; i += blockDim.x) atomicAdd(i < P.C ? P.d_b1 + i : P.d_w2 + (i - P.C), S.gpar[i]);
}

}  // namespace pit
