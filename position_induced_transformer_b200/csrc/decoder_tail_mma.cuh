// K3m: fused decoder tail on the warp-level tensor-core path (mma.sync m16n8k8, 3xTF32) for shared meshes.
//
// Same contract as decoder_tail.cuh (pit.decoder, pit.py:124-127: cross position-attention `up` + `de` MLP), but
// the contraction of a row's 6..40 kept weights with the B*C-wide latent rows is no longer a per-row gather on the
// CUDA cores.  Rows are taken in tiles of 16 consecutive points.  On a spatially coherent mesh the union of the
// columns kept by ANY row of a tile is barely larger than one row's kept set (12..16 columns at Darcy-421), so
//
//     pre^T [(b,c) x 16 rows] = Y^T [(b,c) x candidates] . P^T [candidates x 16 rows]        (per head, summed)
//
// is a small dense product at ~50 % density -- worth moving to the tensor pipe, which has ~8x the FMA rate of the
// fp32 pipe even after the 3xTF32 split.  The mask itself is still decided exactly as in the SIMT kernels (bit-exact
// d2, per-head cut, fp32 weights); only the products are split:  hi*hi + lo*hi + hi*lo with fp32 accumulation
// (error ~2^-21, inside the 1e-5 parity budget).
//
// Work split.  A CTA (4 warps) walks its tiles in rounds of 4:
//   phase 1  warp w "prepares" tile w of the round: scans the M columns held in registers against the 16 rows
//            (head-independent d2-space pre-filter), ballot-compacts the candidate list, evaluates the unnormalised
//            weights P[h][row][cand] (and d2) into a shared-memory block, and the row sums;
//   phase 2  for every prepared tile, warp w owns the 128-column chunk(s) w, w+4, ... of the B*C-wide hidden vector:
//            operand A = Y^T straight from global/L1 (each thread reads 16 contiguous floats of two candidate rows:
//            the column order inside an m16 tile is permuted so that fragments are 128-bit loads), operand B = P^T
//            from the shared block (conflict-free with the 36-float pitch), accumulators [128 cols x 16 rows] in
//            64 registers; epilogue = bias + exact GELU + C->O projection on the fragments.
// The transposed orientation (columns on the MMA m axis, rows on n) is what makes the backward cheap: the
// accumulator fragment of g1^T is, register for register, the A fragment of dY^T = g1^T . P (reduction over rows).
//
// Tiles whose candidate list exceeds one 32-column block (incoherent meshes, unmasked decoders) are handled by
// recomputing further blocks on the fly -- slower, same results.
#pragma once
#include "decoder_tail.cuh"

namespace pit {

constexpr int TM_ROWS = 16;    // rows per tile (two n8 MMA tiles)
constexpr int TM_KT = 32;      // candidates per weight block
constexpr int TM_LD = 36;      // pitch of a block row in floats: fragment reads hit 32 distinct banks
constexpr int TM_CHUNK = 128;  // hidden-vector columns per warp pass (eight m16 MMA tiles)
constexpr int TM_ROUND = TALL_WARPS;  // tiles prepared per round (one per warp)

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// 3xTF32 split.  The tensor pipe reads the upper 19 bits of an fp32 register, so the raw value serves as the
// (truncated) high part and only the residual needs arithmetic; weights use a rounded high part (half the residual).
__device__ __forceinline__ uint32_t tm_trunc_lo(float x) { return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u)); }
__device__ __forceinline__ float tm_round_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

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

// Candidate list of a tile: every column within the pre-filter radius of at least one of its rows, in column order.
template <int GEO, int CPL, int NH>
__device__ __forceinline__ int tm_candidates(const TmRow<GEO, NH>& R, const Point<GEO> (&col)[CPL], int M, int lane, float period,
                                             int16_t* cand) {
  uint32_t flags = 0;
  for (int i = 0; i < TM_ROWS; ++i) {
    Point<GEO> o;
    o.x = __shfl_sync(FULL, R.o.x, i);
    o.y = __shfl_sync(FULL, R.o.y, i);
    const float vc = __shfl_sync(FULL, R.vcap, i);
#pragma unroll
    for (int c = 0; c < CPL; ++c) flags |= (dist2<GEO>(o, col[c], period) <= vc) ? (1u << c) : 0u;
  }
  int n = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    const bool f = ((flags >> c) & 1u) && j < M;
    const unsigned m = __ballot_sync(FULL, f);
    if (f) cand[n + __popc(m & lt)] = (int16_t)j;
    n += __popc(m);
  }
  __syncwarp();
  return n;
}

// One block (<= 32 candidates, zero-filled to the next multiple of 8) of unnormalised weights:
// lane -> (row lane&15, candidates of parity lane>>4).  pt = [NH][16][TM_LD], d2t = [16][TM_LD] or null.
template <int GEO, int NH>
__device__ __forceinline__ void tm_block(const TailParams& P, const TmRow<GEO, NH>& R, const float (&s)[NH], float period,
                                         const int16_t* cand, int cnt, int kb, int lane, float* pt, float* d2t, float (&psum)[NH],
                                         float (&pdsum)[NH]) {
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
        if (sc <= R.cut[h]) p = __expf(__fsub_rn(R.top[h], sc));
      }
      psum[h] += p;
      pdsum[h] = fmaf(p, d2, pdsum[h]);
      pt[(h * TM_ROWS + i) * TM_LD + k] = p;
    }
    if (d2t) d2t[i * TM_LD + k] = live ? d2 : 0.f;
  }
}

// Shared-memory image of a prepared tile.
template <int NH, bool BWD>
struct TmTile {
  float p[NH][TM_ROWS][TM_LD];              // unnormalised weights of block 0 (or of the block being processed)
  float d2[BWD ? TM_ROWS : 1][TM_LD];       // squared distances (backward only)
  float inv_l[NH][TM_ROWS];                 // 1 / row sum (0 for rows past the end)
  float m[NH][TM_ROWS];                     // backward: sum_j P^ d2
  int cnt;
  int pad[3];
};

// pre^T += Y_h^T . P_h^T for one block of one tile and one 128-column chunk.
//   acc[mt][nt][.]: column 16*g + mt (+8 for registers 2,3) of the chunk, rows 8*nt + 2*t (+1 for registers 1,3).
// y_chunk points at Y[b, 0, 0, c] for this thread's 16 columns; row j of head h sits NH*C*j + h*C floats further.
template <int NH>
__device__ __forceinline__ void tm_mma_block(float (&acc)[8][2][4], const float (*p)[TM_ROWS][TM_LD], const float (*inv_l)[TM_ROWS],
                                             const int16_t* cand, int cnt, int kb, const float* y_chunk, int C, int g, int t) {
  const int base = kb * TM_KT;
  const int ksteps = (min(cnt - base, TM_KT) + 7) >> 3;
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float inv0 = inv_l[h][g], inv1 = inv_l[h][g + 8];
    for (int ks = 0; ks < ksteps; ++ks) {
      const int ka = base + ks * 8 + t, kc = ka + 4;
      const int ja = ka < cnt ? (int)cand[ka] : 0, jb = kc < cnt ? (int)cand[kc] : 0;
      const float4* ra = reinterpret_cast<const float4*>(y_chunk + ((size_t)ja * NH + h) * C);
      const float4* rb = reinterpret_cast<const float4*>(y_chunk + ((size_t)jb * NH + h) * C);
      float ya[16], yb[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 a = __ldg(ra + v), b = __ldg(rb + v);
        ya[4 * v] = a.x, ya[4 * v + 1] = a.y, ya[4 * v + 2] = a.z, ya[4 * v + 3] = a.w;
        yb[4 * v] = b.x, yb[4 * v + 1] = b.y, yb[4 * v + 2] = b.z, yb[4 * v + 3] = b.w;
      }
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float inv = nt == 0 ? inv0 : inv1;
        const float p0 = p[h][g + 8 * nt][ks * 8 + t] * inv, p1 = p[h][g + 8 * nt][ks * 8 + t + 4] * inv;
        const float h0 = tm_round_hi(p0), h1 = tm_round_hi(p1);
        bh[nt][0] = __float_as_uint(h0), bh[nt][1] = __float_as_uint(h1);
        bl[nt][0] = __float_as_uint(p0 - h0), bl[nt][1] = __float_as_uint(p1 - h1);
      }
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        const uint32_t ah[4] = {__float_as_uint(ya[mt]), __float_as_uint(ya[8 + mt]), __float_as_uint(yb[mt]), __float_as_uint(yb[8 + mt])};
        const uint32_t al[4] = {tm_trunc_lo(ya[mt]), tm_trunc_lo(ya[8 + mt]), tm_trunc_lo(yb[mt]), tm_trunc_lo(yb[8 + mt])};
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma_tf32_16x8x8(acc[mt][nt], al, bh[nt]);
          mma_tf32_16x8x8(acc[mt][nt], ah, bl[nt]);
          mma_tf32_16x8x8(acc[mt][nt], ah, bh[nt]);
        }
      }
    }
  }
}

__host__ __device__ inline size_t tm_align(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline size_t tm_cand_bytes(int M) { return tm_align((size_t)M * 2); }
template <int NH, bool BWD>
__host__ __device__ inline size_t tm_tiles_bytes(int M) {
  return 2 * TM_ROUND * (tm_align(sizeof(TmTile<NH, BWD>)) + tm_cand_bytes(M));
}
__host__ __device__ inline size_t tm_fwd_smem_bytes(int nh, int M, int C, int O) {
  return (nh == 1 ? tm_tiles_bytes<1, false>(M) : tm_tiles_bytes<2, false>(M)) + tm_align((size_t)C * (1 + O) * 4);
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <int GEO, int CPL, int NH>
__global__ void __launch_bounds__(TALL_THREADS, 2) tail_mma_fwd_kernel(const TailParams P) {
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  using Tile = TmTile<NH, false>;
  const size_t tile_stride = tm_align(sizeof(Tile));
  unsigned char* cand_base = tall_smem_raw + 2 * TM_ROUND * tile_stride;
  const size_t cand_stride = tm_cand_bytes(P.M);
  float* par = reinterpret_cast<float*>(cand_base + 2 * TM_ROUND * cand_stride);  // [b1 (C) | W2 (O x C)]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float period = P.period ? __ldg(P.period) : 0.f;
  Point<GEO> col[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
  }
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  for (int i = tid; i < P.C * (1 + P.O); i += TALL_THREADS) par[i] = i < P.C ? __ldg(P.b1 + i) : __ldg(P.w2 + (i - P.C));
  float b2r[TAIL_MAX_OUT];
#pragma unroll
  for (int o = 0; o < TAIL_MAX_OUT; ++o) b2r[o] = o < P.O ? __ldg(P.b2 + o) : 0.f;

  const int chunks = P.B * P.C / TM_CHUNK;
  const int n_tiles = (P.N + TM_ROWS - 1) / TM_ROWS;
  const int tile_begin = blockIdx.x * P.rows_per_unit;  // rows_per_unit counts tiles here
  const int tile_end = min(n_tiles, tile_begin + P.rows_per_unit);
  const int c0 = (16 * g) % P.C;          // this thread's 16 hidden channels (the same in every chunk: C divides 128)
  const int group = P.C / 16;             // lanes-of-g sharing a sample
  int round = 0;
  for (int tb = tile_begin; tb < tile_end; tb += TM_ROUND, ++round) {
    const int in_round = min(TM_ROUND, tile_end - tb);
    const int set = (round & 1) * TM_ROUND;
    // ---- phase 1: one tile per warp ----
    if (warp < in_round) {
      Tile* T = reinterpret_cast<Tile*>(tall_smem_raw + (set + warp) * tile_stride);
      int16_t* cand = reinterpret_cast<int16_t*>(cand_base + (set + warp) * cand_stride);
      const int row = (tb + warp) * TM_ROWS + (lane & 15);
      const TmRow<GEO, NH> R = tm_row<GEO, NH>(P, row, s);
      const int cnt = tm_candidates<GEO, CPL, NH>(R, col, P.M, lane, period, cand);
      float psum[NH], pdsum[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) psum[h] = pdsum[h] = 0.f;
      const int nkb = (cnt + TM_KT - 1) / TM_KT;
      for (int kb = nkb - 1; kb >= 0; --kb)  // block 0 last: it is the one left in the tile for phase 2
        tm_block<GEO, NH>(P, R, s, period, cand, cnt, kb, lane, &T->p[0][0][0], nullptr, psum, pdsum);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float l = psum[h] + __shfl_xor_sync(FULL, psum[h], 16);
        if (lane < TM_ROWS) {
          T->inv_l[h][lane] = R.valid ? 1.f / l : 0.f;
          if (R.valid) P.rowsum[(int64_t)h * P.N + row] = l;
        }
      }
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
      for (int ch0 = 0; ch0 < chunks; ch0 += TALL_WARPS) {
        const int chunk = ch0 + warp;
        const bool active = chunk < chunks;
        const int colbase = chunk * TM_CHUNK + 16 * g;
        const int b = active ? colbase / P.C : 0;
        const float* y_chunk = P.y + (size_t)b * P.M * NH * P.C + c0;
        float acc[8][2][4];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
        for (int kb = 0; kb < nkb; ++kb) {
          if (nkb > 1) {  // rare: the candidate list spans several blocks -> rebuild block kb in place (CTA-uniform branch)
            __syncthreads();
            if (warp == 0) {
              const TmRow<GEO, NH> R = tm_row<GEO, NH>(P, row0 + (lane & 15), s);
              float ps[NH], pd[NH];
#pragma unroll
              for (int h = 0; h < NH; ++h) ps[h] = pd[h] = 0.f;
              tm_block<GEO, NH>(P, R, s, period, cand, cnt, kb, lane, &T->p[0][0][0], nullptr, ps, pd);
            }
            __syncthreads();
          }
          if (active) tm_mma_block<NH>(acc, T->p, T->inv_l, cand, cnt, kb, y_chunk, P.C, g, t);
        }
        if (!active) continue;
        // epilogue: out[b, row, o] = b2[o] + sum_c W2[o, c] gelu(b1[c] + pre[c]); the thread holds channels c0..c0+15 of rows
        // 2t, 2t+1, 2t+8, 2t+9; the lanes of `group` consecutive g share the sample.
        float part[TAIL_MAX_OUT][4];
#pragma unroll
        for (int o = 0; o < TAIL_MAX_OUT; ++o)
#pragma unroll
          for (int q = 0; q < 4; ++q) part[o][q] = 0.f;
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c = c0 + mt + 8 * half;
            const float bias = par[c];
            float hid[4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) hid[nt * 2 + e] = gelu_erf(acc[mt][nt][half * 2 + e] + bias);
#pragma unroll
            for (int o = 0; o < TAIL_MAX_OUT; ++o) {
              if (o < P.O) {
                const float wv = par[(1 + o) * P.C + c];
#pragma unroll
                for (int q = 0; q < 4; ++q) part[o][q] = fmaf(wv, hid[q], part[o][q]);
              }
            }
          }
        }
#pragma unroll
        for (int o = 0; o < TAIL_MAX_OUT; ++o) {
          if (o >= P.O) break;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float vsum = part[o][q];
            for (int off = 1; off < group; off <<= 1) vsum += __shfl_xor_sync(FULL, vsum, 4 * off);
            const int row = row0 + (q >> 1) * 8 + 2 * t + (q & 1);
            if ((g % group) == 0 && row < P.N) P.out[((int64_t)b * P.N + row) * P.O + o] = vsum + b2r[o];
          }
        }
      }
    }
  }
}

}  // namespace pit
