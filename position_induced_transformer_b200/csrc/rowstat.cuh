// K0 rowstat: exact order statistics of the squared distances of every row.
//
// Replaces the full row sort inside torch.quantile (pit.py:49, 136, 197, 255): the mask only
// needs the k_lo-th and k_hi-th smallest d2 of a row (head independent, because multiplying by
// the positive per-head scale and rounding is monotone) plus the row minimum, which is the
// soft-max shift.  Selection is a radix select on the fp32 bit pattern (d2 >= +0, so unsigned
// integer order is numeric order); nothing of size N x M ever reaches HBM.
//
// Two mappings:
//   warp-per-row   M <= 1024: the row's d2 live in registers, 32 one-bit ballot/REDUX steps;
//   block-per-row  larger M: 4 passes of 8-bit digits with a shared-memory histogram,
//                  d2 recomputed on every pass from the (L1/L2 resident) coordinates.
#pragma once
#include "geometry.cuh"

namespace pit {

struct RowstatParams {
  const float* mesh_out;
  const float* mesh_in;
  const float* period;
  float* v_min;
  float* v_lo;
  float* v_hi;
  int rows_total;  // (mesh_batched ? B : 1) * N
  int N, M, sd, mesh_batched;
  int k_lo, k_hi;
  // optional neighbour lists (warp-per-row mapping only, M <= 1024): the columns with d2 <= v_hi of every row -- a superset of
  // what any head keeps -- as up to 32 {column, d2} entries in ascending column order, plus the true count (may exceed 32)
  int16_t* nbr_idx;  // [rows_total, 32] or null
  float* nbr_d2;     // [rows_total, 32]
  int32_t* nbr_cnt;  // [rows_total]
};

// k_lo-th and k_hi-th (k_hi in {k_lo, k_lo + 1}) smallest of the warp's 32 x RR keys (unused entries 0xffffffff).
template <int RR>
__device__ __forceinline__ void rowstat_select(const uint32_t (&key)[RR], int k_lo, int k_hi, uint32_t& lo_out, uint32_t& hi_out) {
  // k_lo-th smallest key, most significant bit first.  `cand` counts the keys that still match the decided
  // bits; once a single candidate is left it is the answer and the remaining bit steps are skipped (on meshes
  // without exact ties this happens after ~log2(M) + a few steps instead of 32).
  uint32_t prefix = 0;
  int k = k_lo;
  int cand = RR * 32;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    int c = 0;
#pragma unroll
    for (int r = 0; r < RR; ++r) c += (((key[r] ^ prefix) >> bit) == 0u) ? 1 : 0;
    c = __reduce_add_sync(FULL, c);
    if (k >= c) {
      k -= c;
      prefix |= 1u << bit;
      cand -= c;
    } else {
      cand = c;
    }
    if (cand == 1 && bit > 0) {
      // the unique key that agrees with `prefix` on bits [bit, 31]
      uint32_t mine = 0;
#pragma unroll
      for (int r = 0; r < RR; ++r)
        if (((key[r] ^ prefix) >> bit) == 0u) mine = key[r];
      prefix = __reduce_or_sync(FULL, mine);
      break;
    }
  }
  const uint32_t lo = prefix;
  uint32_t hi = lo;
  if (k_hi != k_lo) {  // k_hi == k_lo + 1
    int le = 0;
    uint32_t gt = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < RR; ++r) {
      le += (key[r] <= lo) ? 1 : 0;
      if (key[r] > lo) gt = min(gt, key[r]);
    }
    le = __reduce_add_sync(FULL, le);
    gt = __reduce_min_sync(FULL, gt);
    if (k_hi >= le) hi = gt;
  }
  lo_out = lo;
  hi_out = hi;
}

template <int GEO, int R>
__global__ void __launch_bounds__(128) rowstat_warp_kernel(const RowstatParams P) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= P.rows_total) return;
  const int bm = P.mesh_batched ? row / P.N : 0;
  const float* mesh_in = P.mesh_in + (int64_t)bm * P.M * P.sd;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const Point<GEO> o = load_point<GEO>(P.mesh_out, row, P.sd);

  uint32_t key[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = r * 32 + lane;
    key[r] = 0xffffffffu;
    if (j < P.M) key[r] = __float_as_uint(dist2<GEO>(o, load_point<GEO>(mesh_in, j, P.sd), period));
  }
  uint32_t mn = key[0];
#pragma unroll
  for (int r = 1; r < R; ++r) mn = min(mn, key[r]);
  mn = __reduce_min_sync(FULL, mn);

  uint32_t lo, hi;
  bool done = false, listed = false;
  // Small ranks (the locality masks keep a few dozen columns at most): shrink the problem before selecting.  The
  // (k_hi+1)-th smallest of the 32 per-lane minima is an upper bound T of the k_hi-th smallest key (those lanes hold
  // k_hi+1 distinct keys <= T), so only keys <= T can matter; a lane usually holds 0..3 of them.  They are compacted
  // into four registers per lane and the bit-serial selection runs on 4 instead of R keys per lane.  Rows where some
  // lane holds more than four (heavy ties) take the full path below -- same result either way.
  if (P.k_hi < 32) {
    uint32_t lm = key[0];
#pragma unroll
    for (int r = 1; r < R; ++r) lm = min(lm, key[r]);
    int rank = 0;  // position of this lane's minimum among the 32 (ties broken by lane index)
    for (int l = 0; l < 32; ++l) {
      const uint32_t other = __shfl_sync(FULL, lm, l);
      rank += (other < lm || (other == lm && l < lane)) ? 1 : 0;
    }
    const unsigned owner = __ballot_sync(FULL, rank == P.k_hi);
    const uint32_t bound = __shfl_sync(FULL, lm, __ffs(owner) - 1);
    if (bound != 0xffffffffu) {  // fewer than k_hi+1 lanes with a valid key: no bound, full path
      uint32_t c[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      int cr[4] = {0, 0, 0, 0};   // register index r of each candidate (its column is 32 r + lane)
      int n = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (key[r] <= bound) {
          c[3] = c[2], c[2] = c[1], c[1] = c[0], c[0] = key[r];
          cr[3] = cr[2], cr[2] = cr[1], cr[1] = cr[0], cr[0] = r;
          ++n;
        }
      }
      if (!__any_sync(FULL, n > 4)) {
        const int total = __reduce_add_sync(FULL, n);
        if (total <= 32) {
          // At most 32 candidates in the whole row (the usual case: k_hi + 1 of them plus a few): one per lane, ranked by
          // counting -- 32 shuffles instead of a 32-step bit-serial selection -- and the row's neighbour list is the same set.
          __shared__ uint32_t cand_key[4][32];
          __shared__ int16_t cand_col[4][32];
          const int w = threadIdx.x >> 5;
          int before = n;   // inclusive scan of n over the lanes
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, before, o);
            if (lane >= o) before += v;
          }
          before -= n;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < n) {
              cand_key[w][before + q] = c[q];
              cand_col[w][before + q] = (int16_t)(cr[q] * 32 + lane);
            }
          __syncwarp();
          const uint32_t mine = lane < total ? cand_key[w][lane] : 0xffffffffu;
          const int col = lane < total ? (int)cand_col[w][lane] : 0;
          __syncwarp();       // the buffers are reused by the warp's next row
          int rank = 0;
          for (int l = 0; l < 32; ++l) {
            const uint32_t other = __shfl_sync(FULL, mine, l);
            rank += (other < mine || (other == mine && l < lane)) ? 1 : 0;
          }
          const unsigned at_lo = __ballot_sync(FULL, lane < total && rank == P.k_lo);
          const unsigned at_hi = __ballot_sync(FULL, lane < total && rank == P.k_hi);
          lo = __shfl_sync(FULL, mine, __ffs(at_lo) - 1);
          hi = __shfl_sync(FULL, mine, __ffs(at_hi) - 1);
          done = true;
          if (P.nbr_idx) {
            const bool in_list = lane < total && mine <= hi;
            const unsigned m = __ballot_sync(FULL, in_list);
            if (in_list) {
              const int pos = __popc(m & ((1u << lane) - 1u));
              P.nbr_idx[(int64_t)row * 32 + pos] = (int16_t)col;
              P.nbr_d2[(int64_t)row * 32 + pos] = __uint_as_float(mine);
            }
            if (lane == 0) P.nbr_cnt[row] = __popc(m);
            listed = true;
          }
        } else {
          rowstat_select<4>(c, P.k_lo, P.k_hi, lo, hi);
          done = true;
        }
      }
    }
  }
  if (!done) rowstat_select<R>(key, P.k_lo, P.k_hi, lo, hi);
  if (lane == 0) {
    P.v_min[row] = __uint_as_float(mn);
    P.v_lo[row] = __uint_as_float(lo);
    P.v_hi[row] = __uint_as_float(hi);
  }
  if (P.nbr_idx && !listed) {
    // the keys are still in registers: one more pass compacts the candidates of the row
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r * 32 >= P.M) break;             // warp-uniform: nothing but padding from here on
      const bool cand = key[r] <= hi;       // padding keys are 0xffffffff > hi
      const unsigned m = __ballot_sync(FULL, cand);
      if (cand) {
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (pos < 32) {
          P.nbr_idx[(int64_t)row * 32 + pos] = (int16_t)(r * 32 + lane);
          P.nbr_d2[(int64_t)row * 32 + pos] = __uint_as_float(key[r]);
        }
      }
      cnt += __popc(m);
    }
    if (lane == 0) P.nbr_cnt[row] = cnt;
  }
}

constexpr int ROWSTAT_BLOCK = 256;

template <int GEO>
__global__ void __launch_bounds__(ROWSTAT_BLOCK) rowstat_block_kernel(const RowstatParams P) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix, s_min, s_gt;
  __shared__ int s_k, s_le;
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const int bm = P.mesh_batched ? row / P.N : 0;
  const float* mesh_in = P.mesh_in + (int64_t)bm * P.M * P.sd;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const Point<GEO> o = load_point<GEO>(P.mesh_out, row, P.sd);
  // Every thread runs the same number of iterations so the warp-wide match below is convergent.
  const int m_pad = (P.M + ROWSTAT_BLOCK - 1) / ROWSTAT_BLOCK * ROWSTAT_BLOCK;

  if (tid == 0) {
    s_prefix = 0;
    s_k = P.k_lo;
    s_min = 0xffffffffu;
    s_gt = 0xffffffffu;
    s_le = 0;
  }
  uint32_t my_min = 0xffffffffu;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int j = tid; j < m_pad; j += ROWSTAT_BLOCK) {
      bool take = j < P.M;
      uint32_t key = 0;
      if (take) {
        key = __float_as_uint(dist2<GEO>(o, load_point<GEO>(mesh_in, j, P.sd), period));
        if (pass == 0)
          my_min = min(my_min, key);
        else
          take = ((key ^ prefix) >> (shift + 8)) == 0u;
      }
      const uint32_t digit = (key >> shift) & 255u;
      // warp-aggregated histogram update: one atomic per distinct digit in the warp
      const unsigned active = __ballot_sync(FULL, take);
      if (take) {
        const unsigned peers = __match_any_sync(active, digit);
        if ((__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[digit], __popc(peers));
      }
    }
    __syncthreads();
    if (tid == 0) {
      int k = s_k, cum = 0, d = 0;
      for (; d < 255; ++d) {
        if (k < cum + hist[d]) break;
        cum += hist[d];
      }
      s_k = k - cum;
      s_prefix = prefix | ((uint32_t)d << shift);
    }
    __syncthreads();
  }
  const uint32_t lo = s_prefix;
  my_min = __reduce_min_sync(FULL, my_min);
  if ((tid & 31) == 0) atomicMin(&s_min, my_min);
  if (P.k_hi != P.k_lo) {
    int le = 0;
    uint32_t gt = 0xffffffffu;
    for (int j = tid; j < P.M; j += ROWSTAT_BLOCK) {
      const uint32_t key = __float_as_uint(dist2<GEO>(o, load_point<GEO>(mesh_in, j, P.sd), period));
      le += (key <= lo) ? 1 : 0;
      if (key > lo) gt = min(gt, key);
    }
    le = __reduce_add_sync(FULL, le);
    gt = __reduce_min_sync(FULL, gt);
    if ((tid & 31) == 0) {
      atomicAdd(&s_le, le);
      atomicMin(&s_gt, gt);
    }
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t hi = lo;
    if (P.k_hi != P.k_lo && P.k_hi >= s_le) hi = s_gt;
    P.v_min[row] = __uint_as_float(s_min);
    P.v_lo[row] = __uint_as_float(lo);
    P.v_hi[row] = __uint_as_float(hi);
  }
}

}  // namespace pit
