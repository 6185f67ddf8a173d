// K9: masked position-attention over PER-SAMPLE meshes (posatt_cross, pit.py:59-71, as train_naca.py:47-61 and
// train_elasticity.py:41-54 use it), tiled by sample and by 64 consecutive rows.
//
// The generic warp-per-row kernels (local_attention.cuh) gather every kept value row from L2 for every output row (NACA decoder:
// 225 k rows x 15 neighbours x 512 B = 1.7 GB per pass) and run the value gradient as a second sweep with the roles of rows and
// columns exchanged (728 columns x 11 271 rows x 20 samples of distance tests).  Consecutive rows of a mesh are neighbours in
// space, so the kept sets of a tile of rows overlap almost completely.  A CTA owns (sample, 64 consecutive rows):
//   A  the rows' NEIGHBOUR LISTS -- the columns with d2 <= v_hi, a superset of what any head keeps, <= 32 entries {j, d2} per row,
//      written by the rowstat kernel from the sweep it runs anyway (pit_rowstat_lists) -- are loaded: one entry per lane later on.
//      No kernel of the stage sweeps the N x M pairs again (the generic path sweeps them in forward, scale gradient and, with
//      the roles exchanged, value gradient: 164 M pair tests each at the NACA decoder);
//   B  the union of the listed columns is bound to shared-memory SLOTS and their value rows are staged once (cp.async);
//   C  forward: out_i = sum_k P_ik U[slot_k] / l_i from shared memory.
//      backward: dP_ik = <dO_i, U[slot_k]>, delta_i, dS_ik -> ds_h (registers), and dU accumulates p^_ik dO_i into per-slot
//      shared-memory accumulators (shared atomics), flushed once per CTA with vector REDs: ~30x fewer global atomics than a
//      per-pair scatter and no second sweep.
// Columns beyond the slot capacity fall back to global gathers / REDs (correct, slower).  A row may keep at most 32 columns
// (the caller checks rank_hi + 1 <= 24, leaving room for ties at the cut); a longer list poisons the row with NaN rather than
// dropping neighbours silently.
#pragma once
#include "geometry.cuh"

namespace pit {

constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_ROWS = 64;   // rows per CTA
constexpr int ST_KMAX = 32;   // list entries per row = lanes
constexpr int ST_MAX_RANK = 23;

struct SampleTileParams {
  const float* values;    // [B,M,D]
  const float* scale;     // [H]
  const float* v_min;     // [B,N]
  const float* v_lo;
  const float* v_hi;
  float weight;
  const int16_t* nbr_idx; // [B,N,32] neighbour lists of pit_rowstat_lists: the columns with d2 <= v_hi, ascending
  const float* nbr_d2;    // [B,N,32]
  const int32_t* nbr_cnt; // [B,N] true count (> 32: the list is incomplete)
  int B, H, N, M, D, sd;
  int n_slots;            // slot capacity (shared memory)
  // forward
  float* out;
  int64_t ld_out, col_off;
  float* rowsum;          // [B,H,N]
  // backward
  const float* d_out;
  float* d_values;        // [B,M,D] zero-initialised, or null
  float* d_scale;         // [H] zero-initialised, or null
};

__host__ __device__ inline size_t sample_tile_smem_bytes(int M, int D, int n_slots, bool backward) {
  auto up = [](size_t b) { return (b + 15) / 16 * 16; };
  return up((size_t)M * 2) + up((size_t)n_slots * 2) + up((size_t)ST_ROWS * ST_KMAX * 2) + up((size_t)ST_ROWS * ST_KMAX * 4) +
         up((size_t)ST_ROWS * 4) + 32 + (size_t)n_slots * D * 4 * (backward ? 2 : 1);
}

__device__ __forceinline__ void st_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// DC: float4 chunks of the D-wide vector a lane owns (chunk c of lane l covers elements 4 (l + 32 c) .. +3)
template <int NH, int DC, bool BWD>
__global__ void __launch_bounds__(ST_THREADS) sample_tile_kernel(const SampleTileParams P) {
  extern __shared__ __align__(16) unsigned char st_raw[];
  const int M = P.M, D = P.D, NS = P.n_slots;
  unsigned char* cur = st_raw;
  auto carve = [&](size_t bytes) {
    unsigned char* at = cur;
    cur += (bytes + 15) / 16 * 16;
    return at;
  };
  int16_t* slot_of = reinterpret_cast<int16_t*>(carve((size_t)M * 2));
  int16_t* slot_col = reinterpret_cast<int16_t*>(carve((size_t)NS * 2));
  int16_t* LJ = reinterpret_cast<int16_t*>(carve((size_t)ST_ROWS * ST_KMAX * 2));
  float* LD2 = reinterpret_cast<float*>(carve((size_t)ST_ROWS * ST_KMAX * 4));
  int* LCNT = reinterpret_cast<int*>(carve((size_t)ST_ROWS * 4));
  int* counters = reinterpret_cast<int*>(carve(16));       // [0]: slots handed out
  float* dsacc = reinterpret_cast<float*>(carve(16));      // [NH]
  float* US = reinterpret_cast<float*>(carve((size_t)NS * D * 4));
  float* ACC = reinterpret_cast<float*>(cur);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, row0 = blockIdx.x * ST_ROWS;
  const int rows = min(ST_ROWS, P.N - row0);
  const float* vals = P.values + (int64_t)b * M * D;
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);

  for (int j = tid; j < M; j += ST_THREADS) slot_of[j] = -1;
  if (tid < 4) counters[tid] = 0;
  if (tid < NH) dsacc[tid] = 0.f;
  __syncthreads();

  // ---- A: the rows' neighbour lists (built once per forward by the rowstat kernel, whose sweep they come from for free) ----
  for (int r = warp; r < rows; r += ST_WARPS) {
    const int64_t gi = (int64_t)b * P.N + row0 + r;
    const int cnt = __ldg(P.nbr_cnt + gi);
    if (lane < min(cnt, ST_KMAX)) {
      const int j = __ldg(P.nbr_idx + gi * ST_KMAX + lane);
      LJ[r * ST_KMAX + lane] = (int16_t)j;
      LD2[r * ST_KMAX + lane] = __ldg(P.nbr_d2 + gi * ST_KMAX + lane);
      slot_of[j] = -2;      // used by this tile (benign race: every writer stores the same value)
    }
    if (lane == 0) LCNT[r] = cnt;
  }
  __syncthreads();
  // ---- B: bind the used columns to slots, stage their value rows ----
  for (int j = tid; j < M; j += ST_THREADS) {
    if (slot_of[j] == -2) {
      const int sl = atomicAdd(&counters[0], 1);
      if (sl < NS) {
        slot_of[j] = (int16_t)sl;
        slot_col[sl] = (int16_t)j;
      } else {
        slot_of[j] = -1;       // no room: this column is read from / accumulated to global memory
      }
    }
  }
  __syncthreads();
  const int used = min(counters[0], NS);
  const int dv = D / 4;
  for (int c = tid; c < used * dv; c += ST_THREADS) {
    const int sl = c / dv, q = c - sl * dv;
    st_cp16(US + (size_t)sl * D + 4 * q, vals + (int64_t)slot_col[sl] * D + 4 * q);
    if (BWD) *reinterpret_cast<float4*>(ACC + (size_t)sl * D + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- C: one warp per row, one list entry per lane ----
  float ds_part[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) ds_part[h] = 0.f;
  for (int r = warp; r < rows; r += ST_WARPS) {
    const int64_t gi = (int64_t)b * P.N + row0 + r;
    const int cnt_all = LCNT[r];
    const int cnt = min(cnt_all, ST_KMAX);
    const int j = lane < cnt ? (int)LJ[r * ST_KMAX + lane] : -1;
    const float d2 = lane < cnt ? LD2[r * ST_KMAX + lane] : 0.f;
    const int slot = j >= 0 ? (int)slot_of[j] : -1;
    const float v_min = __ldg(P.v_min + gi), v_lo = __ldg(P.v_lo + gi), v_hi = __ldg(P.v_hi + gi);
    const float poison = cnt_all > ST_KMAX ? __int_as_float(0x7fc00000) : 0.f;   // a list that does not fit is not silently cut
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float sc = __fmul_rn(d2, s[h]);
      const bool kept = j >= 0 && sc <= head_threshold(v_lo, v_hi, s[h], P.weight);
      const float p = kept ? expf(__fsub_rn(__fmul_rn(v_min, s[h]), sc)) : 0.f;
      const float l = warp_sum(p);
      const unsigned live = __ballot_sync(FULL, p != 0.f);
      if (!BWD) {
        float4 acc[DC];
#pragma unroll
        for (int c = 0; c < DC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        // uniform trip count (entries not kept by this head carry weight 0): the loop unrolls, so the shuffles and shared-memory
        // loads of several entries are in flight together
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
          const float pk = __shfl_sync(FULL, p, k);
          const int sk = __shfl_sync(FULL, slot, k);       // warp-uniform: the usual case (a slot) stays on 32-bit shared addressing
          const int jk = sk >= 0 ? 0 : __shfl_sync(FULL, j, k);
#pragma unroll
          for (int c = 0; c < DC; ++c) {
            const int e = 4 * (lane + 32 * c);
            if (e < D) {
              const float4 v = sk >= 0 ? *reinterpret_cast<const float4*>(US + (sk * D + e))
                                       : __ldg(reinterpret_cast<const float4*>(vals + (int64_t)jk * D + e));
              acc[c].x = fmaf(pk, v.x, acc[c].x), acc[c].y = fmaf(pk, v.y, acc[c].y);
              acc[c].z = fmaf(pk, v.z, acc[c].z), acc[c].w = fmaf(pk, v.w, acc[c].w);
            }
          }
        }
        const float inv = 1.f / l;
        float* dst = P.out + gi * P.ld_out + P.col_off + (int64_t)h * D;
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          const int e = 4 * (lane + 32 * c);
          if (e < D) *reinterpret_cast<float4*>(dst + e) = make_float4(acc[c].x * inv + poison, acc[c].y * inv + poison, acc[c].z * inv + poison, acc[c].w * inv + poison);
        }
        if (lane == 0) P.rowsum[((int64_t)b * P.H + h) * P.N + row0 + r] = l;
      } else {
        float4 g[DC];
        const float* go = P.d_out + gi * P.ld_out + P.col_off + (int64_t)h * D;
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          const int e = 4 * (lane + 32 * c);
          g[c] = e < D ? __ldg(reinterpret_cast<const float4*>(go + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // dP of every entry, sixteen at a time: each lane forms its partial dot products with all sixteen value rows (independent
        // shared-memory loads), then a reduce-scatter butterfly (8 + 4 + 2 + 1 + 1 shuffles) leaves entry k's total on lane k --
        // instead of one five-step warp reduction per entry
        float dp = 0.f;
        for (int base = 0; base < cnt; base += 16) {
          float v[16];
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const int k = base + kk;      // warp-uniform
            float t = 0.f;
            if (k < cnt) {
              const int sk = __shfl_sync(FULL, slot, k);
              const int jk = sk >= 0 ? 0 : __shfl_sync(FULL, j, k);
#pragma unroll
              for (int c = 0; c < DC; ++c) {
                const int e = 4 * (lane + 32 * c);
                if (e < D) {
                  const float4 u = sk >= 0 ? *reinterpret_cast<const float4*>(US + (sk * D + e))
                                           : __ldg(reinterpret_cast<const float4*>(vals + (int64_t)jk * D + e));
                  t = fmaf(g[c].x, u.x, t), t = fmaf(g[c].y, u.y, t), t = fmaf(g[c].z, u.z, t), t = fmaf(g[c].w, u.w, t);
                }
              }
            }
            v[kk] = t;
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; ++i) {
              const float keep = upper ? v[i + o] : v[i], send = upper ? v[i] : v[i + o];
              v[i] = keep + __shfl_xor_sync(FULL, send, o);
            }
          }
          v[0] += __shfl_xor_sync(FULL, v[0], 16);      // lanes L and L ^ 16 both hold the total of entry base + (L & 15)
          if ((lane >> 4) == (base >> 4)) dp = v[0];
        }
        const float ph = p / l;
        const float delta = warp_sum(ph * dp);
        const float dS = ph * (dp - delta);
        ds_part[h] = fmaf(-dS, d2, ds_part[h]) + poison;
        if (P.d_values) {
          for (unsigned m = live; m; m &= m - 1) {      // entries without weight add nothing: skipped (their atomics are not free)
            const int k = __ffs(m) - 1;
            const float pk = __shfl_sync(FULL, ph, k);
            const int sk = __shfl_sync(FULL, slot, k);
            const int jk = sk >= 0 ? 0 : __shfl_sync(FULL, j, k);
#pragma unroll
            for (int c = 0; c < DC; ++c) {
              const int e = 4 * (lane + 32 * c);
              if (e < D) {
                if (sk >= 0) {
                  // accumulator layout [slot][chunk][component][lane]: the lanes of one atomic hit consecutive banks
                  const int lw = min(32, dv - 32 * c);      // lanes that own a piece of this chunk
                  float* dst = ACC + (sk * D + 128 * c + lane);
                  atomicAdd(dst, pk * g[c].x), atomicAdd(dst + lw, pk * g[c].y), atomicAdd(dst + 2 * lw, pk * g[c].z), atomicAdd(dst + 3 * lw, pk * g[c].w);
                } else {
                  atomicAdd(reinterpret_cast<float4*>(P.d_values + ((int64_t)b * M + jk) * D + e),
                            make_float4(pk * g[c].x, pk * g[c].y, pk * g[c].z, pk * g[c].w));
                }
              }
            }
          }
        }
      }
    }
  }
  if (BWD) {
    // ---- D: flush the slot accumulators and the scale gradient ----
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float v = warp_sum(ds_part[h]);
      if (lane == 0) atomicAdd(dsacc + h, v);
    }
    __syncthreads();
    if (P.d_scale && tid < NH) atomicAdd(P.d_scale + tid, dsacc[tid]);
    if (P.d_values) {
      for (int c = tid; c < used * dv; c += ST_THREADS) {
        const int sl = c / dv, q = c - sl * dv;          // q = lane + 32 chunk: the float4 of elements 4q .. 4q+3
        const int lw = min(32, dv - 32 * (q >> 5));
        const float* a = ACC + (size_t)sl * D + 128 * (q >> 5) + (q & 31);
        atomicAdd(reinterpret_cast<float4*>(P.d_values + ((int64_t)b * M + slot_col[sl]) * D + 4 * q), make_float4(a[0], a[lw], a[2 * lw], a[3 * lw]));
      }
    }
  }
}

}  // namespace pit
