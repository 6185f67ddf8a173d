// K1c: cooperative-CTA position-attention for shared meshes whose column set is small (M <= 1024):
// the decoder stages (N >> M), the latent self stages and small encoders.
//
// A CTA walks a contiguous range of output rows in rounds of ROUND rows.  Each round has two phases:
//
//   phase 1 (warp per row)   the M column coordinates live in registers for the CTA's whole life
//            (CPL = ceil(M/32) points per lane); a warp evaluates one row against all of them --
//            bit-exact d2, per-head scale, quantile cut, exp with the known shift -- and ballot-compacts
//            the kept (column, weight) pairs of every head into shared memory.  No global load sits
//            inside the sweep.
//   phase 2 (thread per value lane)   every thread owns one float4 lane (b, d..d+3) of the B*D-wide
//            value vector and, row after row, runs down the row's compact list: one broadcast LDS for
//            (column, weight), one 128-bit gather of the value row, four FMAs.  There are no shuffles,
//            no divergence and no atomics in the inner loop; output rows leave as full 128-bit stores.
//
// The backward kernel fuses both gradients in one pass over dO (read once, 128-bit):
//   d scale   thread-local partial of  -sum_e dO_e (W_e - m O_e)  with W = sum_j P^ d2 U, O = sum_j P^ U,
//             m = sum_j P^ d2, reduced once per CTA;
//   d values  dU[b,j,:] += P^_ij dO[b,i,h,:] accumulated WITHOUT atomics in a small set of shared-memory
//             slots (the thread owns its lane of every slot); a slot is bound to a column on first
//             touch and all slots are flushed with vector REDs (RED.E.ADD.F32x4) when the set is full
//             or the CTA ends.  Rows are visited in mesh order, so for spatially coherent meshes the
//             touched-column set stays tiny.
#pragma once
#include "geometry.cuh"

namespace pit {

constexpr int COOP_THREADS = 128;
constexpr int COOP_WARPS = COOP_THREADS / 32;
constexpr int COOP_MAX_M = 1024;
constexpr int COOP_MAX_L4 = 4;      // float4 lanes per thread: B*D <= 4 * 128 * 4 = 2048
constexpr int COOP_GATHER = 8;      // independent value-row gathers in flight per thread (forward)
constexpr int COOP_GATHER_BWD = 8;  // same, backward

struct CoopParams {
  const float* mesh_out;  // [N,sd]
  const float* mesh_in;   // [M,sd]
  const float* period;
  const float* values;  // [B,M,D]
  const float* scale;   // [H]
  const float* v_min;
  const float* v_lo;
  const float* v_hi;
  float weight;
  int masked;
  int B, H, N, M, D, sd;
  int lanes4;          // B*D/4
  int round_rows;      // rows per round (multiple of COOP_WARPS)
  int rows_per_cta;    // contiguous rows per CTA (multiple of round_rows)
  // forward
  float* out;
  int64_t ld_out, col_off;
  float* rowsum;  // [H,N]
  // backward
  const float* d_out;
  float* d_values;  // [B,M,D], zero-initialised by the host; may be null
  float* d_scale;   // [H], zero-initialised by the host; may be null
  int n_slots;
};

// Shared-memory carve-up (dynamic): per (round row, head) a segment of M entries.
struct CoopSmem {
  uint16_t* seg_j;  // [round_rows*H*M]
  float* seg_p;     // [round_rows*H*M]  weight (forward: unnormalised; backward: normalised)
  float* seg_pd;    // [round_rows*H*M]  weight * d2 (backward only)
  int* cnt;         // [round_rows*H]
  float* rowaux;    // [round_rows*H]    forward: 1/l ; backward: m = sum P^ d2
  // backward, value gradient
  float4* slot_acc;  // [n_slots*lanes4]
  int16_t* map;      // [M] column -> slot or -1
  int16_t* slot_j;   // [n_slots]
  uint8_t* touched;  // [M]
  int* ctl;          // [0] = slots in use, [1] = overflow flag
};

__host__ __device__ inline size_t coop_align(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t coop_smem_bytes(int round_rows, int H, int M, int lanes4, bool backward, int n_slots) {
  const size_t seg = (size_t)round_rows * H * M;
  size_t b = coop_align(seg * 2) + coop_align(seg * 4) + coop_align((size_t)round_rows * H * 4) * 2;
  if (backward) {
    b += coop_align(seg * 4);
    b += coop_align((size_t)n_slots * lanes4 * 16) + coop_align((size_t)M * 2) + coop_align((size_t)n_slots * 2) + coop_align(M) + 16;
  }
  return b;
}

__device__ inline CoopSmem coop_carve(unsigned char* base, int round_rows, int H, int M, int lanes4, bool backward, int n_slots) {
  CoopSmem s{};
  const size_t seg = (size_t)round_rows * H * M;
  unsigned char* p = base;
  s.seg_p = reinterpret_cast<float*>(p);
  p += coop_align(seg * 4);
  if (backward) {
    s.seg_pd = reinterpret_cast<float*>(p);
    p += coop_align(seg * 4);
    s.slot_acc = reinterpret_cast<float4*>(p);
    p += coop_align((size_t)n_slots * lanes4 * 16);
  }
  s.cnt = reinterpret_cast<int*>(p);
  p += coop_align((size_t)round_rows * H * 4);
  s.rowaux = reinterpret_cast<float*>(p);
  p += coop_align((size_t)round_rows * H * 4);
  s.seg_j = reinterpret_cast<uint16_t*>(p);
  p += coop_align(seg * 2);
  if (backward) {
    s.map = reinterpret_cast<int16_t*>(p);
    p += coop_align((size_t)M * 2);
    s.slot_j = reinterpret_cast<int16_t*>(p);
    p += coop_align((size_t)n_slots * 2);
    s.touched = reinterpret_cast<uint8_t*>(p);
    p += coop_align(M);
    s.ctl = reinterpret_cast<int*>(p);
  }
  return s;
}

// One head of one row against the register-resident columns; appends kept entries to the segment.
// `post` multiplies the stored weight (1 in the forward pass, 1/l in the backward pass).
template <int CPL, bool BACKWARD>
__device__ __forceinline__ int coop_scan_head(const float (&d2)[CPL], int M, int lane, float s, float top, float cut, float post,
                                              uint16_t* seg_j, float* seg_p, float* seg_pd, uint8_t* touched, float& psum,
                                              float& pdsum) {
  int n = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    float p = 0.f;
    if (j < M) {
      const float sc = __fmul_rn(d2[c], s);
      if (sc <= cut) p = expf(__fsub_rn(top, sc));
    }
    const bool keep = p > 0.f;
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) {
      const int pos = n + __popc(m & lt);
      const float pw = p * post;
      seg_j[pos] = (uint16_t)j;
      seg_p[pos] = pw;
      if (BACKWARD) {
        seg_pd[pos] = pw * d2[c];
        if (touched) touched[j] = 1;
      }
    }
    n += __popc(m);
    psum += p;
    if (BACKWARD) pdsum += p * post * d2[c];
  }
  return n;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <int GEO, int CPL, int L4>
__global__ void __launch_bounds__(COOP_THREADS) coop_fwd_kernel(const CoopParams P) {
  extern __shared__ __align__(16) unsigned char coop_smem_raw[];
  const CoopSmem S = coop_carve(coop_smem_raw, P.round_rows, P.H, P.M, P.lanes4, false, 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float period = P.period ? __ldg(P.period) : 0.f;

  Point<GEO> col[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
  }
  int64_t val_off[L4], out_off[L4];
  bool ok[L4];
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    const int q = tid + k * COOP_THREADS;
    ok[k] = q < P.lanes4;
    const int e = q * 4;
    const int b = e / P.D, d = e - b * P.D;
    val_off[k] = (int64_t)b * P.M * P.D + d;
    out_off[k] = (int64_t)b * P.N * P.ld_out + P.col_off + d;
  }

  const int row_begin = blockIdx.x * P.rows_per_cta;
  const int row_end = min(P.N, row_begin + P.rows_per_cta);
  for (int r0 = row_begin; r0 < row_end; r0 += P.round_rows) {
    const int in_round = min(P.round_rows, row_end - r0);
    // ---- phase 1 ----
    for (int w = warp; w < in_round; w += COOP_WARPS) {
      const int r = r0 + w;
      const Point<GEO> o = load_point<GEO>(P.mesh_out, r, P.sd);
      float d2[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) d2[c] = dist2<GEO>(o, col[c], period);
      const float vmin = __ldg(P.v_min + r);
      const float vlo = P.masked ? __ldg(P.v_lo + r) : 0.f, vhi = P.masked ? __ldg(P.v_hi + r) : 0.f;
      for (int h = 0; h < P.H; ++h) {
        const float s = __ldg(P.scale + h);
        const float top = __fmul_rn(vmin, s);
        const float cut = P.masked ? head_threshold(vlo, vhi, s, P.weight) : INFINITY;
        const int seg = (w * P.H + h) * P.M;
        float psum = 0.f, unused = 0.f;
        const int n = coop_scan_head<CPL, false>(d2, P.M, lane, s, top, cut, 1.f, S.seg_j + seg, S.seg_p + seg, nullptr, nullptr,
                                                 psum, unused);
        psum = warp_sum(psum);
        if (lane == 0) {
          S.cnt[w * P.H + h] = n;
          S.rowaux[w * P.H + h] = 1.f / psum;
          P.rowsum[(int64_t)h * P.N + r] = psum;
        }
      }
    }
    __syncthreads();
    // ---- phase 2 ----
    for (int w = 0; w < in_round; ++w) {
      const int r = r0 + w;
      for (int h = 0; h < P.H; ++h) {
        const int n = S.cnt[w * P.H + h];
        const float inv_l = S.rowaux[w * P.H + h];
        const uint16_t* sj = S.seg_j + (w * P.H + h) * P.M;
        const float* sp = S.seg_p + (w * P.H + h) * P.M;
        float4 acc[L4];
#pragma unroll
        for (int k = 0; k < L4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e0 = 0; e0 < n; e0 += COOP_GATHER) {
          // issue a whole batch of independent 128-bit gathers before the first FMA consumes one
          float4 u[COOP_GATHER][L4];
          float pw[COOP_GATHER];
#pragma unroll
          for (int t = 0; t < COOP_GATHER; ++t) {
            const bool live = e0 + t < n;
            const int64_t joff = live ? (int64_t)sj[e0 + t] * P.D : 0;
            pw[t] = live ? sp[e0 + t] : 0.f;
#pragma unroll
            for (int k = 0; k < L4; ++k)
              u[t][k] = (live && ok[k]) ? __ldg(reinterpret_cast<const float4*>(P.values + val_off[k] + joff))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int t = 0; t < COOP_GATHER; ++t) {
#pragma unroll
            for (int k = 0; k < L4; ++k) {
              acc[k].x = fmaf(pw[t], u[t][k].x, acc[k].x);
              acc[k].y = fmaf(pw[t], u[t][k].y, acc[k].y);
              acc[k].z = fmaf(pw[t], u[t][k].z, acc[k].z);
              acc[k].w = fmaf(pw[t], u[t][k].w, acc[k].w);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < L4; ++k) {
          if (ok[k]) {
            const float4 o4 = make_float4(acc[k].x * inv_l, acc[k].y * inv_l, acc[k].z * inv_l, acc[k].w * inv_l);
            *reinterpret_cast<float4*>(P.out + out_off[k] + (int64_t)r * P.ld_out + (int64_t)h * P.D) = o4;
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// backward (scale gradient and, optionally, value gradient) in one pass over d_out
// ---------------------------------------------------------------------------------------
template <int L4>
__device__ __forceinline__ void coop_flush_slots(const CoopParams& P, const CoopSmem& S, const int64_t (&dv_off)[L4],
                                                 const bool (&ok)[L4], int tid) {
  const int used = min(S.ctl[0], P.n_slots);
  for (int sidx = 0; sidx < used; ++sidx) {
    const int64_t joff = (int64_t)S.slot_j[sidx] * P.D;
#pragma unroll
    for (int k = 0; k < L4; ++k) {
      if (ok[k]) {
        float4* cell = S.slot_acc + (size_t)sidx * P.lanes4 + tid + k * COOP_THREADS;
        atomicAdd(reinterpret_cast<float4*>(P.d_values + dv_off[k] + joff), *cell);
        *cell = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

template <int GEO, int CPL, int L4, bool WITH_VALUES>
__global__ void __launch_bounds__(COOP_THREADS) coop_bwd_kernel(const CoopParams P) {
  extern __shared__ __align__(16) unsigned char coop_smem_raw[];
  const CoopSmem S = coop_carve(coop_smem_raw, P.round_rows, P.H, P.M, P.lanes4, true, P.n_slots);
  __shared__ float red[COOP_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const bool want_scale = P.d_scale != nullptr;

  Point<GEO> col[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
  }
  int64_t val_off[L4], g_off[L4];
  bool ok[L4];
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    const int q = tid + k * COOP_THREADS;
    ok[k] = q < P.lanes4;
    const int e = q * 4;
    const int b = e / P.D, d = e - b * P.D;
    val_off[k] = (int64_t)b * P.M * P.D + d;  // same offset addresses values and d_values
    g_off[k] = (int64_t)b * P.N * P.ld_out + P.col_off + d;
  }
  if (WITH_VALUES) {
    for (int j = tid; j < P.M; j += COOP_THREADS) {
      S.map[j] = -1;
      S.touched[j] = 0;
    }
    for (int i = tid; i < P.n_slots * P.lanes4; i += COOP_THREADS) S.slot_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
      S.ctl[0] = 0;
      S.ctl[1] = 0;
    }
  }
  __syncthreads();

  float ds_part = 0.f;  // this thread's share of sum_e dO_e (W_e - m O_e), all heads interleaved below
  // one partial per head would need H registers; heads are few, so keep a small fixed array
  float ds_head[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) ds_head[h] = 0.f;
  (void)ds_part;

  const int row_begin = blockIdx.x * P.rows_per_cta;
  const int row_end = min(P.N, row_begin + P.rows_per_cta);
  for (int r0 = row_begin; r0 < row_end; r0 += P.round_rows) {
    const int in_round = min(P.round_rows, row_end - r0);
    // Pull the next round's d_out rows into L2 while this round is being processed (each thread its own lane).
    {
      const int nr0 = r0 + P.round_rows;
      const int n_next = min(P.round_rows, row_end - nr0);
      for (int w = 0; w < n_next; ++w) {
        for (int h = 0; h < P.H; ++h) {
#pragma unroll
          for (int k = 0; k < L4; ++k) {
            if (ok[k]) {
              const float* q = P.d_out + g_off[k] + (int64_t)(nr0 + w) * P.ld_out + (int64_t)h * P.D;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
          }
        }
      }
    }
    // ---- phase 1: normalised weights (l is known from the forward pass) ----
    for (int w = warp; w < in_round; w += COOP_WARPS) {
      const int r = r0 + w;
      const Point<GEO> o = load_point<GEO>(P.mesh_out, r, P.sd);
      float d2[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) d2[c] = dist2<GEO>(o, col[c], period);
      const float vmin = __ldg(P.v_min + r);
      const float vlo = P.masked ? __ldg(P.v_lo + r) : 0.f, vhi = P.masked ? __ldg(P.v_hi + r) : 0.f;
      for (int h = 0; h < P.H; ++h) {
        const float s = __ldg(P.scale + h);
        const float top = __fmul_rn(vmin, s);
        const float cut = P.masked ? head_threshold(vlo, vhi, s, P.weight) : INFINITY;
        const float inv_l = 1.f / __ldg(P.rowsum + (int64_t)h * P.N + r);
        const int seg = (w * P.H + h) * P.M;
        float psum = 0.f, pdsum = 0.f;
        const int n = coop_scan_head<CPL, true>(d2, P.M, lane, s, top, cut, inv_l, S.seg_j + seg, S.seg_p + seg, S.seg_pd + seg,
                                                WITH_VALUES ? S.touched : nullptr, psum, pdsum);
        pdsum = warp_sum(pdsum);
        if (lane == 0) {
          S.cnt[w * P.H + h] = n;
          S.rowaux[w * P.H + h] = pdsum;  // m = sum_j P^ d2
        }
      }
    }
    __syncthreads();
    if (WITH_VALUES) {
      // bind a slot to every column touched in this round; flush everything once if the set is full
      for (int attempt = 0; attempt < 2; ++attempt) {
        for (int j = tid; j < P.M; j += COOP_THREADS) {
          if (S.touched[j] && S.map[j] < 0) {
            const int sidx = atomicAdd(&S.ctl[0], 1);
            if (sidx < P.n_slots) {
              S.map[j] = (int16_t)sidx;
              S.slot_j[sidx] = (int16_t)j;
            } else {
              S.ctl[1] = 1;
            }
          }
        }
        __syncthreads();
        const bool overflow = S.ctl[1] != 0;
        if (!overflow) break;
        if (attempt == 0) {
          coop_flush_slots<L4>(P, S, val_off, ok, tid);
          __syncthreads();
          for (int j = tid; j < P.M; j += COOP_THREADS) S.map[j] = -1;
          if (tid == 0) {
            S.ctl[0] = 0;
            S.ctl[1] = 0;
          }
          __syncthreads();
        } else {
          // more distinct columns in one round than slots: the unbound ones go straight to global REDs
          if (tid == 0) {
            S.ctl[0] = P.n_slots;
            S.ctl[1] = 0;
          }
        }
      }
      for (int j = tid; j < P.M; j += COOP_THREADS) S.touched[j] = 0;
      __syncthreads();
    }
    // ---- phase 2 ----
    for (int w = 0; w < in_round; ++w) {
      const int r = r0 + w;
      for (int h = 0; h < P.H; ++h) {
        const int n = S.cnt[w * P.H + h];
        const float m = S.rowaux[w * P.H + h];
        const uint16_t* sj = S.seg_j + (w * P.H + h) * P.M;
        const float* sp = S.seg_p + (w * P.H + h) * P.M;
        const float* spd = S.seg_pd + (w * P.H + h) * P.M;
        float4 g[L4], acc_o[L4], acc_w[L4];
#pragma unroll
        for (int k = 0; k < L4; ++k) {
          g[k] = ok[k] ? __ldg(reinterpret_cast<const float4*>(P.d_out + g_off[k] + (int64_t)r * P.ld_out + (int64_t)h * P.D))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
          acc_o[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          acc_w[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int e0 = 0; e0 < n; e0 += COOP_GATHER_BWD) {
          float4 u[COOP_GATHER_BWD][L4];
          float pw[COOP_GATHER_BWD], pdw[COOP_GATHER_BWD];
          int jj[COOP_GATHER_BWD];
#pragma unroll
          for (int t = 0; t < COOP_GATHER_BWD; ++t) {
            const bool live = e0 + t < n;
            jj[t] = live ? (int)sj[e0 + t] : -1;
            pw[t] = live ? sp[e0 + t] : 0.f;
            pdw[t] = live ? spd[e0 + t] : 0.f;
            if (want_scale) {
#pragma unroll
              for (int k = 0; k < L4; ++k)
                u[t][k] = (live && ok[k]) ? __ldg(reinterpret_cast<const float4*>(P.values + val_off[k] + (int64_t)jj[t] * P.D))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (want_scale) {
#pragma unroll
            for (int t = 0; t < COOP_GATHER_BWD; ++t) {
#pragma unroll
              for (int k = 0; k < L4; ++k) {
                acc_o[k].x = fmaf(pw[t], u[t][k].x, acc_o[k].x);
                acc_o[k].y = fmaf(pw[t], u[t][k].y, acc_o[k].y);
                acc_o[k].z = fmaf(pw[t], u[t][k].z, acc_o[k].z);
                acc_o[k].w = fmaf(pw[t], u[t][k].w, acc_o[k].w);
                acc_w[k].x = fmaf(pdw[t], u[t][k].x, acc_w[k].x);
                acc_w[k].y = fmaf(pdw[t], u[t][k].y, acc_w[k].y);
                acc_w[k].z = fmaf(pdw[t], u[t][k].z, acc_w[k].z);
                acc_w[k].w = fmaf(pdw[t], u[t][k].w, acc_w[k].w);
              }
            }
          }
          if (WITH_VALUES) {
#pragma unroll
            for (int t = 0; t < COOP_GATHER_BWD; ++t) {
              if (jj[t] < 0) continue;
              const int sidx = S.map[jj[t]];
#pragma unroll
              for (int k = 0; k < L4; ++k) {
                if (!ok[k]) continue;
                const float4 add = make_float4(pw[t] * g[k].x, pw[t] * g[k].y, pw[t] * g[k].z, pw[t] * g[k].w);
                if (sidx >= 0) {
                  float4* cell = S.slot_acc + (size_t)sidx * P.lanes4 + tid + k * COOP_THREADS;
                  float4 cur = *cell;
                  cur.x += add.x;
                  cur.y += add.y;
                  cur.z += add.z;
                  cur.w += add.w;
                  *cell = cur;
                } else {
                  atomicAdd(reinterpret_cast<float4*>(P.d_values + val_off[k] + (int64_t)jj[t] * P.D), add);
                }
              }
            }
          }
        }
        if (want_scale) {
          float dot = 0.f;
#pragma unroll
          for (int k = 0; k < L4; ++k) {
            dot = fmaf(g[k].x, acc_w[k].x - m * acc_o[k].x, dot);
            dot = fmaf(g[k].y, acc_w[k].y - m * acc_o[k].y, dot);
            dot = fmaf(g[k].z, acc_w[k].z - m * acc_o[k].z, dot);
            dot = fmaf(g[k].w, acc_w[k].w - m * acc_o[k].w, dot);
          }
#pragma unroll
          for (int hh = 0; hh < 8; ++hh)
            if (hh == h) ds_head[hh] += dot;
        }
      }
    }
    __syncthreads();
  }
  if (WITH_VALUES) {
    coop_flush_slots<L4>(P, S, val_off, ok, tid);
  }
  if (want_scale) {
    for (int h = 0; h < P.H && h < 8; ++h) {
      float v = warp_sum(ds_head[h]);
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < COOP_WARPS; ++w) t += red[w];
        atomicAdd(P.d_scale + h, -t);
      }
      __syncthreads();
    }
  }
}

}  // namespace pit
