// K1t: position-attention for shared meshes whose column set is small (M <= 1024, H <= 2):
// the decoder stages (N >> M), the latent self stages and small encoders.
//
// Common first step ("scan"): the M column coordinates live in registers for the kernel's whole life
// (CPL = ceil(M/32) points per lane).  A warp evaluates one output row against all of them -- bit-exact
// d2, per-head scale, quantile cut, exp with the known shift -- and ballot-compacts the columns kept by
// ANY head into one 16-byte entry {column, d2, weight of head 0, weight of head 1} in shared memory.
// No global load sits inside the sweep, and both heads share one gather of the value row later on.
//
//   tall_fwd_kernel   warp-independent: the warp that scanned a row also contracts it.  Each lane owns
//                     L4 float4 lanes of the B*D-wide value vector; entries are consumed in batches so
//                     that several 128-bit gathers are in flight before the first FMA; rows leave as
//                     full 128-bit stores.  No block-level synchronisation.
//   tall_bwd_kernel   CTA-cooperative, one pass over dO (read once, 128-bit) for BOTH gradients:
//                     d scale  thread-local partial of -sum_e dO_e (W_e - m O_e), W = sum_j P^ d2 U,
//                              O = sum_j P^ U, m = sum_j P^ d2; one block reduction at the very end;
//                     d values dU[b,j,:] += sum_h P^_hij dO[b,i,h,:] accumulated WITHOUT atomics in a
//                              small set of shared-memory slots (a thread owns its lane of every slot);
//                              a slot is bound to a column on first touch, and all slots are flushed
//                              with vector REDs (RED.E.ADD.F32x4) when the set is full or the CTA ends.
//                              Rows are visited in mesh order, so the touched set stays tiny on
//                              spatially coherent meshes.
#pragma once
#include "geometry.cuh"

namespace pit {

constexpr int TALL_THREADS = 128;
constexpr int TALL_WARPS = TALL_THREADS / 32;
constexpr int TALL_MAX_M = 1024;
constexpr int TALL_MAX_H = 2;

struct TallParams {
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
  int lanes4;        // B*D/4 float4 lanes in a row's value vector
  int rows_per_unit;  // contiguous rows per warp (forward) or per CTA (backward)
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

// Scan one row.  `post[h]` multiplies the stored weight (1 forward, 1/l backward).  Returns the entry count.
// `vcap` is a head-independent pre-filter in d2 space: a kept column satisfies fl(d2*s) <= T <= fl(v_hi*s), and
// rounding is monotone, so d2 > v_hi*(1+2^-20) can never be kept; whole 32-column groups without a candidate
// skip the per-head work.  The exact per-head test still decides.
template <int GEO, int CPL, int NH, bool BACKWARD>
__device__ __forceinline__ int tall_scan_row(const Point<GEO>& o, const Point<GEO> (&col)[CPL], int M, int lane, float period,
                                             float vcap, const float (&s)[NH], const float (&top)[NH], const float (&cut)[NH],
                                             const float (&post)[NH], float4* seg, uint8_t* touched, float (&psum)[NH],
                                             float (&pdsum)[NH]) {
  int n = 0;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    const float d2 = dist2<GEO>(o, col[c], period);
    const bool cand = (j < M) && (d2 <= vcap);
    if (!__any_sync(FULL, cand)) continue;
    float p[NH];
    bool any = false;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      p[h] = 0.f;
      if (cand) {
        const float sc = __fmul_rn(d2, s[h]);
        if (sc <= cut[h]) p[h] = __expf(__fsub_rn(top[h], sc));
      }
      psum[h] += p[h];
      if (BACKWARD) {
        p[h] *= post[h];
        pdsum[h] = fmaf(p[h], d2, pdsum[h]);
      }
      any = any || (p[h] > 0.f);
    }
    const unsigned m = __ballot_sync(FULL, any);
    if (any) {
      const int pos = n + __popc(m & lt);
      seg[pos] = make_float4(__int_as_float(j), d2, p[0], NH > 1 ? p[NH - 1] : 0.f);
      if (BACKWARD && touched) touched[j] = 1;
    }
    n += __popc(m);
  }
  return n;
}

// ---------------------------------------------------------------------------------------
// forward: one warp per row, lanes own L4 float4 lanes each (chunk blockIdx.y of 32*L4 lanes)
// ---------------------------------------------------------------------------------------
template <int GEO, int CPL, int NH, int L4>
__global__ void __launch_bounds__(TALL_THREADS) tall_fwd_kernel(const TallParams P) {
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  constexpr int G = 8 / L4;  // entries gathered per batch: G * L4 = 8 independent 128-bit loads in flight per lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* seg = reinterpret_cast<float4*>(tall_smem_raw) + (size_t)warp * (CPL * 32);
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

  int64_t val_off[L4], out_off[L4];
  bool ok[L4];
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    const int q = (blockIdx.y * L4 + k) * 32 + lane;
    ok[k] = q < P.lanes4;
    const int e = q * 4;
    const int b = e / P.D, d = e - b * P.D;
    val_off[k] = (int64_t)b * P.M * P.D + d;
    out_off[k] = (int64_t)b * P.N * P.ld_out + P.col_off + d;
  }

  const int64_t gw = (int64_t)blockIdx.x * TALL_WARPS + warp;
  const int64_t row_begin = gw * P.rows_per_unit;
  const int row_end = (int)min((int64_t)P.N, row_begin + P.rows_per_unit);
  for (int r = (int)row_begin; r < row_end; ++r) {
    const Point<GEO> o = load_point<GEO>(P.mesh_out, r, P.sd);
    const float vmin = __ldg(P.v_min + r);
    const float vlo = P.masked ? __ldg(P.v_lo + r) : 0.f, vhi = P.masked ? __ldg(P.v_hi + r) : 0.f;
    float top[NH], cut[NH], post[NH], psum[NH], pdsum[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      top[h] = __fmul_rn(vmin, s[h]);
      cut[h] = P.masked ? head_threshold(vlo, vhi, s[h], P.weight) : INFINITY;
      post[h] = 1.f;
      psum[h] = 0.f;
      pdsum[h] = 0.f;
    }
    const float vcap = P.masked ? vhi * 1.000001f : INFINITY;
    const int n = tall_scan_row<GEO, CPL, NH, false>(o, col, P.M, lane, period, vcap, s, top, cut, post, seg, nullptr, psum, pdsum);
    __syncwarp();
    float inv_l[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      psum[h] = warp_sum(psum[h]);
      inv_l[h] = 1.f / psum[h];
      if (lane == 0 && blockIdx.y == 0) P.rowsum[(int64_t)h * P.N + r] = psum[h];
    }

    float4 acc[NH][L4];
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int k = 0; k < L4; ++k) acc[h][k] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int e0 = 0; e0 < n; e0 += G) {
      float4 ent[G];
      float4 u[G][L4];
#pragma unroll
      for (int t = 0; t < G; ++t) {
        const bool live = e0 + t < n;
        ent[t] = live ? seg[e0 + t] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t joff = (int64_t)__float_as_int(ent[t].x) * P.D;
#pragma unroll
        for (int k = 0; k < L4; ++k)
          u[t][k] = (live && ok[k]) ? __ldg(reinterpret_cast<const float4*>(P.values + val_off[k] + joff))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int t = 0; t < G; ++t) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float pw = h == 0 ? ent[t].z : ent[t].w;
#pragma unroll
          for (int k = 0; k < L4; ++k) {
            acc[h][k].x = fmaf(pw, u[t][k].x, acc[h][k].x);
            acc[h][k].y = fmaf(pw, u[t][k].y, acc[h][k].y);
            acc[h][k].z = fmaf(pw, u[t][k].z, acc[h][k].z);
            acc[h][k].w = fmaf(pw, u[t][k].w, acc[h][k].w);
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
#pragma unroll
      for (int k = 0; k < L4; ++k) {
        if (ok[k]) {
          const float4 o4 = make_float4(acc[h][k].x * inv_l[h], acc[h][k].y * inv_l[h], acc[h][k].z * inv_l[h], acc[h][k].w * inv_l[h]);
          *reinterpret_cast<float4*>(P.out + out_off[k] + (int64_t)r * P.ld_out + (int64_t)h * P.D) = o4;
        }
      }
    }
    __syncwarp();  // the segment is rewritten by the next row
  }
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
struct TallBwdSmem {
  float4* seg;       // [TALL_WARPS][M_pad] entries
  float4* slot_acc;  // [n_slots][lanes4]
  float* rowm;       // [TALL_WARPS][2]   m = sum_j P^ d2 per head
  int* cnt;          // [TALL_WARPS]
  int16_t* map;      // [M] column -> slot or -1
  int16_t* slot_j;   // [n_slots]
  uint8_t* touched;  // [M]
  int* ctl;          // [0] slots in use, [1] overflow flag
};

__host__ __device__ inline size_t tall_align(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t tall_bwd_smem_bytes(int cpl, int M, int lanes4, int n_slots) {
  return (size_t)TALL_WARPS * cpl * 32 * 16 + tall_align((size_t)n_slots * lanes4 * 16) + tall_align(TALL_WARPS * 2 * 4) +
         tall_align(TALL_WARPS * 4) + tall_align((size_t)M * 2) + tall_align((size_t)n_slots * 2 + 2) + tall_align(M) + 16;
}

__device__ inline TallBwdSmem tall_bwd_carve(unsigned char* p, int cpl, int M, int lanes4, int n_slots) {
  TallBwdSmem s{};
  s.seg = reinterpret_cast<float4*>(p);
  p += (size_t)TALL_WARPS * cpl * 32 * 16;
  s.slot_acc = reinterpret_cast<float4*>(p);
  p += tall_align((size_t)n_slots * lanes4 * 16);
  s.rowm = reinterpret_cast<float*>(p);
  p += tall_align(TALL_WARPS * 2 * 4);
  s.cnt = reinterpret_cast<int*>(p);
  p += tall_align(TALL_WARPS * 4);
  s.map = reinterpret_cast<int16_t*>(p);
  p += tall_align((size_t)M * 2);
  s.slot_j = reinterpret_cast<int16_t*>(p);
  p += tall_align((size_t)n_slots * 2 + 2);
  s.touched = reinterpret_cast<uint8_t*>(p);
  p += tall_align(M);
  s.ctl = reinterpret_cast<int*>(p);
  return s;
}

template <int L4>
__device__ __forceinline__ void tall_flush_slots(const TallParams& P, const TallBwdSmem& S, const int64_t (&val_off)[L4],
                                                 const bool (&ok)[L4], int tid) {
  const int used = min(S.ctl[0], P.n_slots);
  for (int sidx = 0; sidx < used; ++sidx) {
    const int64_t joff = (int64_t)S.slot_j[sidx] * P.D;
#pragma unroll
    for (int k = 0; k < L4; ++k) {
      if (ok[k]) {
        float4* cell = S.slot_acc + (size_t)sidx * P.lanes4 + tid + k * TALL_THREADS;
        atomicAdd(reinterpret_cast<float4*>(P.d_values + val_off[k] + joff), *cell);
        *cell = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// L4 here: float4 lanes per THREAD (lanes4 <= L4 * 128).
template <int GEO, int CPL, int NH, int L4, bool WITH_VALUES>
__global__ void __launch_bounds__(TALL_THREADS) tall_bwd_kernel(const TallParams P) {
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  const TallBwdSmem S = tall_bwd_carve(tall_smem_raw, CPL, P.M, P.lanes4, P.n_slots);
  __shared__ float red[TALL_WARPS];
  constexpr int G = (L4 >= 4) ? 1 : (4 / L4);  // entries per gather batch
  constexpr int SEG = CPL * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float period = P.period ? __ldg(P.period) : 0.f;
  const bool want_scale = P.d_scale != nullptr;

  Point<GEO> col[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int j = c * 32 + lane;
    col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
  }
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);

  int64_t val_off[L4], g_off[L4];
  bool ok[L4];
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    const int q = tid + k * TALL_THREADS;
    ok[k] = q < P.lanes4;
    const int e = q * 4;
    const int b = e / P.D, d = e - b * P.D;
    val_off[k] = (int64_t)b * P.M * P.D + d;  // addresses both values and d_values
    g_off[k] = (int64_t)b * P.N * P.ld_out + P.col_off + d;
  }
  if (WITH_VALUES) {
    for (int j = tid; j < P.M; j += TALL_THREADS) {
      S.map[j] = -1;
      S.touched[j] = 0;
    }
    for (int i = tid; i < P.n_slots * P.lanes4; i += TALL_THREADS) S.slot_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
      S.ctl[0] = 0;
      S.ctl[1] = 0;
    }
  }
  __syncthreads();

  float ds_head[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) ds_head[h] = 0.f;

  const int row_begin = blockIdx.x * P.rows_per_unit;
  const int row_end = min(P.N, row_begin + P.rows_per_unit);
  for (int r0 = row_begin; r0 < row_end; r0 += TALL_WARPS) {
    const int in_round = min(TALL_WARPS, row_end - r0);
    // Pull the next round's d_out rows into L2 while this round is processed (each thread its own lanes).
    {
      const int nr0 = r0 + TALL_WARPS;
      const int n_next = min(TALL_WARPS, row_end - nr0);
      for (int w = 0; w < n_next; ++w) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
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
    // ---- phase 1: one row per warp, normalised weights (l is known from the forward pass) ----
    if (warp < in_round) {
      const int r = r0 + warp;
      const Point<GEO> o = load_point<GEO>(P.mesh_out, r, P.sd);
      const float vmin = __ldg(P.v_min + r);
      const float vlo = P.masked ? __ldg(P.v_lo + r) : 0.f, vhi = P.masked ? __ldg(P.v_hi + r) : 0.f;
      float top[NH], cut[NH], post[NH], psum[NH], pdsum[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        top[h] = __fmul_rn(vmin, s[h]);
        cut[h] = P.masked ? head_threshold(vlo, vhi, s[h], P.weight) : INFINITY;
        post[h] = 1.f / __ldg(P.rowsum + (int64_t)h * P.N + r);
        psum[h] = 0.f;
        pdsum[h] = 0.f;
      }
      const float vcap = P.masked ? vhi * 1.000001f : INFINITY;
      const int n = tall_scan_row<GEO, CPL, NH, true>(o, col, P.M, lane, period, vcap, s, top, cut, post, S.seg + (size_t)warp * SEG,
                                                      WITH_VALUES ? S.touched : nullptr, psum, pdsum);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float m = warp_sum(pdsum[h]);
        if (lane == 0) S.rowm[warp * 2 + h] = m;
      }
      if (lane == 0) S.cnt[warp] = n;
    }
    __syncthreads();
    if (WITH_VALUES) {
      // bind a slot to every column touched in this round; flush everything once if the set is full
      for (int attempt = 0; attempt < 2; ++attempt) {
        for (int j = tid; j < P.M; j += TALL_THREADS) {
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
        __syncthreads();  // every thread has read the flag before thread 0 may reset it
        if (!overflow) break;
        if (attempt == 0) {
          tall_flush_slots<L4>(P, S, val_off, ok, tid);
          __syncthreads();
          for (int j = tid; j < P.M; j += TALL_THREADS) S.map[j] = -1;
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
      for (int j = tid; j < P.M; j += TALL_THREADS) S.touched[j] = 0;
      __syncthreads();
    }
    // ---- phase 2: every thread owns its float4 lanes; rows of the round one after the other ----
    for (int w = 0; w < in_round; ++w) {
      const int r = r0 + w;
      const int n = S.cnt[w];
      const float4* seg = S.seg + (size_t)w * SEG;
      // d scale: -sum_e dO_e (W_e - m O_e) = -sum_e dO_e Z_e with Z = sum_j P^_j (d2_j - m) U_j  (one accumulator per head)
      float4 g[NH][L4], acc_z[NH][L4];
      float mrow[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        mrow[h] = S.rowm[w * 2 + h];
#pragma unroll
        for (int k = 0; k < L4; ++k) {
          g[h][k] = ok[k] ? __ldg(reinterpret_cast<const float4*>(P.d_out + g_off[k] + (int64_t)r * P.ld_out + (int64_t)h * P.D))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
          acc_z[h][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      for (int e0 = 0; e0 < n; e0 += G) {
        float4 ent[G];
        float4 u[G][L4];
#pragma unroll
        for (int t = 0; t < G; ++t) {
          const bool live = e0 + t < n;
          ent[t] = live ? seg[e0 + t] : make_float4(__int_as_float(-1), 0.f, 0.f, 0.f);
          if (want_scale) {
            const int64_t joff = (int64_t)max(__float_as_int(ent[t].x), 0) * P.D;
#pragma unroll
            for (int k = 0; k < L4; ++k)
              u[t][k] = (live && ok[k]) ? __ldg(reinterpret_cast<const float4*>(P.values + val_off[k] + joff))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int t = 0; t < G; ++t) {
          const int j = __float_as_int(ent[t].x);
          const float d2 = ent[t].y;
          if (want_scale) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const float pz = (h == 0 ? ent[t].z : ent[t].w) * (d2 - mrow[h]);
#pragma unroll
              for (int k = 0; k < L4; ++k) {
                acc_z[h][k].x = fmaf(pz, u[t][k].x, acc_z[h][k].x);
                acc_z[h][k].y = fmaf(pz, u[t][k].y, acc_z[h][k].y);
                acc_z[h][k].z = fmaf(pz, u[t][k].z, acc_z[h][k].z);
                acc_z[h][k].w = fmaf(pz, u[t][k].w, acc_z[h][k].w);
              }
            }
          }
          if (WITH_VALUES && j >= 0) {
            const int sidx = S.map[j];
#pragma unroll
            for (int k = 0; k < L4; ++k) {
              if (!ok[k]) continue;
              float4 add = make_float4(ent[t].z * g[0][k].x, ent[t].z * g[0][k].y, ent[t].z * g[0][k].z, ent[t].z * g[0][k].w);
              if (NH > 1) {
                add.x = fmaf(ent[t].w, g[NH - 1][k].x, add.x);
                add.y = fmaf(ent[t].w, g[NH - 1][k].y, add.y);
                add.z = fmaf(ent[t].w, g[NH - 1][k].z, add.z);
                add.w = fmaf(ent[t].w, g[NH - 1][k].w, add.w);
              }
              if (sidx >= 0) {
                float4* cell = S.slot_acc + (size_t)sidx * P.lanes4 + tid + k * TALL_THREADS;
                float4 cur = *cell;
                cur.x += add.x;
                cur.y += add.y;
                cur.z += add.z;
                cur.w += add.w;
                *cell = cur;
              } else {
                atomicAdd(reinterpret_cast<float4*>(P.d_values + val_off[k] + (int64_t)j * P.D), add);
              }
            }
          }
        }
      }
      if (want_scale) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          float dot = 0.f;
#pragma unroll
          for (int k = 0; k < L4; ++k) {
            dot = fmaf(g[h][k].x, acc_z[h][k].x, dot);
            dot = fmaf(g[h][k].y, acc_z[h][k].y, dot);
            dot = fmaf(g[h][k].z, acc_z[h][k].z, dot);
            dot = fmaf(g[h][k].w, acc_z[h][k].w, dot);
          }
          ds_head[h] += dot;
        }
      }
    }
    __syncthreads();
  }
  if (WITH_VALUES) tall_flush_slots<L4>(P, S, val_off, ok, tid);
  if (want_scale) {
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float v = warp_sum(ds_head[h]);
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < TALL_WARPS; ++w) t += red[w];
        atomicAdd(P.d_scale + h, -t);
      }
      __syncthreads();
    }
  }
}

}  // namespace pit
