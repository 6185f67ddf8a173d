// K3: fused decoder tail -- pit.decoder (pit.py:124-127): cross position-attention `up` followed by the
// two-layer `de` MLP -- for shared meshes with a small latent set (M <= 1024, H <= 2).
//
// The reference writes the attention output (B x N x H*D fp32: 726 MB at Darcy-421), reads it back for
// Linear(H*D -> C), writes/reads the hidden activations for GELU and Linear(C -> O), and replays all of it
// in backward.  Here nothing N-sized except the final output (B x N x O) touches HBM:
//
//   * the first Linear is pushed through the attention, which is linear in its values:
//       W1 (concat_h sum_j A_h[n,j] U[b,j,:]) = sum_h sum_j A_h[n,j] Y[b,j,h,:],   Y[b,j,h,:] = W1[:, hD:(h+1)D] U[b,j,:]
//     Y lives on the latent mesh (B x M x H x C, 0.5 MB) and is produced by the caller with one small GEMM;
//   * the kernel gathers Y rows with the attention weights (same scan / entry list as the tall kernels), adds
//     b1, applies the exact (erf) GELU and the C -> O projection in registers, and stores O floats per point;
//   * backward recomputes the hidden pre-activation from the same gather, forms
//       g1 = gelu'(pre) * (W2^T dOut),  dY[b,j,h,:] += A_h[n,j] g1,  ds_h = -sum_c g1_c sum_j A_h (d2 - m_h) Y_c,
//     accumulates dY without atomics in shared-memory slots (as tall_bwd_kernel does for dU) and reduces
//     db1, dW2, db2 once per CTA.  dW1 and dU follow from dY through the caller's small GEMM.
#pragma once
#include "tall_attention.cuh"

namespace pit {

constexpr int TAIL_MAX_OUT = 4;

struct TailParams {
  const float* mesh_out;  // [N,sd]
  const float* mesh_in;   // [M,sd]
  const float* period;
  const float* y;      // [B,M,H,C]
  const float* scale;  // [H]
  const float* v_min;
  const float* v_lo;
  const float* v_hi;
  float weight;
  int masked;
  int B, H, N, M, C, O, sd;
  int lanes4;  // B*C/4
  int rows_per_unit;
  const float* b1;  // [C]
  const float* w2;  // [O,C]
  const float* b2;  // [O]
  // forward
  float* out;     // [B,N,O]
  float* rowsum;  // [H,N]
  // backward
  const float* d_out;  // [B,N,O]
  float* d_y;          // [B,M,H,C] zero-initialised
  float* d_scale;      // [H]       zero-initialised
  float* d_b1;         // [C]       zero-initialised
  float* d_w2;         // [O,C]     zero-initialised
  float* d_b2;         // [O]       zero-initialised
  int n_slots;
};

// Exact-GELU (erf form, torch's default) and its derivative from ONE exponential: with z = |x|/sqrt(2),
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),  t = 1/(1 + p z)      (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7)
// and exp(-z^2) = exp(-x^2/2) is also the Gaussian density needed by the derivative:
//   gelu(x) = x Phi(x),   gelu'(x) = Phi(x) + x exp(-x^2/2) / sqrt(2 pi),   Phi = (1 + erf(x/sqrt 2)) / 2.
// The 1.5e-7 absolute error is two orders below the 1e-5 parity budget; libm's erff costs ~3x the instructions.
__device__ __forceinline__ void gelu_pair(float x, float& g, float& dg) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  const float e = __expf(-z * z);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, e, 1.f);
  const float phi = 0.5f * (1.f + copysignf(erf_abs, x));
  g = x * phi;
  dg = fmaf(x * 0.3989422804014327f, e, phi);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float g, dg;
  gelu_pair(x, g, dg);
  return g;
}

// Sum over the `seg` consecutive lanes (seg a power of two <= 32) that share a sample.
__device__ __forceinline__ float seg_sum(float v, int seg) {
  for (int o = seg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// forward: one warp per row; a lane owns L4 float4 lanes (b, c..c+3) of the B*C-wide hidden vector
// ---------------------------------------------------------------------------------------
template <int GEO, int CPL, int NH, int L4>
__global__ void __launch_bounds__(TALL_THREADS, 4) tail_fwd_kernel(const TailParams P) {
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  constexpr int G = (NH * L4 >= 8) ? 2 : (8 / (NH * L4));  // entries per gather batch: >= 8 independent 128-bit loads in flight
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

  const int c4_per_b = P.C / 4;
  const int segw = c4_per_b < 32 ? c4_per_b : 32;
  // Lane -> (sample, channel) bookkeeping as 32-bit element offsets (the host guarantees B*M*H*C < 2^31); lanes past
  // the end are clamped onto the last valid lane so that every gather is unconditional, and only skip the final store.
  int y_off[L4];
  int bidx[L4], cidx[L4];
  bool ok[L4];
  // b1 and W2 live in shared memory behind the per-warp entry segments: [b1 (C) | W2 (O x C)]
  float* par = reinterpret_cast<float*>(tall_smem_raw + (size_t)TALL_WARPS * CPL * 32 * 16);
  for (int i = threadIdx.x; i < P.C * (1 + P.O); i += TALL_THREADS) par[i] = i < P.C ? __ldg(P.b1 + i) : __ldg(P.w2 + (i - P.C));
  __syncthreads();
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    int q = (blockIdx.y * L4 + k) * 32 + lane;
    ok[k] = q < P.lanes4;
    q = min(q, P.lanes4 - 1);
    bidx[k] = q / c4_per_b;
    cidx[k] = (q - bidx[k] * c4_per_b) * 4;
    y_off[k] = bidx[k] * P.M * NH * P.C + cidx[k];
    asm volatile("" : "+r"(y_off[k]), "+r"(cidx[k]), "+r"(bidx[k]));  // keep them in registers: no re-derivation inside the loops
  }
  const int row_stride = NH * P.C;
  // C/4 > 32: one sample spans k_per_b consecutive register groups, folded into the first ("lead") one
  const int k_per_b = c4_per_b > 32 ? c4_per_b / 32 : 1;
  bool k_lead[L4], writer[L4];
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    k_lead[k] = (k % k_per_b) == 0;
    writer[k] = ok[k] && k_lead[k] && (lane % segw) == 0;
  }
  float b2r[TAIL_MAX_OUT];
#pragma unroll
  for (int o = 0; o < TAIL_MAX_OUT; ++o) b2r[o] = o < P.O ? __ldg(P.b2 + o) : 0.f;

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

    // full batches of G entries (all gathers issued before the first FMA), then the remainder one by one
    int e0 = 0;
    for (; e0 + G <= n; e0 += G) {
      float4 ent[G];
      float4 u[G][NH][L4];
#pragma unroll
      for (int t = 0; t < G; ++t) {
        ent[t] = seg[e0 + t];
        const int jo = __float_as_int(ent[t].x) * row_stride;
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int k = 0; k < L4; ++k) u[t][h][k] = __ldg(reinterpret_cast<const float4*>(P.y + (y_off[k] + jo + h * P.C)));
      }
#pragma unroll
      for (int t = 0; t < G; ++t) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float pw = h == 0 ? ent[t].z : ent[t].w;
#pragma unroll
          for (int k = 0; k < L4; ++k) {
            acc[h][k].x = fmaf(pw, u[t][h][k].x, acc[h][k].x);
            acc[h][k].y = fmaf(pw, u[t][h][k].y, acc[h][k].y);
            acc[h][k].z = fmaf(pw, u[t][h][k].z, acc[h][k].z);
            acc[h][k].w = fmaf(pw, u[t][h][k].w, acc[h][k].w);
          }
        }
      }
    }
    for (; e0 < n; ++e0) {
      const float4 ent = seg[e0];
      const int jo = __float_as_int(ent.x) * row_stride;
      float4 u[NH][L4];
#pragma unroll
      for (int h = 0; h < NH; ++h)
#pragma unroll
        for (int k = 0; k < L4; ++k) u[h][k] = __ldg(reinterpret_cast<const float4*>(P.y + (y_off[k] + jo + h * P.C)));
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float pw = h == 0 ? ent.z : ent.w;
#pragma unroll
        for (int k = 0; k < L4; ++k) {
          acc[h][k].x = fmaf(pw, u[h][k].x, acc[h][k].x);
          acc[h][k].y = fmaf(pw, u[h][k].y, acc[h][k].y);
          acc[h][k].z = fmaf(pw, u[h][k].z, acc[h][k].z);
          acc[h][k].w = fmaf(pw, u[h][k].w, acc[h][k].w);
        }
      }
    }
    // hidden = gelu(b1 + sum_h acc_h / l_h); out[b, r, o] = b2[o] + <W2[o, :], hidden[b, :]>
    float part[TAIL_MAX_OUT][L4];
#pragma unroll
    for (int k = 0; k < L4; ++k) {
      float4 pre = *reinterpret_cast<const float4*>(par + cidx[k]);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        pre.x = fmaf(acc[h][k].x, inv_l[h], pre.x);
        pre.y = fmaf(acc[h][k].y, inv_l[h], pre.y);
        pre.z = fmaf(acc[h][k].z, inv_l[h], pre.z);
        pre.w = fmaf(acc[h][k].w, inv_l[h], pre.w);
      }
      const float4 hid = make_float4(gelu_erf(pre.x), gelu_erf(pre.y), gelu_erf(pre.z), gelu_erf(pre.w));
#pragma unroll
      for (int oo = 0; oo < TAIL_MAX_OUT; ++oo) {
        part[oo][k] = 0.f;
        if (oo < P.O) {
          const float4 wv = *reinterpret_cast<const float4*>(par + (1 + oo) * P.C + cidx[k]);
          part[oo][k] = hid.x * wv.x + hid.y * wv.y + hid.z * wv.z + hid.w * wv.w;
        }
      }
    }
    // reduce over the lanes (and, for C > 128, the consecutive k) that share a sample
#pragma unroll
    for (int oo = 0; oo < TAIL_MAX_OUT; ++oo) {
      if (oo >= P.O) break;
#pragma unroll
      for (int k = 0; k < L4; ++k) {
        if (!k_lead[k]) continue;  // uniform across the warp
        float v = part[oo][k];
#pragma unroll
        for (int kk = 1; kk < L4; ++kk)
          if (kk < k_per_b && k + kk < L4) v += part[oo][k + kk];
        v = seg_sum(v, segw);
        if (writer[k]) P.out[((int64_t)bidx[k] * P.N + r) * P.O + oo] = v + b2r[oo];
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------
// backward: CTA-cooperative (a thread owns float4 lanes (b, c..c+3)), one pass over the rows
// ---------------------------------------------------------------------------------------
struct TailBwdSmem {
  float4* seg;       // [TALL_WARPS][SEG]
  float4* slot_acc;  // [n_slots][NH][lanes4]
  float* rowm;       // [TALL_WARPS][2]
  int* cnt;          // [TALL_WARPS]
  int16_t* map;      // [M]
  int16_t* slot_j;   // [n_slots]
  uint8_t* touched;  // [M]
  int* ctl;
  float* red;        // [TALL_WARPS] scratch for the final reductions
  float* par;        // [C + O*C]: b1 then W2
};

__host__ __device__ inline size_t tail_bwd_smem_bytes(int cpl, int M, int lanes4, int nh, int n_slots, int C, int O) {
  return (size_t)TALL_WARPS * cpl * 32 * 16 + tall_align((size_t)n_slots * nh * lanes4 * 16) + tall_align(TALL_WARPS * 2 * 4) +
         tall_align(TALL_WARPS * 4) + tall_align((size_t)M * 2) + tall_align((size_t)n_slots * 2 + 2) + tall_align(M) + 16 + 64 +
         tall_align((size_t)C * (1 + O) * 4);
}

__device__ inline TailBwdSmem tail_bwd_carve(unsigned char* p, int cpl, int M, int lanes4, int nh, int n_slots) {
  TailBwdSmem s{};
  s.seg = reinterpret_cast<float4*>(p);
  p += (size_t)TALL_WARPS * cpl * 32 * 16;
  s.slot_acc = reinterpret_cast<float4*>(p);
  p += tall_align((size_t)n_slots * nh * lanes4 * 16);
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
  p += 16;
  s.red = reinterpret_cast<float*>(p);
  p += 64;
  s.par = reinterpret_cast<float*>(p);
  return s;
}

template <int NH, int L4>
__device__ __forceinline__ void tail_flush_slots(const TailParams& P, const TailBwdSmem& S, const int (&y_off)[L4],
                                                 const bool (&ok)[L4], int tid) {
  const int used = min(S.ctl[0], P.n_slots);
  for (int sidx = 0; sidx < used; ++sidx) {
    const int joff = (int)S.slot_j[sidx] * NH * P.C;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
#pragma unroll
      for (int k = 0; k < L4; ++k) {
        if (ok[k]) {
          float4* cell = S.slot_acc + ((size_t)sidx * NH + h) * P.lanes4 + tid + k * TALL_THREADS;
          atomicAdd(reinterpret_cast<float4*>(P.d_y + (y_off[k] + joff + h * P.C)), *cell);
          *cell = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
}

template <int GEO, int CPL, int NH, int L4>
__global__ void __launch_bounds__(TALL_THREADS, (L4 == 1) ? 4 : 1) tail_bwd_kernel(const TailParams P) {
  extern __shared__ __align__(16) unsigned char tall_smem_raw[];
  const TailBwdSmem S = tail_bwd_carve(tall_smem_raw, CPL, P.M, P.lanes4, NH, P.n_slots);
  constexpr int SEG = CPL * 32;
  constexpr int G = (NH * L4 >= 8) ? 1 : (8 / (NH * L4) > 4 ? 4 : 8 / (NH * L4));  // gather batch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

  const int c4_per_b = P.C / 4;
  int y_off[L4];
  int bidx[L4], cidx[L4];
  bool ok[L4];
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    int q = tid + k * TALL_THREADS;
    ok[k] = q < P.lanes4;
    q = min(q, P.lanes4 - 1);
    bidx[k] = q / c4_per_b;
    cidx[k] = (q - bidx[k] * c4_per_b) * 4;
    y_off[k] = bidx[k] * P.M * NH * P.C + cidx[k];
    asm volatile("" : "+r"(y_off[k]), "+r"(cidx[k]), "+r"(bidx[k]));
  }
  const int row_stride = NH * P.C;
  for (int i = tid; i < P.C * (1 + P.O); i += TALL_THREADS) S.par[i] = i < P.C ? __ldg(P.b1 + i) : __ldg(P.w2 + (i - P.C));
  for (int j = tid; j < P.M; j += TALL_THREADS) {
    S.map[j] = -1;
    S.touched[j] = 0;
  }
  for (int i = tid; i < P.n_slots * NH * P.lanes4; i += TALL_THREADS) S.slot_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    S.ctl[0] = 0;
    S.ctl[1] = 0;
  }
  __syncthreads();

  float ds_head[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) ds_head[h] = 0.f;
  float4 db1[L4], dw2[TAIL_MAX_OUT][L4];
  float db2[TAIL_MAX_OUT];
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    db1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int o = 0; o < TAIL_MAX_OUT; ++o) dw2[o][k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int o = 0; o < TAIL_MAX_OUT; ++o) db2[o] = 0.f;

  const int row_begin = blockIdx.x * P.rows_per_unit;
  const int row_end = min(P.N, row_begin + P.rows_per_unit);
  for (int r0 = row_begin; r0 < row_end; r0 += TALL_WARPS) {
    const int in_round = min(TALL_WARPS, row_end - r0);
    // ---- phase 1: one row per warp, normalised weights ----
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
                                                      S.touched, psum, pdsum);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float m = warp_sum(pdsum[h]);
        if (lane == 0) S.rowm[warp * 2 + h] = m;
      }
      if (lane == 0) S.cnt[warp] = n;
    }
    __syncthreads();
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
        tail_flush_slots<NH, L4>(P, S, y_off, ok, tid);
        __syncthreads();
        for (int j = tid; j < P.M; j += TALL_THREADS) S.map[j] = -1;
        if (tid == 0) {
          S.ctl[0] = 0;
          S.ctl[1] = 0;
        }
        __syncthreads();
      } else {
        if (tid == 0) {
          S.ctl[0] = P.n_slots;
          S.ctl[1] = 0;
        }
      }
    }
    for (int j = tid; j < P.M; j += TALL_THREADS) S.touched[j] = 0;
    __syncthreads();
    // ---- phase 2 ----
    for (int w = 0; w < in_round; ++w) {
      const int r = r0 + w;
      const int n = S.cnt[w];
      const float4* seg = S.seg + (size_t)w * SEG;
      float mrow[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) mrow[h] = S.rowm[w * 2 + h];
      // pass A: hidden pre-activation H = sum_h sum_j P^ Y  and  Z_h = sum_j P^ (d2 - m_h) Y
      float4 acc_h[L4], acc_z[NH][L4];
#pragma unroll
      for (int k = 0; k < L4; ++k) {
        acc_h[k] = *reinterpret_cast<const float4*>(S.par + cidx[k]);
#pragma unroll
        for (int h = 0; h < NH; ++h) acc_z[h][k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      int e0 = 0;
      for (; e0 + G <= n; e0 += G) {
        float4 ent[G];
        float4 u[G][NH][L4];
#pragma unroll
        for (int t = 0; t < G; ++t) {
          ent[t] = seg[e0 + t];
          const int jo = __float_as_int(ent[t].x) * row_stride;
#pragma unroll
          for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int k = 0; k < L4; ++k) u[t][h][k] = __ldg(reinterpret_cast<const float4*>(P.y + (y_off[k] + jo + h * P.C)));
        }
#pragma unroll
        for (int t = 0; t < G; ++t) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            const float pw = h == 0 ? ent[t].z : ent[t].w;
            const float pz = pw * (ent[t].y - mrow[h]);
#pragma unroll
            for (int k = 0; k < L4; ++k) {
              acc_h[k].x = fmaf(pw, u[t][h][k].x, acc_h[k].x);
              acc_h[k].y = fmaf(pw, u[t][h][k].y, acc_h[k].y);
              acc_h[k].z = fmaf(pw, u[t][h][k].z, acc_h[k].z);
              acc_h[k].w = fmaf(pw, u[t][h][k].w, acc_h[k].w);
              acc_z[h][k].x = fmaf(pz, u[t][h][k].x, acc_z[h][k].x);
              acc_z[h][k].y = fmaf(pz, u[t][h][k].y, acc_z[h][k].y);
              acc_z[h][k].z = fmaf(pz, u[t][h][k].z, acc_z[h][k].z);
              acc_z[h][k].w = fmaf(pz, u[t][h][k].w, acc_z[h][k].w);
            }
          }
        }
      }
      for (; e0 < n; ++e0) {
        const float4 ent = seg[e0];
        const int jo = __float_as_int(ent.x) * row_stride;
        float4 u[NH][L4];
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int k = 0; k < L4; ++k) u[h][k] = __ldg(reinterpret_cast<const float4*>(P.y + (y_off[k] + jo + h * P.C)));
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float pw = h == 0 ? ent.z : ent.w;
          const float pz = pw * (ent.y - mrow[h]);
#pragma unroll
          for (int k = 0; k < L4; ++k) {
            acc_h[k].x = fmaf(pw, u[h][k].x, acc_h[k].x);
            acc_h[k].y = fmaf(pw, u[h][k].y, acc_h[k].y);
            acc_h[k].z = fmaf(pw, u[h][k].z, acc_h[k].z);
            acc_h[k].w = fmaf(pw, u[h][k].w, acc_h[k].w);
            acc_z[h][k].x = fmaf(pz, u[h][k].x, acc_z[h][k].x);
            acc_z[h][k].y = fmaf(pz, u[h][k].y, acc_z[h][k].y);
            acc_z[h][k].z = fmaf(pz, u[h][k].z, acc_z[h][k].z);
            acc_z[h][k].w = fmaf(pz, u[h][k].w, acc_z[h][k].w);
          }
        }
      }
      // g1 = gelu'(pre) * (W2^T dOut[b, r, :]); parameter-gradient partials stay in registers
      float4 g1[L4];
#pragma unroll
      for (int k = 0; k < L4; ++k) {
        float go[TAIL_MAX_OUT];
#pragma unroll
        for (int o = 0; o < TAIL_MAX_OUT; ++o) go[o] = (ok[k] && o < P.O) ? __ldg(P.d_out + ((int64_t)bidx[k] * P.N + r) * P.O + o) : 0.f;
        const float4 pre = acc_h[k];
        float4 hid, dhid;
        gelu_pair(pre.x, hid.x, dhid.x);
        gelu_pair(pre.y, hid.y, dhid.y);
        gelu_pair(pre.z, hid.z, dhid.z);
        gelu_pair(pre.w, hid.w, dhid.w);
        float4 up = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < TAIL_MAX_OUT; ++o) {
          if (o >= P.O) continue;
          const float4 wv = *reinterpret_cast<const float4*>(S.par + (1 + o) * P.C + cidx[k]);
          up.x = fmaf(go[o], wv.x, up.x);
          up.y = fmaf(go[o], wv.y, up.y);
          up.z = fmaf(go[o], wv.z, up.z);
          up.w = fmaf(go[o], wv.w, up.w);
          dw2[o][k].x = fmaf(go[o], hid.x, dw2[o][k].x);
          dw2[o][k].y = fmaf(go[o], hid.y, dw2[o][k].y);
          dw2[o][k].z = fmaf(go[o], hid.z, dw2[o][k].z);
          dw2[o][k].w = fmaf(go[o], hid.w, dw2[o][k].w);
          if (cidx[k] == 0 && ok[k]) db2[o] += go[o];
        }
        g1[k] = make_float4(up.x * dhid.x, up.y * dhid.y, up.z * dhid.z, up.w * dhid.w);
        if (!ok[k]) g1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        db1[k].x += g1[k].x;
        db1[k].y += g1[k].y;
        db1[k].z += g1[k].z;
        db1[k].w += g1[k].w;
#pragma unroll
        for (int h = 0; h < NH; ++h)
          ds_head[h] += g1[k].x * acc_z[h][k].x + g1[k].y * acc_z[h][k].y + g1[k].z * acc_z[h][k].z + g1[k].w * acc_z[h][k].w;
      }
      // pass B: dY[b, j, h, :] += P^_hj g1
      for (int e = 0; e < n; ++e) {
        const float4 ent = seg[e];
        const int j = __float_as_int(ent.x);
        const int sidx = S.map[j];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float pw = h == 0 ? ent.z : ent.w;
#pragma unroll
          for (int k = 0; k < L4; ++k) {
            if (!ok[k]) continue;
            const float4 add = make_float4(pw * g1[k].x, pw * g1[k].y, pw * g1[k].z, pw * g1[k].w);
            if (sidx >= 0) {
              float4* cell = S.slot_acc + ((size_t)sidx * NH + h) * P.lanes4 + tid + k * TALL_THREADS;
              float4 cur = *cell;
              cur.x += add.x;
              cur.y += add.y;
              cur.z += add.z;
              cur.w += add.w;
              *cell = cur;
            } else {
              atomicAdd(reinterpret_cast<float4*>(P.d_y + (y_off[k] + j * row_stride + h * P.C)), add);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  tail_flush_slots<NH, L4>(P, S, y_off, ok, tid);

  // ---- parameter gradients: one reduction per CTA ----
  // scale: sum over all threads
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float v = warp_sum(ds_head[h]);
    __syncthreads();
    if (lane == 0) S.red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < TALL_WARPS; ++w) t += S.red[w];
      atomicAdd(P.d_scale + h, -t);
    }
  }
  // b2: sum over all threads
#pragma unroll
  for (int o = 0; o < TAIL_MAX_OUT; ++o) {
    if (o >= P.O) continue;  // uniform
    const float v = warp_sum(db2[o]);
    __syncthreads();
    if (lane == 0) S.red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < TALL_WARPS; ++w) t += S.red[w];
      atomicAdd(P.d_b2 + o, t);
    }
  }
  // b1 and W2: indexed by the hidden channel, summed over the samples -> straight REDs (C*O+C addresses per CTA)
#pragma unroll
  for (int k = 0; k < L4; ++k) {
    if (!ok[k]) continue;
    atomicAdd(reinterpret_cast<float4*>(P.d_b1 + cidx[k]), db1[k]);
#pragma unroll
    for (int o = 0; o < TAIL_MAX_OUT; ++o)
      if (o < P.O) atomicAdd(reinterpret_cast<float4*>(P.d_w2 + (int64_t)o * P.C + cidx[k]), dw2[o][k]);
  }
}

}  // namespace pit
