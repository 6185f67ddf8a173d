// K5: the whole processor of a shared-mesh PiT model (pit.py:114-122) in ONE launch forward and ONE launch backward.
//
//   for k in range(n_blocks):   X <- gelu( mlp_k( cat(X, A_k,0 X, ..., A_k,H-1 X) ) )        A_k,h = softmax_j(-s_k,h d2)
//
// on a small latent mesh (N = 8 x 16..32 points, hidden width D = 32 or 64, H <= 2 heads): every product of a block is
// a few MFLOP, so the unfused path (a dense-attention launch, two cuBLAS GEMMs and two epilogue launches per block and
// direction, ~36 % of a Darcy-421 step) is pure launch latency and sub-wave grids.  Here one thread-block CLUSTER owns a
// sample: CTA `t` of the cluster owns rows [t TR, (t+1) TR) of the latent mesh for the whole depth of the processor, the
// only exchange between the CTAs of a cluster is the block output X_{k+1} (forward) and the attention-input gradient dC
// (backward), which go through L2 between two cluster barriers; samples never interact.  All products run on
// mma.sync.m16n8k8 TF32 with fp32 accumulation from shared-memory operands: the attention products split 3xTF32 (parity
// with the fp32 reference to ~1e-6), the Linear products 3xTF32 or plain TF32 as torch's matmul precision asks
// ('highest' / 'high', pit.py:2 ships 'high').
//
// Forward, per block (CTA = row tile x sample):
//   P[h][r][j]   = exp2(-d2(r, j) s_h log2e)  from coordinates (v_min = 0: self stage), row sums l -> 1/l
//   C_h          = (P_h X) / l                              [TR x D]  per head          (X = block input, all N rows, smem)
//   Z1           = [X_tile | C_0 | C_1] W1^T + b1 ;  H1 = gelu(Z1)
//   Z2           = H1 W2^T + b2 ;  X_next tile = gelu(Z2)
//   saved for the backward: C, Z1, Z2, l, X_next (the block inputs)
// Backward, per block in reverse (G = gradient w.r.t. the block output on the CTA's rows, kept in shared memory):
//   phase 1 (row tile)   dZ2 = G gelu'(Z2); db2, dW2 += dZ2^T H1; dH1 = dZ2 W2; dZ1 = dH1 gelu'(Z1); db1; dW1 += dZ1^T cat;
//                        [dXdirect | dC_0 | dC_1] = dZ1 W1;  delta = <dC_h, C_h>;  dP_h = dC_h X^T;
//                        ds_h -= sum P^ (dP - delta) d2;  dC -> L2 scratch
//   cluster barrier
//   phase 2 (column tile = the same rows, as COLUMNS of the attention)   G_prev = dXdirect + sum_h P^_h[:, tile]^T dC_h
// Parameter gradients leave each CTA as vector REDs into zero-initialised buffers.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "decoder_tail_mma.cuh"
#include "geometry.cuh"
#include "mlp_epilogue.cuh"

namespace pit {

constexpr int PB_THREADS = 256;
constexpr int PB_WARPS = PB_THREADS / 32;
constexpr int PB_MAX_BLOCKS = 8;

struct ProcWeights {
  const float *w1, *b1, *w2, *b2;  // mlp1.weight [D,(1+H)D], mlp1.bias [D], mlp2.weight [D,D], mlp2.bias [D]
};
struct ProcGrads {
  float *d_w1, *d_b1, *d_w2, *d_b2;  // zero-initialised
};

struct ProcParams {
  int geo, sd, B, N, n_blocks;
  const float* mesh;    // [N,sd]
  const float* period;  // device scalar or null
  const float* x0;      // [B,N,D]
  const float* scale;   // [n_blocks,H]
  ProcWeights w[PB_MAX_BLOCKS];
  float* saved;  // n_blocks x proc_saved_stride floats
  float* out;    // [B,N,D]
  // backward only
  const float* d_out;  // [B,N,D]
  float* d_x0;         // [B,N,D]
  float* d_scale;      // [n_blocks,H] zero-initialised
  ProcGrads g[PB_MAX_BLOCKS];
  float* scratch;  // [2][B,N,H*D]
};

// Layout of one block's slice of `saved` (floats): C [B,N,H*D] | Z1 [B,N,D] | Z2 [B,N,D] | l [H,N] | X_next [B,N,D]
struct ProcSaved {
  int64_t c, z1, z2, l, x, stride;
};
__host__ __device__ inline ProcSaved proc_saved_layout(int64_t B, int64_t N, int64_t H, int64_t D) {
  ProcSaved s;
  s.c = 0;
  s.z1 = B * N * H * D;
  s.z2 = s.z1 + B * N * D;
  s.l = s.z2 + B * N * D;
  s.x = s.l + H * N;
  s.stride = s.x + B * N * D;
  s.stride = (s.stride + 3) / 4 * 4;  // keeps every slice 16-byte aligned
  return s;
}

__device__ __forceinline__ float pb_dist2(int geo, float ox, float oy, float ix, float iy, float period) {
  // plain arithmetic: no mask is decided on these distances (global stage)
  if (geo == GEO_EUCLID1) {
    const float dx = ox - ix;
    return dx * dx;
  } else if (geo == GEO_EUCLID2) {
    const float dx = ox - ix, dy = oy - iy;
    return fmaf(dy, dy, dx * dx);
  } else if (geo == GEO_PERIODIC1) {
    float m = fabsf(ox - ix);
    m = fminf(m, period - m);
    return m * m;
  }
  float mx = fabsf(ox - ix), my = fabsf(oy - iy);
  mx = fminf(mx, period - mx);
  my = fminf(my, period - my);
  return fmaf(my, my, mx * mx);
}

__device__ __forceinline__ float pb_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void pb_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void pb_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pb_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Squared distances of the CTA's TR points to every point of the mesh, handed to f(r, j, h, d2) for r < TR, j < N, h < NH.
// A thread stays on one column j (its coordinates in registers) when N divides the CTA size: no index arithmetic in the loop.
// (d2 is symmetric in its two points for every variant, so the same walk serves the transposed tile of the backward.)
template <int GEO, int NH, int TR, typename F>
__device__ __forceinline__ void pb_weights_geo(int N, int row0, const float* XY, float period, F f) {
  if (PB_THREADS % N == 0) {
    const int j = threadIdx.x % N;
    const float jx = XY[2 * j], jy = XY[2 * j + 1];
#pragma unroll 4
    for (int r = threadIdx.x / N; r < TR; r += PB_THREADS / N) {
      const float d2 = pb_dist2(GEO, XY[2 * (row0 + r)], XY[2 * (row0 + r) + 1], jx, jy, period);
#pragma unroll
      for (int h = 0; h < NH; ++h) f(r, j, h, d2);
    }
  } else {
    for (int idx = threadIdx.x; idx < TR * N; idx += PB_THREADS) {
      const int r = idx / N, j = idx - r * N;
      const float d2 = pb_dist2(GEO, XY[2 * (row0 + r)], XY[2 * (row0 + r) + 1], XY[2 * j], XY[2 * j + 1], period);
#pragma unroll
      for (int h = 0; h < NH; ++h) f(r, j, h, d2);
    }
  }
}
template <int NH, int TR, typename F>
__device__ __forceinline__ void pb_weights(int geo, int N, int row0, const float* XY, float period, F f) {
  if (geo == GEO_EUCLID2) pb_weights_geo<GEO_EUCLID2, NH, TR>(N, row0, XY, period, f);
  else if (geo == GEO_EUCLID1) pb_weights_geo<GEO_EUCLID1, NH, TR>(N, row0, XY, period, f);
  else if (geo == GEO_PERIODIC1) pb_weights_geo<GEO_PERIODIC1, NH, TR>(N, row0, XY, period, f);
  else pb_weights_geo<GEO_PERIODIC2, NH, TR>(N, row0, XY, period, f);
}

// Pulls `bytes` of global memory (weights of the block that is about to run) into L1 while the attention product runs.
__device__ __forceinline__ void pb_prefetch_l1(const float* base, int bytes) {
  for (int off = threadIdx.x * 128; off < bytes; off += PB_THREADS * 128)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(base) + off));
}

// 3xTF32 operand split: hi = x rounded to the nearest TF32, lo = x - hi (exact; the tensor pipe drops its low bits, an
// unbiased 2^-21 relative error because lo takes either sign).  Leaving the rounding of hi to the pipe as well (truncation)
// is cheaper but biases every operand toward zero: ~1e-6 per product, which adds up over the twelve chained products of a
// four-block processor (measured 2.2e-5 against the oracle instead of 2.8e-6).  Without X3 the operand is rounded to TF32.
template <bool X3>
__device__ __forceinline__ void pb_split(float x, uint32_t& hi, uint32_t& lo) {
  const float h = tm_round_hi(x);
  hi = __float_as_uint(h);
  lo = X3 ? __float_as_uint(x - h) : 0u;
}

// Lane pointers of the mma.sync.m16n8k8 fragments (g = lane / 4, t = lane % 4):
//   A fragment of m-tile mt: rows 16 mt + g (+8), k = t (+4);   B fragment of n-tile nt: k = t (+4), column 8 nt + g.
template <int MT>
struct PbA {
  const float* p[MT][2];
};
template <int NT>
struct PbB {
  const float* p[NT];
};
// A stored row-major [m][k] with row pitch lda (k contiguous: A_SK = 1)
template <int MT>
__device__ __forceinline__ PbA<MT> pb_a_rows(const float* base, int lda) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  PbA<MT> a;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) a.p[mt][half] = base + (size_t)(mt * 16 + g + 8 * half) * lda + t;
  return a;
}
// A stored transposed [k][m] with pitch lda (A_SK = lda)
template <int MT>
__device__ __forceinline__ PbA<MT> pb_a_cols(const float* base, int lda) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  PbA<MT> a;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) a.p[mt][half] = base + (size_t)t * lda + mt * 16 + g + 8 * half;
  return a;
}
// B stored [k][n] with pitch ldb (B_SK = ldb)
template <int NT>
__device__ __forceinline__ PbB<NT> pb_b_rows(const float* base, int ldb) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  PbB<NT> b;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) b.p[nt] = base + (size_t)t * ldb + nt * 8 + g;
  return b;
}
// B stored [n][k] with pitch ldb (k contiguous: B_SK = 1)
template <int NT>
__device__ __forceinline__ PbB<NT> pb_b_cols(const float* base, int ldb) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  PbB<NT> b;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) b.p[nt] = base + (size_t)(nt * 8 + g) * ldb + t;
  return b;
}

// Warp-level product on mma.sync.m16n8k8: acc[mt][nt] += A[16 mt + ., k] B[k, 8 nt + .] over `ksteps` (a multiple of PB_CHAIN) steps
// of 8.  A_SK / B_SK: compile-time distance (floats) between consecutive k of the operand, so every load of a chain of four
// k-steps has an immediate offset.  B_LDG: operand B lives in global memory (weights), read through the read-only path.
// X3: hi*hi + lo*hi + hi*lo (pb_split); otherwise one product on operands rounded to the nearest TF32.
// Accumulator fragment: acc[..][0] = (row g, col 2t), [1] = (g, 2t+1), [2] = (g+8, 2t), [3] = (g+8, 2t+1).
template <int MT, int NT, bool X3, int A_SK, int B_SK, bool B_LDG = false, int PB_CHAIN = 4>
__device__ __forceinline__ void pb_gemm(float (&acc)[MT][NT][4], int ksteps, PbA<MT> a, PbB<NT> b) {
  // The tensor pipe accumulates with truncation: a chain of 3 x 32 accumulating MMAs drifts by a few 1e-6 (relative, biased).
  // The 3xTF32 products therefore run in chains of four k-steps from zero and are added to `acc` by the fp32 pipe
  // (round to nearest), as Ootomo & Yokota do for their error-corrected TF32 GEMM.
  for (int kc = 0; kc < ksteps; kc += PB_CHAIN) {
    float part[MT][NT][4];
    if (X3) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) part[mt][nt][i] = 0.f;
    }
    float(&dst)[MT][NT][4] = X3 ? part : acc;
#pragma unroll
    for (int kq = 0; kq < PB_CHAIN; ++kq) {
      uint32_t ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float v[4] = {a.p[mt][0][(8 * kq) * A_SK], a.p[mt][1][(8 * kq) * A_SK], a.p[mt][0][(8 * kq + 4) * A_SK],
                            a.p[mt][1][(8 * kq + 4) * A_SK]};
#pragma unroll
        for (int i = 0; i < 4; ++i) pb_split<X3>(v[i], ah[mt][i], al[mt][i]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float* q = b.p[nt];
        const float v[2] = {B_LDG ? __ldg(q + (8 * kq) * B_SK) : q[(8 * kq) * B_SK], B_LDG ? __ldg(q + (8 * kq + 4) * B_SK) : q[(8 * kq + 4) * B_SK]};
#pragma unroll
        for (int i = 0; i < 2; ++i) pb_split<X3>(v[i], bh[nt][i], bl[nt][i]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          if (X3) {
            mma_tf32_16x8x8(dst[mt][nt], al[mt], bh[nt]);
            mma_tf32_16x8x8(dst[mt][nt], ah[mt], bl[nt]);
          }
          mma_tf32_16x8x8(dst[mt][nt], ah[mt], bh[nt]);
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) a.p[mt][half] += 8 * PB_CHAIN * A_SK;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b.p[nt] += 8 * PB_CHAIN * B_SK;
    if (X3) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mt][nt][i] += part[mt][nt][i];
    }
  }
}

template <int MT, int NT>
__device__ __forceinline__ void pb_zero(float (&acc)[MT][NT][4]) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
}

// Compile-time shape facts shared by both kernels.
template <int D, int NH, int TR>
struct ProcShape {
  static constexpr int HD = NH * D, CW = D + HD;
  static constexpr int MTR = TR / 16;                 // m-tiles over the CTA's rows
  static constexpr int LDX_F = D + 8;                 // forward X pitch: B fragments (k = point, n = feature) conflict-free
  static constexpr int LDX_B = D + 4;                 // backward X pitch: B fragments (k = feature, n = point) conflict-free
  static constexpr int LDC = CW + 4, LDD = D + 4, LDH = HD + 4;
  // attention-shaped products [TR x N] [N x HD]: n-tiles per item, m-tiles per item
  static constexpr int NT_ATT = HD >= 128 ? 2 : 1;
  static constexpr int NG_ATT = HD / 8 / NT_ATT;
  static constexpr int MT_ATT = NG_ATT >= PB_WARPS ? MTR : 1;
  static constexpr int ITEMS_ATT = (MTR / MT_ATT) * NG_ATT;
  // Linear-shaped products with D output columns: [TR x K] [K x D]
  static constexpr int MT_LIN = D / 8 >= PB_WARPS ? MTR : 1;
  static constexpr int ITEMS_LIN = (MTR / MT_LIN) * (D / 8);
  // weight-gradient products [D x TR] [TR x cols]
  static constexpr int MT_W = D >= 64 ? 2 : 1, NT_W = D >= 64 ? 2 : 1;
  static constexpr int CHAIN_W = TR / 8 >= 4 ? 4 : TR / 8;   // their reduction runs over the TR rows only
  static_assert(TR % 16 == 0 && D % 16 == 0 && PB_THREADS % D == 0, "unsupported tile shape");
  static_assert(ITEMS_LIN <= PB_WARPS, "the backward keeps one Linear-shaped item per warp in registers");
};

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <int D, int NH, int TR>
__host__ __device__ constexpr size_t proc_fwd_smem_floats(int N) {
  using S = ProcShape<D, NH, TR>;
  return (size_t)NH * TR * (N + 4) + (size_t)N * S::LDX_F + (size_t)TR * S::LDC + (size_t)TR * S::LDD + 2 * (size_t)N + NH * TR;
}

template <int D, int NH, int TR, bool LIN3>
__global__ void __launch_bounds__(PB_THREADS, 1) processor_fwd_kernel(const ProcParams P) {
  using S = ProcShape<D, NH, TR>;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int HD = S::HD, CW = S::CW, LDX = S::LDX_F, LDC = S::LDC, LDD = S::LDD;
  const int N = P.N, LDP = N + 4;
  extern __shared__ __align__(16) float pb_smem[];
  float* PS = pb_smem;                           // [NH][TR][LDP] unnormalised weights of the CTA's rows
  float* XS = PS + (size_t)NH * TR * LDP;        // [N][LDX]      block input of this sample
  float* CAT = XS + (size_t)N * LDX;             // [TR][LDC]     [X tile | C_0 | C_1]
  float* HS = CAT + (size_t)TR * LDC;            // [TR][LDD]     gelu(Z1)
  float* XY = HS + (size_t)TR * LDD;             // [N][2]
  float* INVL = XY + 2 * (size_t)N;              // [NH][TR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, row0 = blockIdx.x * TR;
  const ProcSaved L = proc_saved_layout(P.B, N, NH, D);
  const float period = P.period ? __ldg(P.period) : 0.f;
  for (int i = tid; i < N; i += PB_THREADS) {
    XY[2 * i] = __ldg(P.mesh + (int64_t)i * P.sd);
    XY[2 * i + 1] = P.sd == 2 ? __ldg(P.mesh + (int64_t)i * P.sd + 1) : 0.f;
  }
  const float* xin = P.x0 + (int64_t)b * N * D;
  for (int k = 0; k < P.n_blocks; ++k) {
    float* sv = P.saved + (int64_t)k * L.stride;
    const ProcWeights W = P.w[k];
    // block input of this sample -> shared memory (L2 path: rows written by the other CTAs of the cluster)
    for (int c = tid; c < N * (D / 4); c += PB_THREADS) {
      const int r = c / (D / 4), q = c - r * (D / 4);
      pb_cp16(XS + (size_t)r * LDX + 4 * q, xin + (int64_t)r * D + 4 * q);
    }
    pb_cp_commit();
    pb_prefetch_l1(W.w1, D * CW * 4);
    pb_prefetch_l1(W.w2, D * D * 4);
    __syncthreads();  // XY visible (first block)
    float sc2[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) sc2[h] = __ldg(P.scale + k * NH + h) * 1.4426950408889634f;
    pb_weights<NH, TR>(P.geo, N, row0, XY, period, [&](int r, int j, int h, float d2) {
      PS[(size_t)(h * TR + r) * LDP + j] = pb_exp2(-d2 * sc2[h]);
    });
    __syncthreads();
    for (int p0 = warp * 4; p0 < NH * TR; p0 += PB_WARPS * 4) {   // four rows per pass: independent load / add chains
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = lane; j < N; j += 32)
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] += PS[(size_t)(p0 + q) * LDP + j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] += __shfl_xor_sync(FULL, s[q], o);
      if (lane < 4) {
        const int pr = p0 + lane;
        const float sum = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
        INVL[pr] = 1.f / sum;
        if (b == 0) sv[L.l + (int64_t)(pr / TR) * N + row0 + (pr % TR)] = sum;
      }
    }
    pb_cp_wait();
    __syncthreads();
    // X tile -> first D columns of the concatenation
    for (int idx = tid; idx < TR * (D / 4); idx += PB_THREADS) {
      const int r = idx / (D / 4), q = idx - r * (D / 4);
      *reinterpret_cast<float4*>(CAT + (size_t)r * LDC + 4 * q) = *reinterpret_cast<const float4*>(XS + (size_t)(row0 + r) * LDX + 4 * q);
    }
    // attention: C_h = (P_h X) / l
    for (int item = warp; item < S::ITEMS_ATT; item += PB_WARPS) {
      const int mg = item / S::NG_ATT, ng = item - mg * S::NG_ATT;
      const int c0 = ng * S::NT_ATT * 8, h = c0 / D, dc = c0 - h * D;
      float acc[S::MT_ATT][S::NT_ATT][4];
      pb_zero(acc);
      pb_gemm<S::MT_ATT, S::NT_ATT, true, 1, LDX>(acc, N / 8, pb_a_rows<S::MT_ATT>(PS + (size_t)(h * TR + mg * S::MT_ATT * 16) * LDP, LDP),
                                                   pb_b_rows<S::NT_ATT>(XS + dc, LDX));
#pragma unroll
      for (int mt = 0; mt < S::MT_ATT; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = (mg * S::MT_ATT + mt) * 16 + g + 8 * half;
          const float il = INVL[h * TR + r];
#pragma unroll
          for (int nt = 0; nt < S::NT_ATT; ++nt) {
            const int col = c0 + nt * 8 + 2 * t;
            const float2 v = make_float2(acc[mt][nt][2 * half] * il, acc[mt][nt][2 * half + 1] * il);
            *reinterpret_cast<float2*>(CAT + (size_t)r * LDC + D + col) = v;
            *reinterpret_cast<float2*>(sv + L.c + ((int64_t)b * N + row0 + r) * HD + col) = v;
          }
        }
    }
    __syncthreads();
    // Z1 = cat W1^T + b1, H1 = gelu(Z1)
    for (int item = warp; item < S::ITEMS_LIN; item += PB_WARPS) {
      const int mg = item / (D / 8), n0 = (item - mg * (D / 8)) * 8;
      float acc[S::MT_LIN][1][4];
      pb_zero(acc);
      pb_gemm<S::MT_LIN, 1, LIN3, 1, 1, true>(acc, CW / 8, pb_a_rows<S::MT_LIN>(CAT + (size_t)(mg * S::MT_LIN * 16) * LDC, LDC),
                                               pb_b_cols<1>(W.w1 + (int64_t)n0 * CW, CW));
      const float2 bias = __ldg(reinterpret_cast<const float2*>(W.b1 + n0 + 2 * t));
#pragma unroll
      for (int mt = 0; mt < S::MT_LIN; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = (mg * S::MT_LIN + mt) * 16 + g + 8 * half, col = n0 + 2 * t;
          const float2 z = make_float2(acc[mt][0][2 * half] + bias.x, acc[mt][0][2 * half + 1] + bias.y);
          *reinterpret_cast<float2*>(sv + L.z1 + ((int64_t)b * N + row0 + r) * D + col) = z;
          *reinterpret_cast<float2*>(HS + (size_t)r * LDD + col) = make_float2(tm_gelu(z.x), tm_gelu(z.y));
        }
    }
    __syncthreads();
    // Z2 = H1 W2^T + b2, X_next = gelu(Z2)
    float* xout = (k == P.n_blocks - 1) ? P.out : sv + L.x;
    for (int item = warp; item < S::ITEMS_LIN; item += PB_WARPS) {
      const int mg = item / (D / 8), n0 = (item - mg * (D / 8)) * 8;
      float acc[S::MT_LIN][1][4];
      pb_zero(acc);
      pb_gemm<S::MT_LIN, 1, LIN3, 1, 1, true>(acc, D / 8, pb_a_rows<S::MT_LIN>(HS + (size_t)(mg * S::MT_LIN * 16) * LDD, LDD),
                                               pb_b_cols<1>(W.w2 + (int64_t)n0 * D, D));
      const float2 bias = __ldg(reinterpret_cast<const float2*>(W.b2 + n0 + 2 * t));
#pragma unroll
      for (int mt = 0; mt < S::MT_LIN; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = (mg * S::MT_LIN + mt) * 16 + g + 8 * half, col = n0 + 2 * t;
          const float2 z = make_float2(acc[mt][0][2 * half] + bias.x, acc[mt][0][2 * half + 1] + bias.y);
          *reinterpret_cast<float2*>(sv + L.z2 + ((int64_t)b * N + row0 + r) * D + col) = z;
          *reinterpret_cast<float2*>(xout + ((int64_t)b * N + row0 + r) * D + col) = make_float2(tm_gelu(z.x), tm_gelu(z.y));
        }
    }
    if (k + 1 < P.n_blocks) {
      cluster.sync();  // every row of X_next of this sample is in L2
      xin = sv + L.x + (int64_t)b * N * D;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
template <int D, int NH, int TR>
__host__ __device__ constexpr size_t proc_bwd_smem_floats(int N) {
  using S = ProcShape<D, NH, TR>;
  return (size_t)NH * TR * (N + 4) + (size_t)N * S::LDX_B + (size_t)TR * S::LDC + 3 * (size_t)TR * S::LDD + (size_t)TR * S::LDH +
         2 * (size_t)N + (size_t)PB_MAX_BLOCKS * NH * N + NH * TR + PB_THREADS + 4;
}

template <int D, int NH, int TR, bool LIN3>
__global__ void __launch_bounds__(PB_THREADS, 1) processor_bwd_kernel(const ProcParams P) {
  using S = ProcShape<D, NH, TR>;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int HD = S::HD, CW = S::CW, LDX = S::LDX_B, LDC = S::LDC, LDD = S::LDD, LDH = S::LDH;
  const int N = P.N, LDP = N + 4;
  extern __shared__ __align__(16) float pb_smem[];
  float* PS = pb_smem;                       // [NH][TR][LDP] phase 1: P^ of the CTA's rows; phase 2: P^[:, tile]^T
  float* XS = PS + (size_t)NH * TR * LDP;    // [N][LDX]      phase 1: block input; phase 2: dC_h of the sample
  float* CAT = XS + (size_t)N * LDX;         // [TR][LDC]     [X tile | C_0 | C_1]
  float* GB = CAT + (size_t)TR * LDC;        // [TR][LDD]     G, later dXdirect, later G of the previous block
  float* AB = GB + (size_t)TR * LDD;         // [TR][LDD]     dZ2, later dZ1
  float* H1 = AB + (size_t)TR * LDD;         // [TR][LDD]     gelu(Z1)
  float* DC = H1 + (size_t)TR * LDD;         // [TR][LDH]     dC tile
  float* XY = DC + (size_t)TR * LDH;         // [N][2]
  float* ILALL = XY + 2 * (size_t)N;         // [n_blocks][NH][N]  1 / l of every row of every block
  float* DELTA = ILALL + (size_t)PB_MAX_BLOCKS * NH * N;  // [NH][TR]
  float* RED = DELTA + NH * TR;              // [PB_THREADS]
  float* DSACC = RED + PB_THREADS;           // [NH]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, row0 = blockIdx.x * TR;
  const ProcSaved L = proc_saved_layout(P.B, N, NH, D);
  const float period = P.period ? __ldg(P.period) : 0.f;
  for (int i = tid; i < N; i += PB_THREADS) {
    XY[2 * i] = __ldg(P.mesh + (int64_t)i * P.sd);
    XY[2 * i + 1] = P.sd == 2 ? __ldg(P.mesh + (int64_t)i * P.sd + 1) : 0.f;
  }
  for (int idx = tid; idx < TR * D; idx += PB_THREADS) {
    const int r = idx / D, c = idx - r * D;
    GB[(size_t)r * LDD + c] = __ldg(P.d_out + ((int64_t)b * N + row0 + r) * D + c);
  }
  for (int i = tid; i < P.n_blocks * NH * N; i += PB_THREADS) {
    const int k = i / (NH * N);
    ILALL[i] = 1.f / __ldg(P.saved + (int64_t)k * L.stride + L.l + (i - k * NH * N));
  }
  for (int k = P.n_blocks - 1; k >= 0; --k) {
    const float* sv = P.saved + (int64_t)k * L.stride;
    const ProcWeights W = P.w[k];
    const ProcGrads G = P.g[k];
    const float* xin = (k == 0 ? P.x0 : P.saved + (int64_t)(k - 1) * L.stride + L.x) + (int64_t)b * N * D;
    float* dcs = P.scratch + ((int64_t)(k & 1) * P.B + b) * N * HD;
    const int64_t grow = (int64_t)b * N + row0;  // first global row of the tile
    // stage the block input (all rows) and the concatenation tile
    for (int c = tid; c < N * (D / 4); c += PB_THREADS) {
      const int r = c / (D / 4), q = c - r * (D / 4);
      pb_cp16(XS + (size_t)r * LDX + 4 * q, xin + (int64_t)r * D + 4 * q);
    }
    for (int c = tid; c < TR * (CW / 4); c += PB_THREADS) {
      const int r = c / (CW / 4), q = c - r * (CW / 4);
      const float* src = 4 * q < D ? xin + (int64_t)(row0 + r) * D + 4 * q : sv + L.c + (grow + r) * HD + (4 * q - D);
      pb_cp16(CAT + (size_t)r * LDC + 4 * q, src);
    }
    pb_cp_commit();
    pb_prefetch_l1(W.w1, D * CW * 4);
    pb_prefetch_l1(W.w2, D * D * 4);
    float sc2[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) sc2[h] = __ldg(P.scale + k * NH + h) * 1.4426950408889634f;
    const float* IL = ILALL + (size_t)k * NH * N;
    if (tid < NH) DSACC[tid] = 0.f;
    __syncthreads();  // XY, IL, GB
    // normalised weights of the CTA's rows
    pb_weights<NH, TR>(P.geo, N, row0, XY, period, [&](int r, int j, int h, float d2) {
      PS[(size_t)(h * TR + r) * LDP + j] = pb_exp2(-d2 * sc2[h]) * IL[h * N + row0 + r];
    });
    // dZ2 = G gelu'(Z2), H1 = gelu(Z1), db2 (a thread stays on one column)
    {
      const int c = tid % D;
      float colsum = 0.f;
      for (int r = tid / D; r < TR; r += PB_THREADS / D) {
        const float z2 = __ldg(sv + L.z2 + (grow + r) * D + c), z1 = __ldg(sv + L.z1 + (grow + r) * D + c);
        float act, dact;
        tm_gelu_pair(z2, act, dact);
        const float dz2 = GB[(size_t)r * LDD + c] * dact;
        AB[(size_t)r * LDD + c] = dz2;
        H1[(size_t)r * LDD + c] = tm_gelu(z1);
        colsum += dz2;
      }
      RED[tid] = colsum;
    }
    __syncthreads();
    if (tid < D) {
      float s = 0.f;
      for (int q = tid; q < PB_THREADS; q += D) s += RED[q];
      atomicAdd(G.d_b2 + tid, s);
    }
    // dW2 += dZ2^T H1
    for (int item = warp; item < (D / 16 / S::MT_W) * (D / 8 / S::NT_W); item += PB_WARPS) {
      const int mg = item / (D / 8 / S::NT_W), ng = item - mg * (D / 8 / S::NT_W);
      const int m0 = mg * S::MT_W * 16, n0 = ng * S::NT_W * 8;
      float acc[S::MT_W][S::NT_W][4];
      pb_zero(acc);
      pb_gemm<S::MT_W, S::NT_W, LIN3, LDD, LDD, false, S::CHAIN_W>(acc, TR / 8, pb_a_cols<S::MT_W>(AB + m0, LDD), pb_b_rows<S::NT_W>(H1 + n0, LDD));
#pragma unroll
      for (int mt = 0; mt < S::MT_W; ++mt)
#pragma unroll
        for (int nt = 0; nt < S::NT_W; ++nt)
#pragma unroll
          for (int half = 0; half < 2; ++half)
            atomicAdd(reinterpret_cast<float2*>(G.d_w2 + (int64_t)(m0 + mt * 16 + g + 8 * half) * D + n0 + nt * 8 + 2 * t),
                      make_float2(acc[mt][nt][2 * half], acc[mt][nt][2 * half + 1]));
    }
    // dH1 = dZ2 W2 (kept in registers until every warp has read dZ2), then dZ1 = dH1 gelu'(Z1) over dZ2's buffer
    {
      float acc[(S::ITEMS_LIN + PB_WARPS - 1) / PB_WARPS][S::MT_LIN][1][4];
      int slot = 0;
      for (int item = warp; item < S::ITEMS_LIN; item += PB_WARPS, ++slot) {
        const int mg = item / (D / 8), n0 = (item - mg * (D / 8)) * 8;
        pb_zero(acc[slot]);
        pb_gemm<S::MT_LIN, 1, LIN3, 1, D, true>(acc[slot], D / 8, pb_a_rows<S::MT_LIN>(AB + (size_t)(mg * S::MT_LIN * 16) * LDD, LDD),
                                                 pb_b_rows<1>(W.w2 + n0, D));
      }
      __syncthreads();
      slot = 0;
      for (int item = warp; item < S::ITEMS_LIN; item += PB_WARPS, ++slot) {
        const int mg = item / (D / 8), n0 = (item - mg * (D / 8)) * 8, col = n0 + 2 * t;
        float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
        for (int mt = 0; mt < S::MT_LIN; ++mt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int r = (mg * S::MT_LIN + mt) * 16 + g + 8 * half;
            const float2 z1 = __ldg(reinterpret_cast<const float2*>(sv + L.z1 + (grow + r) * D + col));
            float a0, d0, a1, d1;
            tm_gelu_pair(z1.x, a0, d0);
            tm_gelu_pair(z1.y, a1, d1);
            const float2 v = make_float2(acc[slot][mt][0][2 * half] * d0, acc[slot][mt][0][2 * half + 1] * d1);
            *reinterpret_cast<float2*>(AB + (size_t)r * LDD + col) = v;
            cs0 += v.x, cs1 += v.y;
          }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          cs0 += __shfl_xor_sync(FULL, cs0, o);
          cs1 += __shfl_xor_sync(FULL, cs1, o);
        }
        if (g == 0) {
          atomicAdd(G.d_b1 + col, cs0);
          atomicAdd(G.d_b1 + col + 1, cs1);
        }
      }
    }
    pb_cp_wait();
    __syncthreads();  // dZ1, cat tile, X
    // dW1 += dZ1^T cat
    for (int item = warp; item < (D / 16 / S::MT_W) * (CW / 8 / S::NT_W); item += PB_WARPS) {
      const int mg = item / (CW / 8 / S::NT_W), ng = item - mg * (CW / 8 / S::NT_W);
      const int m0 = mg * S::MT_W * 16, n0 = ng * S::NT_W * 8;
      float acc[S::MT_W][S::NT_W][4];
      pb_zero(acc);
      pb_gemm<S::MT_W, S::NT_W, LIN3, LDD, LDC, false, S::CHAIN_W>(acc, TR / 8, pb_a_cols<S::MT_W>(AB + m0, LDD), pb_b_rows<S::NT_W>(CAT + n0, LDC));
#pragma unroll
      for (int mt = 0; mt < S::MT_W; ++mt)
#pragma unroll
        for (int nt = 0; nt < S::NT_W; ++nt)
#pragma unroll
          for (int half = 0; half < 2; ++half)
            atomicAdd(reinterpret_cast<float2*>(G.d_w1 + (int64_t)(m0 + mt * 16 + g + 8 * half) * CW + n0 + nt * 8 + 2 * t),
                      make_float2(acc[mt][nt][2 * half], acc[mt][nt][2 * half + 1]));
    }
    // [dXdirect | dC] = dZ1 W1
    for (int item = warp; item < (S::MTR / S::MT_LIN) * (CW / 8); item += PB_WARPS) {
      const int mg = item / (CW / 8), n0 = (item - mg * (CW / 8)) * 8;
      float acc[S::MT_LIN][1][4];
      pb_zero(acc);
      pb_gemm<S::MT_LIN, 1, LIN3, 1, CW, true>(acc, D / 8, pb_a_rows<S::MT_LIN>(AB + (size_t)(mg * S::MT_LIN * 16) * LDD, LDD),
                                                pb_b_rows<1>(W.w1 + n0, CW));
#pragma unroll
      for (int mt = 0; mt < S::MT_LIN; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = (mg * S::MT_LIN + mt) * 16 + g + 8 * half, col = n0 + 2 * t;
          const float2 v = make_float2(acc[mt][0][2 * half], acc[mt][0][2 * half + 1]);
          if (n0 < D) {
            *reinterpret_cast<float2*>(GB + (size_t)r * LDD + col) = v;
          } else {
            *reinterpret_cast<float2*>(DC + (size_t)r * LDH + col - D) = v;
            *reinterpret_cast<float2*>(dcs + (int64_t)(row0 + r) * HD + col - D) = v;
          }
        }
    }
    __syncthreads();
    // delta[h][r] = <dC_h[r], C_h[r]>
    for (int pr = warp; pr < NH * TR; pr += PB_WARPS) {
      const int h = pr / TR, r = pr - h * TR;
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s += DC[(size_t)r * LDH + h * D + d] * CAT[(size_t)r * LDC + D + h * D + d];
      s = warp_sum(s);
      if (lane == 0) DELTA[pr] = s;
    }
    __syncthreads();
    // dP_h = dC_h X^T and the scale gradient ds_h = -sum P^ (dP - delta) d2
    {
      constexpr int NT_P = 2;
      float part[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) part[h] = 0.f;
      const int groups = N / (8 * NT_P);
      for (int item = warp; item < NH * groups; item += PB_WARPS) {
        const int h = item / groups, j0 = (item - h * groups) * 8 * NT_P;
        float acc[S::MTR][NT_P][4];
        pb_zero(acc);
        pb_gemm<S::MTR, NT_P, true, 1, 1>(acc, D / 8, pb_a_rows<S::MTR>(DC + h * D, LDH), pb_b_cols<NT_P>(XS + (size_t)j0 * LDX, LDX));
        float sum = 0.f;
#pragma unroll
        for (int mt = 0; mt < S::MTR; ++mt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int r = mt * 16 + g + 8 * half;
            const float dl = DELTA[h * TR + r], ox = XY[2 * (row0 + r)], oy = XY[2 * (row0 + r) + 1];
#pragma unroll
            for (int nt = 0; nt < NT_P; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = j0 + nt * 8 + 2 * t + e;
                const float d2 = pb_dist2(P.geo, ox, oy, XY[2 * j], XY[2 * j + 1], period);
                sum = fmaf(PS[(size_t)(h * TR + r) * LDP + j] * (acc[mt][nt][2 * half + e] - dl), d2, sum);
              }
          }
#pragma unroll
        for (int hh = 0; hh < NH; ++hh)
          if (hh == h) part[hh] += sum;
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float s = warp_sum(part[h]);
        if (lane == 0) atomicAdd(DSACC + h, -s);
      }
    }
    cluster.sync();  // every dC tile of this sample is in L2; also a CTA barrier (PS, XS, DSACC)
    if (tid < NH) atomicAdd(P.d_scale + k * NH + tid, DSACC[tid]);
    // phase 2 needs the transposed weights of the CTA's COLUMNS, PT[h][jt][i] = P^_h[i][row0 + jt] = P_h[i][row0 + jt] / l_i.  The
    // unnormalised weights are symmetric (d2 is, bit for bit), so PT[h][jt][i] = P_h[row0 + jt][i] / l_i: the tile of phase 1,
    // P^_h[row0 + jt][i] = P_h[row0 + jt][i] / l_{row0 + jt}, rescaled in place by l_{row0 + jt} / l_i -- no second round of exponentials.
    auto stage_dc = [&](int h) {     // dC_h of the whole sample -> XS (free since the dP products; the peers' rows come through L2)
      for (int c = tid; c < N * (D / 4); c += PB_THREADS) {
        const int r = c / (D / 4), q = c - r * (D / 4);
        pb_cp16(XS + (size_t)r * LDX + 4 * q, dcs + (int64_t)r * HD + h * D + 4 * q);
      }
      pb_cp_commit();
    };
    stage_dc(0);                     // in flight while the tile is rescaled
    for (int idx = tid; idx < NH * TR * (N / 4); idx += PB_THREADS) {
      const int row = idx / (N / 4), i4 = (idx - row * (N / 4)) * 4;      // row = h * TR + jt
      const int h = row / TR, jt = row - h * TR;
      const float lrow = 1.f / IL[h * N + row0 + jt];
      float4* pp = reinterpret_cast<float4*>(PS + (size_t)row * LDP + i4);
      const float4 il = *reinterpret_cast<const float4*>(IL + h * N + i4);
      float4 v = *pp;
      v.x *= lrow * il.x, v.y *= lrow * il.y, v.z *= lrow * il.z, v.w *= lrow * il.w;
      *pp = v;
    }
    {
      float acc[S::MT_LIN][1][4];
      pb_zero(acc);
      const int item = warp;  // ITEMS_LIN <= PB_WARPS
      const int mg = item / (D / 8), n0 = (item - mg * (D / 8)) * 8;
      for (int h = 0; h < NH; ++h) {
        if (h > 0) stage_dc(h);
        pb_cp_wait();
        __syncthreads();
        if (item < S::ITEMS_LIN) {
          pb_gemm<S::MT_LIN, 1, true, 1, LDX>(acc, N / 8, pb_a_rows<S::MT_LIN>(PS + (size_t)(h * TR + mg * S::MT_LIN * 16) * LDP, LDP),
                                              pb_b_rows<1>(XS + n0, LDX));
        }
        __syncthreads();
      }
      if (item < S::ITEMS_LIN) {
#pragma unroll
        for (int mt = 0; mt < S::MT_LIN; ++mt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int r = (mg * S::MT_LIN + mt) * 16 + g + 8 * half, col = n0 + 2 * t;
            float2* gb = reinterpret_cast<float2*>(GB + (size_t)r * LDD + col);
            const float2 d = *gb;
            const float2 v = make_float2(acc[mt][0][2 * half] + d.x, acc[mt][0][2 * half + 1] + d.y);
            if (k > 0) *gb = v;
            else *reinterpret_cast<float2*>(P.d_x0 + (grow + r) * D + col) = v;
          }
      }
    }
    __syncthreads();
  }
}

}  // namespace pit
