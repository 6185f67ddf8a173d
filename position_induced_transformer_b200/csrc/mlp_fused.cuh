// K8: a whole kaiming_mlp (pit.py:13-26) with a narrow input -- the encoder lift `en_layer` of the shared-mesh models
// (pit.py:110-111: Linear(H (in_dim + space_dim) -> hid) -> GELU -> Linear(hid -> hid), then the caller's GELU) -- in ONE launch
// forward and ONE launch backward.
//
// On the latent grid the lift is a [B 256 x 4..6] x [6 x 64] product followed by a [2048 x 64] x [64 x 64] one: the cuBLAS path
// spends two GEMM launches and two epilogue launches forward and four GEMMs (one of them 23 us for a K = 6 operand) plus two
// epilogues backward, ~65 us of a Darcy-421 step and a quarter of a Sod step, on a few MFLOP.  Here a CTA owns 32 rows:
//   forward    Z1 = X W1^T + b1 on the fp32 pipe (K <= 32: too narrow for an MMA k-step to pay), H1 = gelu(Z1) in shared memory,
//              Z2 = H1 W2^T + b2 on mma.sync (3xTF32 or TF32 by torch's matmul precision, as processor_block.cuh),
//              out = gelu(Z2) or Z2; Z1 and Z2 are saved
//   backward   dZ2 = G gelu'(Z2); db2, dW2 += dZ2^T H1 (mma); dH1 = dZ2 W2 (mma); dZ1 = dH1 gelu'(Z1); db1;
//              dW1 += dZ1^T X and dX = dZ1 W1 on the fp32 pipe; parameter gradients leave as REDs into zeroed buffers.
#pragma once
#include "processor_block.cuh"

namespace pit {

constexpr int MF_ROWS = 32;        // rows per CTA
constexpr int MF_MAX_IN = 32;      // widest input handled on the fp32 pipe

struct MlpFusedParams {
  const float* x;    // [R, K]
  const float* w1;   // [D, K]
  const float* b1;   // [D]
  const float* w2;   // [D, D]
  const float* b2;   // [D]
  float* z1;         // [R, D] saved pre-activations
  float* z2;         // [R, D]
  float* out;        // [R, D]
  int64_t R;
  int K, act_out;
  // backward
  const float* d_out;  // [R, D]
  float* d_x;          // [R, K]
  float* d_w1;         // zero-initialised [D, K]
  float* d_b1;         // [D]
  float* d_w2;         // [D, D]
  float* d_b2;         // [D]
};

template <int D, bool LIN3>
__global__ void __launch_bounds__(PB_THREADS) mlp_fused_fwd_kernel(const MlpFusedParams P) {
  constexpr int LDD = D + 4, MT = MF_ROWS / 16;
  __shared__ __align__(16) float XS[MF_ROWS * MF_MAX_IN];
  __shared__ __align__(16) float W1S[D * MF_MAX_IN];
  __shared__ __align__(16) float HS[MF_ROWS * LDD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * MF_ROWS;
  const int K = P.K;
  for (int i = tid; i < MF_ROWS * K; i += PB_THREADS) {
    const int r = i / K;
    XS[i] = row0 + r < P.R ? __ldg(P.x + row0 * K + i) : 0.f;
  }
  for (int i = tid; i < D * K; i += PB_THREADS) W1S[i] = __ldg(P.w1 + i);
  __syncthreads();
  // Z1 on the fp32 pipe: a thread takes one row and D / 8 consecutive columns
  {
    constexpr int CPT = D / 8;
    const int r = tid >> 3, c0 = (tid & 7) * CPT;
    float z[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) z[c] = __ldg(P.b1 + c0 + c);
    for (int k = 0; k < K; ++k) {
      const float xv = XS[r * K + k];
#pragma unroll
      for (int c = 0; c < CPT; ++c) z[c] = fmaf(xv, W1S[(c0 + c) * K + k], z[c]);
    }
    const bool ok = row0 + r < P.R;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (ok) P.z1[(row0 + r) * D + c0 + c] = z[c];
      HS[r * LDD + c0 + c] = tm_gelu(z[c]);
    }
  }
  __syncthreads();
  // Z2 = H1 W2^T + b2 on mma.sync
  for (int item = warp; item < D / 8; item += PB_WARPS) {
    const int n0 = item * 8;
    float acc[MT][1][4];
    pb_zero(acc);
    pb_gemm<MT, 1, LIN3, 1, 1, true>(acc, D / 8, pb_a_rows<MT>(HS, LDD), pb_b_cols<1>(P.w2 + (int64_t)n0 * D, D));
    const float2 bias = __ldg(reinterpret_cast<const float2*>(P.b2 + n0 + 2 * t));
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int64_t r = row0 + mt * 16 + g + 8 * half;
        if (r >= P.R) continue;
        const int col = n0 + 2 * t;
        const float2 z = make_float2(acc[mt][0][2 * half] + bias.x, acc[mt][0][2 * half + 1] + bias.y);
        *reinterpret_cast<float2*>(P.z2 + r * D + col) = z;
        *reinterpret_cast<float2*>(P.out + r * D + col) = P.act_out ? make_float2(tm_gelu(z.x), tm_gelu(z.y)) : z;
      }
  }
}

template <int D, bool LIN3>
__global__ void __launch_bounds__(PB_THREADS) mlp_fused_bwd_kernel(const MlpFusedParams P) {
  constexpr int LDD = D + 4, MT = MF_ROWS / 16;
  constexpr int MT_W = D >= 64 ? 2 : 1, NT_W = D >= 64 ? 2 : 1;
  __shared__ __align__(16) float XS[MF_ROWS * MF_MAX_IN];
  __shared__ __align__(16) float W1S[D * MF_MAX_IN];
  __shared__ __align__(16) float AB[MF_ROWS * LDD];   // dZ2, later dZ1
  __shared__ __align__(16) float H1[MF_ROWS * LDD];
  __shared__ float RED[PB_THREADS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * MF_ROWS;
  const int K = P.K;
  for (int i = tid; i < MF_ROWS * K; i += PB_THREADS) {
    const int r = i / K;
    XS[i] = row0 + r < P.R ? __ldg(P.x + row0 * K + i) : 0.f;
  }
  for (int i = tid; i < D * K; i += PB_THREADS) W1S[i] = __ldg(P.w1 + i);
  // dZ2 = G act'(Z2), H1 = gelu(Z1), db2 (a thread stays on one column; rows past the end contribute zeros)
  {
    const int c = tid % D;
    float colsum = 0.f;
    for (int r = tid / D; r < MF_ROWS; r += PB_THREADS / D) {
      float dz2 = 0.f, h1 = 0.f;
      if (row0 + r < P.R) {
        const float gq = __ldg(P.d_out + (row0 + r) * D + c);
        float act, dact = 1.f;
        if (P.act_out) tm_gelu_pair(__ldg(P.z2 + (row0 + r) * D + c), act, dact);
        dz2 = gq * dact;
        h1 = tm_gelu(__ldg(P.z1 + (row0 + r) * D + c));
      }
      AB[r * LDD + c] = dz2;
      H1[r * LDD + c] = h1;
      colsum += dz2;
    }
    RED[tid] = colsum;
  }
  __syncthreads();
  if (tid < D) {
    float s = 0.f;
    for (int q = tid; q < PB_THREADS; q += D) s += RED[q];
    atomicAdd(P.d_b2 + tid, s);
  }
  // dW2 += dZ2^T H1
  for (int item = warp; item < (D / 16 / MT_W) * (D / 8 / NT_W); item += PB_WARPS) {
    const int mg = item / (D / 8 / NT_W), ng = item - mg * (D / 8 / NT_W);
    const int m0 = mg * MT_W * 16, n0 = ng * NT_W * 8;
    float acc[MT_W][NT_W][4];
    pb_zero(acc);
    pb_gemm<MT_W, NT_W, LIN3, LDD, LDD>(acc, MF_ROWS / 8, pb_a_cols<MT_W>(AB + m0, LDD), pb_b_rows<NT_W>(H1 + n0, LDD));
#pragma unroll
    for (int mt = 0; mt < MT_W; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT_W; ++nt)
#pragma unroll
        for (int half = 0; half < 2; ++half)
          atomicAdd(reinterpret_cast<float2*>(P.d_w2 + (int64_t)(m0 + mt * 16 + g + 8 * half) * D + n0 + nt * 8 + 2 * t),
                    make_float2(acc[mt][nt][2 * half], acc[mt][nt][2 * half + 1]));
  }
  // dH1 = dZ2 W2 (one n-tile per warp, kept in registers until every warp has read dZ2), then dZ1 = dH1 gelu'(Z1) over dZ2's buffer
  static_assert(D / 8 <= PB_WARPS, "one n-tile of the hidden width per warp");
  {
    float acc[MT][1][4];
    pb_zero(acc);
    const int n0 = warp * 8;
    if (warp < D / 8) pb_gemm<MT, 1, LIN3, 1, D, true>(acc, D / 8, pb_a_rows<MT>(AB, LDD), pb_b_rows<1>(P.w2 + n0, D));
    __syncthreads();
    if (warp < D / 8) {
      const int col = n0 + 2 * t;
      float cs0 = 0.f, cs1 = 0.f;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = mt * 16 + g + 8 * half;
          float2 v = make_float2(0.f, 0.f);
          if (row0 + r < P.R) {
            const float2 z1 = __ldg(reinterpret_cast<const float2*>(P.z1 + (row0 + r) * D + col));
            float a0, d0, a1, d1;
            tm_gelu_pair(z1.x, a0, d0);
            tm_gelu_pair(z1.y, a1, d1);
            v = make_float2(acc[mt][0][2 * half] * d0, acc[mt][0][2 * half + 1] * d1);
          }
          *reinterpret_cast<float2*>(AB + r * LDD + col) = v;
          cs0 += v.x, cs1 += v.y;
        }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        cs0 += __shfl_xor_sync(FULL, cs0, o);
        cs1 += __shfl_xor_sync(FULL, cs1, o);
      }
      if (g == 0) {
        atomicAdd(P.d_b1 + col, cs0);
        atomicAdd(P.d_b1 + col + 1, cs1);
      }
    }
  }
  __syncthreads();
  // dW1[c][k] += sum_r dZ1[r][c] X[r][k]   and   dX[r][k] = sum_c dZ1[r][c] W1[c][k], both on the fp32 pipe
  for (int o = tid; o < D * K; o += PB_THREADS) {
    const int c = o / K, k = o - c * K;
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < MF_ROWS; ++r) s = fmaf(AB[r * LDD + c], XS[r * K + k], s);
    atomicAdd(P.d_w1 + o, s);
  }
  if (P.d_x) {
    for (int o = tid; o < MF_ROWS * K; o += PB_THREADS) {
      const int r = o / K, k = o - r * K;
      if (row0 + r >= P.R) continue;
      float s = 0.f;
#pragma unroll 8
      for (int c = 0; c < D; ++c) s = fmaf(AB[r * LDD + c], W1S[c * K + k], s);
      P.d_x[(row0 + r) * K + k] = s;
    }
  }
}

}  // namespace pit
