// K5: relative Lp error of the reference's utils.py:60-98 (RelLpNorm, p = 1 or 2):
//   loss = sum_b mean_o ||true[b,:,o] - pred[b,:,o]||_p / ||true[b,:,o]||_p
// torch runs it as ~8 launches forward (two of them 8-row norm reductions of 9 us each at Darcy-421) and ~6 backward;
// here: one partial-sum pass + a one-block finalize forward, one elementwise pass backward.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "geometry.cuh"

namespace pit {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_OUT = 4;

struct LossParams {
  const float* truth;  // [B, L, O]
  const float* pred;   // [B, L, O]
  float* sums;         // [B, O, 2]: sum |e|^p, sum |t|^p (zero-initialised); finalize turns them into the two norms
  float* loss;         // scalar
  const float* d_loss; // scalar (backward)
  float* d_pred;       // [B, L, O]
  int64_t L;
  int B, O, p;
};

template <int P_ORD>
__global__ void __launch_bounds__(LOSS_THREADS) rel_lp_partial_kernel(const LossParams P) {
  __shared__ float red[LOSS_THREADS / 32][LOSS_MAX_OUT][2];
  const int b = blockIdx.y;
  const int64_t n = P.L * P.O;
  const float* t = P.truth + (int64_t)b * n;
  const float* q = P.pred + (int64_t)b * n;
  float se[LOSS_MAX_OUT] = {0.f, 0.f, 0.f, 0.f}, st[LOSS_MAX_OUT] = {0.f, 0.f, 0.f, 0.f};
  // the grid stride is a multiple of O (the host rounds it), so a thread stays on one output variable
  const int64_t stride = (int64_t)gridDim.x * LOSS_THREADS;
  const int64_t i0 = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  const int o = (int)(i0 % P.O);
  float ae = 0.f, at = 0.f;
  for (int64_t i = i0; i < n; i += stride) {
    const float tv = __ldg(t + i), e = tv - __ldg(q + i);
    ae += P_ORD == 2 ? e * e : fabsf(e);
    at += P_ORD == 2 ? tv * tv : fabsf(tv);
  }
#pragma unroll
  for (int k = 0; k < LOSS_MAX_OUT; ++k) {
    se[k] = warp_sum(k == o ? ae : 0.f);
    st[k] = warp_sum(k == o ? at : 0.f);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < LOSS_MAX_OUT; ++k) red[warp][k][0] = se[k], red[warp][k][1] = st[k];
  __syncthreads();
  if (threadIdx.x < 2 * P.O) {
    const int k = threadIdx.x >> 1, w2 = threadIdx.x & 1;
    float acc = 0.f;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) acc += red[w][k][w2];
    atomicAdd(P.sums + ((int64_t)b * P.O + k) * 2 + w2, acc);
  }
}

template <int P_ORD>
__global__ void rel_lp_finalize_kernel(const LossParams P) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < P.B * P.O; i += 32) {
    float ne = P.sums[2 * i], nt = P.sums[2 * i + 1];
    if (P_ORD == 2) ne = sqrtf(ne), nt = sqrtf(nt);
    P.sums[2 * i] = ne;  // keep the norms for the backward pass
    P.sums[2 * i + 1] = nt;
    acc += ne / nt;
  }
  acc = warp_sum(acc);
  if (threadIdx.x == 0) *P.loss = acc / P.O;
}

template <int P_ORD>
__global__ void __launch_bounds__(LOSS_THREADS) rel_lp_backward_kernel(const LossParams P) {
  const int b = blockIdx.y;
  const int64_t n = P.L * P.O;
  const float* t = P.truth + (int64_t)b * n;
  const float* q = P.pred + (int64_t)b * n;
  float* dq = P.d_pred + (int64_t)b * n;
  const int64_t stride = (int64_t)gridDim.x * LOSS_THREADS;
  const int64_t i0 = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  const int o = (int)(i0 % P.O);
  const float ne = P.sums[((int64_t)b * P.O + o) * 2], nt = P.sums[((int64_t)b * P.O + o) * 2 + 1];
  // d/d pred of ||t - q||_p / ||t||_p / O:  p = 2: -(t - q) / (||e|| ||t|| O),  p = 1: -sign(t - q) / (||t|| O); zero where ||e|| = 0
  const float g = __ldg(P.d_loss) / (nt * P.O);
  const float c = P_ORD == 2 ? (ne > 0.f ? g / ne : 0.f) : g;
  for (int64_t i = i0; i < n; i += stride) {
    const float e = __ldg(t + i) - __ldg(q + i);
    dq[i] = P_ORD == 2 ? -c * e : (e > 0.f ? -c : (e < 0.f ? c : 0.f));
  }
}

}  // namespace pit
