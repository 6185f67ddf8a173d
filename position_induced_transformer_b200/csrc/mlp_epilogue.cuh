// K4: epilogues of the kaiming_mlp Linears (pit.py:21-26) and of the GELUs around them (pit.py:111, 121).
//
//   bias_act_fwd_kernel   out = act(z + bias)                               one 128-bit pass
//   bias_act_bwd_kernel   d_z = d_out * act'(z + bias),  d_bias = colsum(d_z)  one 128-bit pass; a thread keeps its
//                         column group for the whole grid-stride walk, so the column sums are register partials,
//                         combined per CTA in shared memory and finished with one RED per column and CTA
// Both are HBM/L2-bound elementwise kernels (the latent activations are 0.5-10 MB); the point is the launch count:
// autograd runs GELU-backward and the bias reduction as two kernels, the latter at 11 us for a 2048 x 64 matrix.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pit {

constexpr int EPI_THREADS = 256;

// torch's exact GELU: x * 0.5 * (1 + erf(x / sqrt(2))) and its derivative
__device__ __forceinline__ float epi_gelu(float x) { return x * 0.5f * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float epi_gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = expf(-0.5f * x * x) * 0.3989422804014327f;
  return fmaf(x, pdf, cdf);
}

struct EpiParams {
  const float* z;
  const float* bias;
  const float* d_out;
  float* out;     // forward: out; backward: d_z
  float* d_bias;  // zero-initialised
  int64_t rows;
  int cols4;  // cols / 4; divides EPI_THREADS
  int gelu;
};

template <bool GELU>
__global__ void __launch_bounds__(EPI_THREADS) bias_act_fwd_kernel(const EpiParams P) {
  const int c4 = threadIdx.x % P.cols4;
  const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias) + c4);
  const int64_t total = P.rows * P.cols4;
  for (int64_t i = (int64_t)blockIdx.x * EPI_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EPI_THREADS) {
    float4 v = __ldg(reinterpret_cast<const float4*>(P.z) + i);
    v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
    if (GELU) v = make_float4(epi_gelu(v.x), epi_gelu(v.y), epi_gelu(v.z), epi_gelu(v.w));
    reinterpret_cast<float4*>(P.out)[i] = v;
  }
}

template <bool GELU>
__global__ void __launch_bounds__(EPI_THREADS) bias_act_bwd_kernel(const EpiParams P) {
  __shared__ float4 part[EPI_THREADS];
  const int c4 = threadIdx.x % P.cols4;
  const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias) + c4);
  const int64_t total = P.rows * P.cols4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // the grid stride is a multiple of cols4, so a thread stays on its column group
  for (int64_t i = (int64_t)blockIdx.x * EPI_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EPI_THREADS) {
    float4 g = __ldg(reinterpret_cast<const float4*>(P.d_out) + i);
    if (GELU) {
      float4 v = __ldg(reinterpret_cast<const float4*>(P.z) + i);
      v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
      g.x *= epi_gelu_grad(v.x), g.y *= epi_gelu_grad(v.y), g.z *= epi_gelu_grad(v.z), g.w *= epi_gelu_grad(v.w);
    }
    reinterpret_cast<float4*>(P.out)[i] = g;
    acc.x += g.x, acc.y += g.y, acc.z += g.z, acc.w += g.w;
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < P.cols4) {
    float4 s = part[threadIdx.x];
    for (int t = threadIdx.x + P.cols4; t < EPI_THREADS; t += P.cols4) {
      const float4 o = part[t];
      s.x += o.x, s.y += o.y, s.z += o.z, s.w += o.w;
    }
    atomicAdd(reinterpret_cast<float4*>(P.d_bias) + threadIdx.x, s);
  }
}

}  // namespace pit
