// K7: the gradient all-reduce of a data-parallel step fused with the Adam update, over NVLink peer memory.
//
// PiT models are small (8.6 k - 1.27 M parameters), so the one collective of a training step -- the SUM of the flat
// gradient over the ranks (utils.py:98 sums the loss over the batch) -- is latency, not bandwidth: NCCL's all-reduce plus
// the pack copy in front of it and the optimizer launch behind it cost ~60 us of a 0.9 ms step at 8 GPUs.  Here ONE kernel
// per rank does all of it over symmetric memory (every rank maps every other rank's bucket through NVSwitch):
//
//   1. gather   this rank's gradients (wherever autograd left them) -> its own bucket[parity] in symmetric memory
//   2. signal   when every CTA of the rank has written, a system-scope release store of (step + 1) into the flag word
//               `rank` of EVERY peer's flag block; then each CTA spins (system-scope acquire) until all `world` words of
//               the local flag block show step + 1
//   3. reduce + update   thread i reads element i of all `world` buckets (coalesced 128-byte peer reads), sums them in rank
//               order (every rank gets the same bits), and applies Adam to its slice of the flat parameter / moment buffers
//
// Buckets are double-buffered by the parity of the step counter, which lives on the device so that a replayed CUDA graph
// alternates them by itself: a rank can overwrite bucket[p] only two steps later, after it has seen every peer's flag of the
// step in between -- which a peer publishes only after it has finished reading.  No trailing barrier, no NCCL call.
//
// Adam as torch.optim.Adam (no weight decay, no amsgrad -- what every reference script uses, train_darcy.py:115):
//   m = lerp(m, g, 1-b1);  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) m / (sqrt(v) / sqrt(1-b2^t) + eps)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pit {

constexpr int ARA_THREADS = 256;
constexpr int ARA_MAX_TENSORS = 64;
constexpr int ARA_MAX_WORLD = 16;
constexpr int ARA_FLAG_WORDS = 64;  // flag block at the start of every rank's symmetric region (256 bytes)

struct AllReduceAdamParams {
  int world, rank, n_tensors;
  int64_t total;                            // elements of the flat buffers: sum of the tensor sizes, each rounded up to a multiple of 4
  const float* grad[ARA_MAX_TENSORS];       // this rank's gradients, in parameter order
  int32_t numel[ARA_MAX_TENSORS];
  float* region[ARA_MAX_WORLD];             // symmetric regions of all ranks: [flags: 64 x u32][bucket 0: stride][bucket 1: stride]
  int64_t stride;                           // elements per bucket (>= total, multiple of 4)
  float* param;                             // flat parameters / moments of this rank
  float* exp_avg;
  float* exp_avg_sq;
  int32_t* step;                            // device step counter t (number of updates done so far)
  uint32_t* arrive;                         // [3] device words: CTAs that have written (monotone), CTAs that have finished (monotone), error flag
  const float* lr;                          // device scalar
  float beta1, beta2, eps;
};

__device__ __forceinline__ void ara_store_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ara_load_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t ara_add_acq_rel(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint64_t ara_globaltimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t ARA_TIMEOUT_NS = 2000000000ull;  // 2 s

// Gradient group `q` (four consecutive elements of the flat index space; a group never straddles two tensors because every
// tensor starts on a multiple of four) read from wherever autograd left the tensor.
__device__ __forceinline__ float4 ara_load_grad(const AllReduceAdamParams& P, const int64_t* offs, int64_t q) {
  const int64_t i = 4 * q;
  int k = 0;
  while (i >= offs[k + 1]) ++k;
  const float* src = P.grad[k];
  const int64_t e = i - offs[k], n = P.numel[k];
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!src) return g;
  if (e + 4 <= n && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) return *reinterpret_cast<const float4*>(src + e);
  if (e < n) g.x = src[e];
  if (e + 1 < n) g.y = src[e + 1];
  if (e + 2 < n) g.z = src[e + 2];
  if (e + 3 < n) g.w = src[e + 3];
  return g;
}

// Every CTA of the launch must be resident at once (they wait for each other): two CTAs per SM by the launch bounds, and the host
// sizes the grid from cudaOccupancyMaxActiveBlocksPerMultiprocessor.  (A first version assumed two per SM while ptxas had given the
// kernel 158 registers -- one CTA per SM -- and deadlocked into its 2 s timeout on every step once the model needed more than 148 CTAs.)
template <bool MULTI>   // MULTI: world > 1
__global__ void __launch_bounds__(ARA_THREADS, 2) allreduce_adam_kernel(const AllReduceAdamParams P) {
  __shared__ int64_t offs[ARA_MAX_TENSORS + 1];
  const int tid = threadIdx.x;
  const int t = *reinterpret_cast<volatile int32_t*>(P.step);   // updates done so far; every CTA reads it before anyone bumps it
  const int parity = t & 1;
  const uint32_t epoch = (uint32_t)t + 1u;
  const int64_t bucket = ARA_FLAG_WORDS + (int64_t)parity * P.stride;
  if (tid == 0) {
    int64_t o = 0;
    for (int k = 0; k < P.n_tensors; ++k) {
      offs[k] = o;
      o += (P.numel[k] + 3) & ~3;       // every tensor starts on a 16-byte boundary of the flat buffers
    }
    offs[P.n_tensors] = o;
  }
  __syncthreads();
  const int64_t gtid = (int64_t)blockIdx.x * ARA_THREADS + tid, gsize = (int64_t)gridDim.x * ARA_THREADS;
  const int64_t groups = P.total / 4;
  if (MULTI) {
    // 1. gather into this rank's bucket (the same thread -> group mapping as the reduction below)
    float4* mine = reinterpret_cast<float4*>(P.region[P.rank] + bucket);
    for (int64_t q = gtid; q < groups; q += gsize) mine[q] = ara_load_grad(P, offs, q);
    // 2. every CTA of this rank has written -> publish; then wait for every rank
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const uint32_t target = epoch * gridDim.x;
      const uint32_t seen = ara_add_acq_rel(P.arrive, 1u) + 1u;      // acq_rel: the last CTA has observed every CTA's writes
      if (seen == target) {
        __threadfence_system();
        for (int r = 0; r < P.world; ++r) ara_store_release_sys(reinterpret_cast<uint32_t*>(P.region[r]) + P.rank, epoch);
      }
      const uint32_t* flags = reinterpret_cast<const uint32_t*>(P.region[P.rank]);
      const uint64_t t0 = ara_globaltimer();
      for (int r = 0; r < P.world; ++r)
        while ((int32_t)(ara_load_acquire_sys(flags + r) - epoch) < 0) {
          if (ara_globaltimer() - t0 > ARA_TIMEOUT_NS) {      // a peer is gone: never hang the GPU -- flag the error and carry on
            atomicExch(P.arrive + 2, 1u);
            break;
          }
        }
    }
    __syncthreads();
  }
  // 3. reduce + Adam, four elements per thread and trip
  const float lr = *P.lr;
  const float tf = (float)(t + 1);
  const float bc1 = 1.f - powf(P.beta1, tf), bc2 = 1.f - powf(P.beta2, tf);
  const float step_size = lr / bc1, inv_bc2_sqrt = rsqrtf(bc2);
  for (int64_t q = gtid; q < groups; q += gsize) {
    float4 g;
    if (MULTI) {
      float4 part[ARA_MAX_WORLD];
#pragma unroll
      for (int r = 0; r < ARA_MAX_WORLD; ++r)      // every peer read in flight before the first add
        if (r < P.world) part[r] = __ldcg(reinterpret_cast<const float4*>(P.region[r] + bucket) + q);
      g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < ARA_MAX_WORLD; ++r)      // rank order: every rank computes the same bits
        if (r < P.world) g.x += part[r].x, g.y += part[r].y, g.z += part[r].z, g.w += part[r].w;
    } else {
      g = ara_load_grad(P, offs, q);
    }
    float4 m = reinterpret_cast<float4*>(P.exp_avg)[q], v = reinterpret_cast<float4*>(P.exp_avg_sq)[q], p = reinterpret_cast<float4*>(P.param)[q];
    const float gg[4] = {g.x, g.y, g.z, g.w};
    float mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w}, pp[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = fmaf(1.f - P.beta1, gg[e] - mm[e], mm[e]);
      vv[e] = fmaf(1.f - P.beta2, gg[e] * gg[e], P.beta2 * vv[e]);
      pp[e] -= step_size * mm[e] / (sqrtf(vv[e]) * inv_bc2_sqrt + P.eps);
    }
    reinterpret_cast<float4*>(P.exp_avg)[q] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(P.exp_avg_sq)[q] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(P.param)[q] = make_float4(pp[0], pp[1], pp[2], pp[3]);
  }
  // the last CTA to finish bumps the step counter (everyone read it at the top)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(P.arrive + 1, 1u) + 1u;
    if (done == epoch * gridDim.x) *reinterpret_cast<volatile int32_t*>(P.step) = t + 1;
  }
}

}  // namespace pit
