// Stand-alone probe of tcgen05.mma kind::f16 (BF16 operands, FP32 accumulate): A K-major SWIZZLE_128B,
// B MN-major SWIZZLE_128B (16-bit canonical atoms: 8 K-rows x 128 bytes = 64 columns).  One CTA, M=128, N=64, K=16*ksteps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/probe/umma_probe_bf16 scripts/probe/umma_probe_bf16.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../position_induced_transformer_b200/csrc/dense_attention.cuh"
using namespace pit;

struct ProbeArgs { int mode; int ksteps; float* d_out; const float* a; const float* b; };

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(160, 1) probe_kernel(ProbeArgs P) {
  extern __shared__ unsigned char raw_smem[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t raw = smem_u32(raw_smem);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* tp = raw_smem + (tiles - raw);
  unsigned char* a_tile = tp;          // 128 rows x 128 B (64 bf16 along K) = 16 KB
  unsigned char* b_tile = tp + 16384;  // 64 k x 64 n bf16 = 8 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16384 + 8192) / 4; i += blockDim.x) reinterpret_cast<float*>(tp)[i] = 0.f;
  if (tid == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc<64>(&tmem_base_smem);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // A: K-major SW128: element (r, k): chunk = k / 8 (16 B = 8 bf16)
  if (tid < 128) {
    const int r = tid;
    for (int k = 0; k < 64; ++k) {
      const uint32_t off = a_chunk_offset(r, k >> 3) + (k & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(a_tile + off) = __float2bfloat16(P.a[r * 64 + k]);
    }
  }
  // B: MN-major SW128 (16-bit): atom = 8 k-rows x 128 B (64 n). element (k, n): atom (n/64, k/8); row k%8; chunk (n%64)/8 ^ (k%8)
  for (int idx = tid; idx < 64 * 64; idx += blockDim.x) {
    const int k = idx / 64, n = idx % 64;
    uint32_t off;
    if (P.mode == 0) off = (uint32_t)((k >> 3) * 1024 + (k & 7) * 128 + ((((n & 63) >> 3) ^ (k & 7)) << 4) + (n & 7) * 2);  // single n-atom (N=64)
    else off = a_chunk_offset(n, k >> 3) + (k & 7) * 2;  // K-major B (rows = n)
    *reinterpret_cast<__nv_bfloat16*>(b_tile + off) = __float2bfloat16(P.b[idx]);
  }
  fence_async_shared();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    if (lane == 0) {
      // idesc: c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), b_major bit16, N>>3 <<17, M>>4 <<24
      const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      for (int kg = 0; kg < P.ksteps; ++kg) {  // K = 16 per MMA = 32 bytes of an A row = two 8-row k-groups of B
        uint64_t ad = umma_desc(tiles + kg * 32, 16, 1024), bd; uint32_t idesc;
        if (P.mode == 0) { bd = umma_desc(tiles + 16384 + kg * 2048, 8192 /*LBO: next 64 columns (unused, N=64)*/, 1024 /*SBO: next 8 k rows*/); idesc = idesc_base | (1u << 16); }
        else { bd = umma_desc(tiles + 16384 + kg * 32, 16, 1024); idesc = idesc_base; }
        umma_bf16(tmem_base, ad, bd, idesc, kg > 0 ? 1u : 0u);
      }
      umma_commit(&done_bar);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int e = 0; e < 32; ++e) P.d_out[tid * 64 + c0 + e] = v[e];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<64>(tmem_base); }
}

int main() {
  std::vector<float> a(128 * 64), b(64 * 64), d(128 * 64);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < 64; ++k) a[r * 64 + k] = float((r * 7 + k * 3) % 5 - 2);
  for (int k = 0; k < 64; ++k) for (int n = 0; n < 64; ++n) b[k * 64 + n] = float((k * 5 + n * 11) % 7 - 3);
  float *da, *db, *dd;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode) for (int ksteps = 1; ksteps <= 4; ksteps += 3) {
    cudaMemset(dd, 0xff, d.size() * 4);
    ProbeArgs P{mode, ksteps, dd, da, db};
    probe_kernel<<<1, 160, 64 * 1024>>>(P);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    float ref0[4];
    for (int r = 0; r < 128; ++r) for (int n = 0; n < 64; ++n) {
      float acc = 0; for (int k = 0; k < 16 * ksteps; ++k) acc += a[r * 64 + k] * b[k * 64 + n];
      if (r == 0 && n < 4) ref0[n] = acc;
      if (d[r * 64 + n] != acc) ++bad;
    }
    printf("bf16 mode %d (%s B) ksteps %d: mismatches %d / 8192 | d[0][0..3] = %g %g %g %g | ref = %g %g %g %g\n", mode, mode == 0 ? "MN-major" : "K-major", ksteps, bad,
           d[0], d[1], d[2], d[3], ref0[0], ref0[1], ref0[2], ref0[3]);
  }
  return 0;
}
