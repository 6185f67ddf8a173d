// Stand-alone probe of tcgen05.mma kind::tf32 operand layouts (one CTA, one MMA M=128 N=64 K=8..32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe/umma_probe scripts/probe/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../position_induced_transformer_b200/csrc/dense_attention.cuh"
using namespace pit;

struct ProbeArgs {
  int mode;       // layout hypothesis
  int ksteps;     // number of K=8 MMAs (1..4)
  float* d_out;   // [128][64]
  const float* a; // [128][32] row-major (r,k)
  const float* b; // [32][64]  row-major (k,n)
};

__global__ void __launch_bounds__(160, 1) probe_kernel(ProbeArgs P) {
  extern __shared__ unsigned char raw_smem[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t raw = smem_u32(raw_smem);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  unsigned char* tp = raw_smem + (tiles - raw);
  unsigned char* a_tile = tp;            // 16 KB
  unsigned char* b_tile = tp + 16384;    // 8 KB (32 x 64 x 4)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16384 + 8192) / 4; i += blockDim.x) reinterpret_cast<float*>(tp)[i] = 0.f;
  if (tid == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
  if (warp == 4) tmem_alloc<64>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // ---- fill A: K-major. modes 0-2: SW128 ; mode 3: no swizzle (core matrices 8 rows x 16 bytes)
  if (tid < 128) {
    const int r = tid;
    for (int k = 0; k < 32; ++k) {
      const float v = P.a[r * 32 + k];
      uint32_t off;
      if (P.mode != 3) off = a_chunk_offset(r, k >> 2) + (k & 3) * 4;
      else off = (uint32_t)((r >> 3) * 128 + (r & 7) * 16 + (k >> 2) * 2048 + (k & 3) * 4);  // SBO=128 (next 8 rows), LBO=2048 (next 16B along K)
      *reinterpret_cast<float*>(a_tile + off) = v;
    }
  }
  // ---- fill B
  for (int idx = tid; idx < 32 * 64; idx += blockDim.x) {
    const int k = idx / 64, n = idx % 64;
    const float v = P.b[idx];
    uint32_t off;
    if (P.mode == 0 || P.mode == 1) off = (uint32_t)((((n >> 5) << 2) + (k >> 3)) * 1024 + (k & 7) * 128 + ((((n >> 2) & 7) ^ (k & 7)) << 4)) + (n & 3) * 4;  // MN-major SW128 atoms
    else if (P.mode == 2) off = a_chunk_offset(n, k >> 2) + (k & 3) * 4;                           // K-major SW128, rows = n
    else if (P.mode == 4) off = (uint32_t)((n >> 2) * 128 + (k >> 3) * 2048 + (k & 7) * 16 + (n & 3) * 4);   // MN-major, no swizzle: core = 8 k x 16 B; SBO=128 (next 4 n), LBO=2048 (next 8 k)
    else if (P.mode == 5 || P.mode == 6) off = (uint32_t)((n >> 5) * 4096 + (k >> 2) * 512 + (k & 3) * 128 + ((((n & 31) >> 3) ^ (k & 3)) << 5) + (n & 7) * 4);  // MN-major SW128_BASE32B: atom 4 k x 128 B
    else if (P.mode == 7) off = (uint32_t)((n >> 5) * 4096 + (k >> 3) * 1024 + (k & 7) * 128 + ((((n & 31) >> 3) ^ (k & 3)) << 5) + (n & 7) * 4);  // BASE32B with 8-row atoms
    else off = (uint32_t)((n >> 3) * 128 + (n & 7) * 16 + (k >> 2) * 1024 + (k & 3) * 4);          // K-major no swizzle: SBO=128, LBO=1024
    *reinterpret_cast<float*>(b_tile + off) = v;
  }
  fence_async_shared();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    if (lane == 0) {
      const uint32_t a_base = tiles, b_base = tiles + 16384;
      for (int kg = 0; kg < P.ksteps; ++kg) {
        uint64_t ad, bd; uint32_t idesc;
        if (P.mode == 0) { ad = umma_desc(a_base + kg * 32, 16, 1024); bd = umma_desc(b_base + kg * 1024, 4096, 1024); idesc = umma_idesc_tf32(64) | (1u << 16); }
        else if (P.mode == 1) { ad = umma_desc(a_base + kg * 32, 16, 1024); bd = umma_desc(b_base + kg * 1024, 1024, 4096); idesc = umma_idesc_tf32(64) | (1u << 16); }
        else if (P.mode == 2) { ad = umma_desc(a_base + kg * 32, 16, 1024); bd = umma_desc(b_base + kg * 32, 16, 1024); idesc = umma_idesc_tf32(64); }
        else if (P.mode == 4) { ad = umma_desc(a_base + kg * 32, 16, 1024); bd = umma_desc(b_base + kg * 2048, 2048, 128) & ~((uint64_t)7 << 61); idesc = umma_idesc_tf32(64) | (1u << 16); }
        else if (P.mode == 5) { ad = umma_desc(a_base + kg * 32, 16, 1024); bd = (umma_desc(b_base + kg * 1024, 4096, 512) & ~((uint64_t)7 << 61)) | ((uint64_t)1 << 61); idesc = umma_idesc_tf32(64) | (1u << 16); }
        else if (P.mode == 6) { ad = umma_desc(a_base + kg * 32, 16, 1024); bd = (umma_desc(b_base + kg * 1024, 512, 4096) & ~((uint64_t)7 << 61)) | ((uint64_t)1 << 61); idesc = umma_idesc_tf32(64) | (1u << 16); }
        else if (P.mode == 7) { ad = umma_desc(a_base + kg * 32, 16, 1024); bd = (umma_desc(b_base + kg * 1024, 4096, 1024) & ~((uint64_t)7 << 61)) | ((uint64_t)1 << 61); idesc = umma_idesc_tf32(64) | (1u << 16); }
        else {
          ad = umma_desc(a_base + kg * 4096, 2048, 128) & ~((uint64_t)7 << 61);
          bd = umma_desc(b_base + kg * 2048, 1024, 128) & ~((uint64_t)7 << 61);
          idesc = umma_idesc_tf32(64);
        }
        umma_tf32(tmem_base, ad, bd, idesc, kg > 0 ? 1u : 0u);
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
  std::vector<float> a(128 * 32), b(32 * 64), d(128 * 64), ref(128 * 64);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < 32; ++k) a[r * 32 + k] = float((r * 7 + k * 3) % 5 - 2);
  for (int k = 0; k < 32; ++k) for (int n = 0; n < 64; ++n) b[k * 64 + n] = float((k * 5 + n * 11) % 7 - 3);
  float *da, *db, *dd;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dd, d.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 2; mode < 8; ++mode) {
    for (int ksteps = 1; ksteps <= 4; ksteps += 3) {
      cudaMemset(dd, 0xff, d.size() * 4);
      ProbeArgs P{mode, ksteps, dd, da, db};
      probe_kernel<<<1, 160, 64 * 1024>>>(P);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d ksteps %d: CUDA error %s\n", mode, ksteps, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, zeros = 0;
      for (int r = 0; r < 128; ++r) for (int n = 0; n < 64; ++n) {
        float acc = 0; for (int k = 0; k < 8 * ksteps; ++k) acc += a[r * 32 + k] * b[k * 64 + n];
        ref[r * 64 + n] = acc;
        if (d[r * 64 + n] != acc) ++bad;
        if (d[r * 64 + n] == 0.f) ++zeros;
      }
      printf("mode %d ksteps %d: mismatches %d / 8192, zeros %d | d[0][0..7] =", mode, ksteps, bad, zeros);
      for (int n = 0; n < 8; ++n) printf(" %g", d[n]);
      printf(" | ref =");
      for (int n = 0; n < 8; ++n) printf(" %g", ref[n]);
      printf("\n");
    }
  }
  return 0;
}
