// Throughput of the legacy warp-level TF32 MMA (mma.sync.m16n8k8 -> HMMA.1688.F32.TF32) on one SM sub-partition, as used by the
// decoder-tail kernels:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu && ./hmma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int CHAINS>
__global__ void __launch_bounds__(1024) probe(float* out) {
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) d[c][e] = threadIdx.x * 1e-6f;
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 123.456f) out[0] = s;
}

template <int CHAINS>
void run(int warps_per_sm, int sms) {
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  probe<CHAINS><<<sms, 32 * warps_per_sm>>>(out);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  probe<CHAINS><<<sms, 32 * warps_per_sm>>>(out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mmas = (double)sms * warps_per_sm * ITERS * CHAINS;
  const double cyc = ms * 1e-3 * khz * 1e3;
  printf("chains %d warps/SM %2d: %7.3f ms  %6.1f TFLOP/s  %5.2f cycles per HMMA per SM sub-partition (at %d MHz nominal)\n", CHAINS, warps_per_sm, ms,
         mmas * 2048 / (ms * 1e-3) * 1e-12, cyc / (mmas / sms / 4), khz / 1000);
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int w : {4, 8, 16, 32}) run<1>(w, sms);
  for (int w : {4, 8, 16, 32}) run<4>(w, sms);
  for (int w : {4, 8, 16}) run<8>(w, sms);
  return 0;
}
