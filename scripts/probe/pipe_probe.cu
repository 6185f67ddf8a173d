// Throughput probe for the epilogue design of the decoder tail: FFMA vs FFMA2 (packed fp32x2) vs MUFU.EX2 / MUFU.RCP per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float seed) {
  float a[8];
  unsigned long long p[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-6f + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = (unsigned long long)__float_as_uint(a[2 * i]) | ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32);
  const float m = 0.999f + seed * 1e-9f, c = 1e-3f;
  const unsigned long long m2 = (unsigned long long)__float_as_uint(m) | ((unsigned long long)__float_as_uint(m) << 32);
  const unsigned long long c2 = (unsigned long long)__float_as_uint(c) | ((unsigned long long)__float_as_uint(c) << 32);
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {  // 8 FFMA (3-reg)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
    } else if (MODE == 1) {  // 4 FFMA2 = 8 fp32 FMAs
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
    } else if (MODE == 2) {  // 8 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    } else if (MODE == 3) {  // 4 FFMA + 4 LOP3-ish (alu pipe) interleaved
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = fmaf(a[i], m, c);
        a[4 + i] = __uint_as_float((__float_as_uint(a[4 + i]) + 0x1000u) & 0xffffe000u);
      }
    } else if (MODE == 4) {  // GELU-like mix: 2 MUFU + 12 FFMA per element, 4 elements
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x = a[i], e, t;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * x));
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(x, 0.23f, 1.f)));
        float q = fmaf(t, 1.06f, -1.45f);
        q = fmaf(t, q, 1.42f); q = fmaf(t, q, -0.28f); q = fmaf(t, q, 0.25f); q *= t;
        float r = fmaf(-q, e, 1.f);
        a[i] = fmaf(x, r, 0.5f * x) * 0.7f + 0.1f;
      }
    } else if (MODE == 5) {  // 8 FFMA with immediate operands
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 0.999f, 1e-3f);
    } else if (MODE == 6) {  // 8 FMUL
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a[i] * m;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter_thread, int sms) {
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  const int blocks = sms * 8;  // 8 CTAs of 8 warps per SM = full occupancy
  probe<MODE><<<blocks, 256>>>(out, 1.f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  probe<MODE><<<blocks, 256>>>(out, 1.f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double total = (double)blocks * 256 * ITERS * ops_per_iter_thread;
  printf("%-34s %8.3f ms  %8.1f Gop/s  %6.1f op/clk/SM (at %d MHz nominal)\n", name, ms, total / ms * 1e-6, total / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  run<0>("FFMA 3-reg (fp32 FMAs)", 8, sms);
  run<5>("FFMA imm (fp32 FMAs)", 8, sms);
  run<6>("FMUL", 8, sms);
  run<1>("FFMA2 (fp32 FMAs, 2 per lane-op)", 8, sms);
  run<2>("MUFU.EX2", 8, sms);
  run<3>("FFMA + LOP3 pairs (instr)", 12, sms);
  run<4>("GELU-like mix (elements)", 4, sms);
  return 0;
}
