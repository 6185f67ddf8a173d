// Bit-exact squared distances and small warp helpers shared by every kernel.
//
// The quantile mask of the reference (pit.py:49-50) is decided on fp32 values that tie to
// within one ulp on regular grids, so d2 must be rounded exactly like the reference's
// tensor expression: each subtraction, multiplication and addition rounded separately,
// never contracted into an FMA.  The __f*_rn intrinsics guarantee that.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pit {

constexpr unsigned FULL = 0xffffffffu;

// Geometry codes: distance variant x coordinates per point.
enum : int { GEO_EUCLID1 = 0, GEO_EUCLID2 = 1, GEO_PERIODIC1 = 2, GEO_PERIODIC2 = 3 };

template <int GEO>
struct Point {
  float x, y;
};

// Loads one mesh point; `sd` is the row stride in floats (1 or 2). The 1-D periodic variant
// only ever reads coordinate 0 (pit.py:195), whatever the stride.
template <int GEO>
__device__ __forceinline__ Point<GEO> load_point(const float* __restrict__ mesh, int64_t idx, int sd) {
  Point<GEO> p;
  const float* q = mesh + idx * sd;
  p.x = __ldg(q);
  p.y = (GEO == GEO_EUCLID2 || GEO == GEO_PERIODIC2) ? __ldg(q + 1) : 0.f;
  return p;
}

// d2(out, in) in the reference's rounding order.
//   euclid      sum((xo - xi)^2)                         pit.py:47 / 134
//   periodic1d  m = |xo - xi|; m = min(m, l - m); m^2    pit.py:193-195
//   periodic2d  same wrap per axis, then the sum         pit.py:251-253
template <int GEO>
__device__ __forceinline__ float dist2(const Point<GEO>& o, const Point<GEO>& i, float period) {
  if (GEO == GEO_EUCLID1) {
    float dx = __fsub_rn(o.x, i.x);
    return __fmul_rn(dx, dx);
  } else if (GEO == GEO_EUCLID2) {
    float dx = __fsub_rn(o.x, i.x), dy = __fsub_rn(o.y, i.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  } else if (GEO == GEO_PERIODIC1) {
    float m = fabsf(__fsub_rn(o.x, i.x));
    m = fminf(m, __fsub_rn(period, m));
    return __fmul_rn(m, m);
  } else {
    float mx = fabsf(__fsub_rn(o.x, i.x)), my = fabsf(__fsub_rn(o.y, i.y));
    mx = fminf(mx, __fsub_rn(period, mx));
    my = fminf(my, __fsub_rn(period, my));
    return __fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my));
  }
}

// torch.lerp(a, b, w) as ATen evaluates it on both CPU (vectorised fmadd) and CUDA (contracted):
//   w < 0.5 : fma(w,     b - a, a)        else : fma(w - 1, b - a, b)
__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
  float diff = __fsub_rn(b, a);
  return (fabsf(w) < 0.5f) ? __fmaf_rn(w, diff, a) : __fmaf_rn(__fsub_rn(w, 1.0f), diff, b);
}

// Per-head cut of the locality mask: T = lerp(fl(v_lo*s), fl(v_hi*s), w)  (SURVEY 8a step 3).
__device__ __forceinline__ float head_threshold(float v_lo, float v_hi, float s, float w) {
  return torch_lerp(__fmul_rn(v_lo, s), __fmul_rn(v_hi, s), w);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

}  // namespace pit
