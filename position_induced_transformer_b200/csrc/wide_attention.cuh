// K1w: position-attention for shared meshes with FEW output rows and a HUGE column set with narrow values
// (the local encoder: N = 256 latent points, M = 177 241 mesh points, B*D = 24 at Darcy-421).
//
// Roles are transposed with respect to the tall kernels: a lane owns CPL columns for the kernel's whole
// life (coordinates in registers) and the warp walks over ALL rows, whose per-row constants (point, soft-max
// shift and cut of every head) sit in a shared-memory table built once per CTA.  The sweep touches no global
// memory.  Only ~2 % of the (row, column) pairs are kept and they are spatially clustered, so a warp finds
// something to do for ~5 % of the rows; for those the kept lanes gather their value rows, and the per-row sums
// are combined across lanes with shuffles and across warps with a handful of fp32 REDs (the soft-max shift
// is known in advance, so partial sums need no rescaling).
//
//   wide_fwd_kernel     partial[(row,h), e] += sum_j P_hj U[j,e];  rowsum[(h,row)] += sum_j P_hj   (+ finalize)
//   wide_dscale_kernel  per (row,h): A += sum_j P d2 dP_j, Bq += sum_j P dP_j, m += sum_j P d2 with
//                       dP_j = <dO[row,h,:], U[j,:]> a lane-local dot product          (+ the generic finalize)
#pragma once
#include "geometry.cuh"

namespace pit {

constexpr int WIDE_THREADS = 256;  // 8 warps share one row table: 512 columns per CTA pass
constexpr int WIDE_WARPS = WIDE_THREADS / 32;
constexpr int WIDE_CPL = 2;       // columns per lane: 64 columns per warp (more warps: the row walk is latency bound)
constexpr int WIDE_MAX_WIDTH = 32;  // B*D scalars per value row
constexpr int WIDE_MAX_H = 2;

struct WideParams {
  const float* mesh_out;  // [N,sd]
  const float* mesh_in;   // [M,sd]
  const float* period;
  const float* values;  // [B,M,D]
  const float* scale;   // [H]
  const float* v_min;
  const float* v_lo;
  const float* v_hi;
  float weight;
  int masked;
  int B, H, N, M, D, sd;
  int width;  // B*D
  // forward
  float* partial;  // [N*H, width] zero-initialised
  float* rowsum;   // [H,N]       zero-initialised
  // backward
  const float* d_out;
  int64_t ld_out, col_off;
  float* dscale_terms;  // [N*H,3] zero-initialised
};

// Row table entry: x, y, reach (largest d2 a kept column can have), then per head (top, cut).
__host__ __device__ inline int wide_row_words(int H) { return 3 + 2 * H; }

template <int GEO>
__device__ __forceinline__ void wide_build_rows(const WideParams& P, float* rowtab, int* val_off, int* g_off) {
  const int rw = wide_row_words(P.H);
  for (int r = threadIdx.x; r < P.N; r += WIDE_THREADS) {
    const Point<GEO> o = load_point<GEO>(P.mesh_out, r, P.sd);
    float* t = rowtab + (size_t)r * rw;
    t[0] = o.x;
    t[1] = o.y;
    const float vmin = __ldg(P.v_min + r);
    t[2] = P.masked ? __ldg(P.v_hi + r) * 1.000001f : INFINITY;  // see tall_scan_row: d2 above it can never be kept
    for (int h = 0; h < P.H; ++h) {
      const float s = __ldg(P.scale + h);
      t[3 + 2 * h] = __fmul_rn(vmin, s);
      t[4 + 2 * h] = P.masked ? head_threshold(__ldg(P.v_lo + r), __ldg(P.v_hi + r), s, P.weight) : INFINITY;
    }
  }
  for (int e = threadIdx.x; e < P.width; e += WIDE_THREADS) {
    const int b = e / P.D, d = e - b * P.D;
    val_off[e] = b * P.M * P.D + d;  // host guarantees B*M*D < 2^31
    if (g_off) g_off[e] = d;         // + b*N*ld_out handled with 64-bit arithmetic at use
  }
}

// Bounding box of the warp's columns (Euclidean geometries) and the row pre-filter built on it: lane l tests row
// r0 + l -- the distance from the row's point to the box against the row's reach -- and the ballot is the set of rows
// of this group of 32 worth walking.  ~95 % of the (warp, row) pairs are dismissed here for 8 instructions per 32 rows.
template <int GEO>
struct WideBox {
  float x0, x1, y0, y1;
};
template <int GEO>
__device__ __forceinline__ WideBox<GEO> wide_box(const Point<GEO> (&col)[WIDE_CPL], const int (&jcol)[WIDE_CPL]) {
  WideBox<GEO> bx{INFINITY, -INFINITY, INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < WIDE_CPL; ++c) {
    if (jcol[c] >= 0) {
      bx.x0 = fminf(bx.x0, col[c].x), bx.x1 = fmaxf(bx.x1, col[c].x);
      bx.y0 = fminf(bx.y0, col[c].y), bx.y1 = fmaxf(bx.y1, col[c].y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bx.x0 = fminf(bx.x0, __shfl_xor_sync(FULL, bx.x0, o)), bx.x1 = fmaxf(bx.x1, __shfl_xor_sync(FULL, bx.x1, o));
    bx.y0 = fminf(bx.y0, __shfl_xor_sync(FULL, bx.y0, o)), bx.y1 = fmaxf(bx.y1, __shfl_xor_sync(FULL, bx.y1, o));
  }
  return bx;
}
template <int GEO>
__device__ __forceinline__ unsigned wide_rows_to_visit(const WideBox<GEO>& bx, const float* rowtab, int rw, int r0, int N, int lane) {
  const int r = r0 + lane;
  bool visit = r < N;
  if ((GEO == GEO_EUCLID1 || GEO == GEO_EUCLID2) && visit) {
    const float* t = rowtab + (size_t)r * rw;
    const float dx = fmaxf(fmaxf(bx.x0 - t[0], t[0] - bx.x1), 0.f);
    const float dy = GEO == GEO_EUCLID2 ? fmaxf(fmaxf(bx.y0 - t[1], t[1] - bx.y1), 0.f) : 0.f;
    visit = fmaf(dx, dx, dy * dy) <= t[2] * 1.0001f;  // slack for the different rounding of the box distance
  }
  return __ballot_sync(FULL, visit);
}

template <int GEO, int NH, int WPAD>
__global__ void __launch_bounds__(WIDE_THREADS, 3) wide_fwd_kernel(const WideParams P) {
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  const int rw = wide_row_words(NH);
  float* rowtab = reinterpret_cast<float*>(wide_smem_raw);
  int* val_off = reinterpret_cast<int*>(rowtab + (size_t)P.N * rw);
  wide_build_rows<GEO>(P, rowtab, val_off, nullptr);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float period = P.period ? __ldg(P.period) : 0.f;
  // A CTA takes one or more groups of 8 x 64 columns (grid-stride): the row table (and, backward, the upstream-gradient table)
  // is built once per CTA, not once per 256 columns.
  const int64_t stride = (int64_t)gridDim.x * WIDE_WARPS * (32 * WIDE_CPL);
  for (int64_t base = ((int64_t)blockIdx.x * WIDE_WARPS + warp) * (32 * WIDE_CPL); base < P.M; base += stride) {
  Point<GEO> col[WIDE_CPL];
  int jcol[WIDE_CPL];
#pragma unroll
  for (int c = 0; c < WIDE_CPL; ++c) {
    const int64_t j = base + c * 32 + lane;
    jcol[c] = j < P.M ? (int)j : -1;
    col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
  }
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  // the lane's own value rows stay in registers: u[c][e] = U[b_e, j_c, d_e]
  float u[WIDE_CPL][WPAD];
#pragma unroll
  for (int c = 0; c < WIDE_CPL; ++c)
#pragma unroll
    for (int e = 0; e < WPAD; ++e)
      u[c][e] = (e < P.width && jcol[c] >= 0) ? __ldg(P.values + val_off[e] + (int64_t)jcol[c] * P.D) : 0.f;

  const WideBox<GEO> box = wide_box<GEO>(col, jcol);
  for (int r0 = 0; r0 < P.N; r0 += 32)
  for (unsigned todo = wide_rows_to_visit<GEO>(box, rowtab, rw, r0, P.N, lane); todo; todo &= todo - 1) {
    const int r = r0 + __ffs(todo) - 1;
    const float* t = rowtab + (size_t)r * rw;
    Point<GEO> o;
    o.x = t[0];
    o.y = t[1];
    float p[WIDE_CPL][NH];
    bool any = false;
#pragma unroll
    for (int c = 0; c < WIDE_CPL; ++c) {
      const float d2 = dist2<GEO>(o, col[c], period);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        p[c][h] = 0.f;
        if (jcol[c] >= 0) {
          const float sc = __fmul_rn(d2, s[h]);
          if (sc <= t[4 + 2 * h]) p[c][h] = expf(__fsub_rn(t[3 + 2 * h], sc));
        }
        any = any || (p[c][h] > 0.f);
      }
    }
    if (!__any_sync(FULL, any)) continue;
    // ---- rare path: this warp holds kept columns of row r; everything needed is in registers ----
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < WIDE_CPL; ++c) lsum += p[c][h];
      lsum = warp_sum(lsum);
      if (lsum > 0.f) {  // warp-uniform
        // lane e ends up with element e of the warp's partial sum: a reduce-scatter butterfly (31 shuffles) instead of one
        // full warp reduction per element (5 x WPAD shuffles)
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          v[e] = 0.f;
          if (e < WPAD) {
#pragma unroll
            for (int c = 0; c < WIDE_CPL; ++c) v[e] = fmaf(p[c][h], u[c][e], v[e]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const bool upper = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < o; ++i) {
            const float keep = upper ? v[i + o] : v[i], send = upper ? v[i] : v[i + o];
            v[i] = keep + __shfl_xor_sync(FULL, send, o);
          }
        }
        const float mine = v[0];
        if (lane < P.width) atomicAdd(P.partial + ((int64_t)r * NH + h) * P.width + lane, mine);
        if (lane == 0) atomicAdd(P.rowsum + (int64_t)h * P.N + r, lsum);
      }
    }
  }
  }
}

template <int GEO, int NH, int WPAD>
__global__ void __launch_bounds__(WIDE_THREADS, 3) wide_dscale_kernel(const WideParams P) {
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  const int rw = wide_row_words(NH);
  float* rowtab = reinterpret_cast<float*>(wide_smem_raw);
  int* val_off = reinterpret_cast<int*>(rowtab + (size_t)P.N * rw);
  float* gtab = reinterpret_cast<float*>(val_off + P.width);  // [N][NH][WPAD] upstream gradient rows
  wide_build_rows<GEO>(P, rowtab, val_off, nullptr);
  // one (row, head) pair per thread and pass: WPAD independent loads in flight, sample/channel counters instead of divisions
  for (int rh = threadIdx.x; rh < P.N * NH; rh += WIDE_THREADS) {
    const int r = rh / NH, h = rh - r * NH;
    const float* src = P.d_out + (int64_t)r * P.ld_out + P.col_off + (int64_t)h * P.D;
    float* g = gtab + (size_t)rh * WPAD;
    int b = 0, d = 0;
#pragma unroll
    for (int e = 0; e < WPAD; ++e) {
      g[e] = e < P.width ? __ldg(src + (int64_t)b * P.N * P.ld_out + d) : 0.f;
      if (++d == P.D) {
        d = 0;
        ++b;
      }
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float period = P.period ? __ldg(P.period) : 0.f;
  // A CTA takes one or more groups of 8 x 64 columns (grid-stride): the row table (and, backward, the upstream-gradient table)
  // is built once per CTA, not once per 256 columns.
  const int64_t stride = (int64_t)gridDim.x * WIDE_WARPS * (32 * WIDE_CPL);
  for (int64_t base = ((int64_t)blockIdx.x * WIDE_WARPS + warp) * (32 * WIDE_CPL); base < P.M; base += stride) {
  Point<GEO> col[WIDE_CPL];
  int jcol[WIDE_CPL];
#pragma unroll
  for (int c = 0; c < WIDE_CPL; ++c) {
    const int64_t j = base + c * 32 + lane;
    jcol[c] = j < P.M ? (int)j : -1;
    col[c] = load_point<GEO>(P.mesh_in, j < P.M ? j : 0, P.sd);
  }
  float s[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) s[h] = __ldg(P.scale + h);
  float u[WIDE_CPL][WPAD];
#pragma unroll
  for (int c = 0; c < WIDE_CPL; ++c)
#pragma unroll
    for (int e = 0; e < WPAD; ++e)
      u[c][e] = (e < P.width && jcol[c] >= 0) ? __ldg(P.values + val_off[e] + (int64_t)jcol[c] * P.D) : 0.f;

  const WideBox<GEO> box = wide_box<GEO>(col, jcol);
  for (int r0 = 0; r0 < P.N; r0 += 32)
  for (unsigned todo = wide_rows_to_visit<GEO>(box, rowtab, rw, r0, P.N, lane); todo; todo &= todo - 1) {
    const int r = r0 + __ffs(todo) - 1;
    const float* t = rowtab + (size_t)r * rw;
    Point<GEO> o;
    o.x = t[0];
    o.y = t[1];
    float p[WIDE_CPL][NH], d2c[WIDE_CPL];
    bool any = false;
#pragma unroll
    for (int c = 0; c < WIDE_CPL; ++c) {
      d2c[c] = dist2<GEO>(o, col[c], period);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        p[c][h] = 0.f;
        if (jcol[c] >= 0) {
          const float sc = __fmul_rn(d2c[c], s[h]);
          if (sc <= t[4 + 2 * h]) p[c][h] = expf(__fsub_rn(t[3 + 2 * h], sc));
        }
        any = any || (p[c][h] > 0.f);
      }
    }
    if (!__any_sync(FULL, any)) continue;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float kept_w = 0.f;
#pragma unroll
      for (int c = 0; c < WIDE_CPL; ++c) kept_w += p[c][h];
      kept_w = warp_sum(kept_w);
      if (kept_w > 0.f) {  // warp-uniform
        const float* g = gtab + ((size_t)r * NH + h) * WPAD;
        float a_sum = 0.f, b_sum = 0.f, m_sum = 0.f;
#pragma unroll
        for (int c = 0; c < WIDE_CPL; ++c) {
          float dp = 0.f;  // <dO[row,h,:], U[j_c,:]>
#pragma unroll
          for (int e = 0; e < WPAD; ++e) dp = fmaf(g[e], u[c][e], dp);
          const float pd = p[c][h] * d2c[c];
          a_sum = fmaf(pd, dp, a_sum);
          b_sum = fmaf(p[c][h], dp, b_sum);
          m_sum += pd;
        }
        a_sum = warp_sum(a_sum);
        b_sum = warp_sum(b_sum);
        m_sum = warp_sum(m_sum);
        if (lane == 0) {
          float* dst = P.dscale_terms + ((int64_t)r * NH + h) * 3;
          atomicAdd(dst + 0, a_sum);
          atomicAdd(dst + 1, b_sum);
          atomicAdd(dst + 2, m_sum);
        }
      }
    }
  }
  }
}

}  // namespace pit
