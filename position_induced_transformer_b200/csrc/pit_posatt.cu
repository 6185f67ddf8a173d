// libpit_posatt.so -- C ABI (include/pit_posatt.h) over the sm_100a position-attention kernels.
// Host side only validates, picks a launch shape and enqueues; it never allocates or syncs.
#include "../../include/pit_posatt.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "local_attention.cuh"
#include "rowstat.cuh"

namespace {

thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

#define PIT_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) return fail(PIT_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

#define PIT_LAUNCHED()                                                                        \
  do {                                                                                        \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                       \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess) return fail(PIT_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e_)); \
  } while (0)

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;  // B200
  }
  return cached;
}

int check_problem(const pit_problem_t* p) {
  if (!p) return fail(PIT_ERR_ARG, "problem is null");
  if (p->variant < PIT_EUCLID || p->variant > PIT_PERIODIC2D) return fail(PIT_ERR_ARG, "unknown variant %d", p->variant);
  if (p->space_dim != 1 && p->space_dim != 2) return fail(PIT_ERR_ARG, "space_dim must be 1 or 2, got %d", p->space_dim);
  if (p->variant == PIT_PERIODIC2D && p->space_dim != 2) return fail(PIT_ERR_ARG, "periodic2d needs space_dim 2");
  if (p->batch < 1 || p->n_head < 1 || p->n_out < 1 || p->n_in < 1 || p->dim < 1)
    return fail(PIT_ERR_ARG, "batch/n_head/n_out/n_in/dim must be positive");
  if ((int64_t)p->batch * p->n_in * p->dim >= (1ll << 40)) return fail(PIT_ERR_ARG, "values tensor too large");
  return PIT_OK;
}

int geo_of(const pit_problem_t* p) {
  if (p->variant == PIT_PERIODIC1D) return pit::GEO_PERIODIC1;
  if (p->variant == PIT_PERIODIC2D) return pit::GEO_PERIODIC2;
  return p->space_dim == 1 ? pit::GEO_EUCLID1 : pit::GEO_EUCLID2;
}

// Launch shape of the warp-per-row kernels.
struct Shape {
  int vec, a, chunks, n_split, split_len;
  int64_t items;
  int width;
};

// `owners`: rows (forward / dscale: rows_total*H) or columns (dvalues: cols_total) that each get a warp;
// `red_len`: length of the index the warp sweeps.
Shape make_shape(const pit_problem_t* p, int64_t owners, int red_len, bool allow_split) {
  Shape s;
  s.items = owners;
  s.width = p->mesh_batched ? p->dim : p->batch * p->dim;
  s.vec = (p->dim % 4 == 0) ? 4 : 1;
  const int groups = (s.width + 32 * s.vec - 1) / (32 * s.vec);  // 32-lane groups needed to cover the vector
  s.a = groups >= 4 ? 4 : (groups >= 2 ? 2 : 1);
  s.chunks = (groups + s.a - 1) / s.a;
  s.n_split = 1;
  s.split_len = (red_len + 31) / 32 * 32;
  if (allow_split) {
    const int64_t warps = owners * s.chunks;
    const int64_t target = (int64_t)sm_count() * 32;
    if (warps < target && red_len > 1024) {
      int64_t want = (target + warps - 1) / warps;
      int64_t most = (red_len + 511) / 512;
      int64_t n = want < most ? want : most;
      if (n > 1) {
        int len = (int)((red_len + n - 1) / n);
        len = (len + 127) / 128 * 128;
        s.split_len = len;
        s.n_split = (red_len + len - 1) / len;
      }
    }
  }
  return s;
}

pit::AttnParams base_params(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                            const float* values, const float* scale, const pit_rowstat_t* st) {
  pit::AttnParams P{};
  P.mesh_out = mesh_out;
  P.mesh_in = mesh_in;
  P.period = p->variant == PIT_EUCLID ? nullptr : period;
  P.values = values;
  P.scale = scale;
  P.v_min = st->v_min;
  P.v_lo = st->v_lo;
  P.v_hi = st->v_hi;
  P.weight = st->weight;
  P.masked = st->masked;
  P.B = p->batch;
  P.H = p->n_head;
  P.N = p->n_out;
  P.M = p->n_in;
  P.D = p->dim;
  P.sd = p->space_dim;
  P.mesh_batched = p->mesh_batched;
  P.width = p->mesh_batched ? p->dim : p->batch * p->dim;
  return P;
}

int check_stat(const pit_problem_t* p, const pit_rowstat_t* st, const float* period) {
  if (!st || !st->v_min) return fail(PIT_ERR_ARG, "rowstat.v_min is required");
  if (st->masked && (!st->v_lo || !st->v_hi)) return fail(PIT_ERR_ARG, "masked stage needs rowstat.v_lo / v_hi");
  if (p->variant != PIT_EUCLID && !period) return fail(PIT_ERR_ARG, "periodic variant needs the period pointer");
  return PIT_OK;
}

inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// Dispatch a <GEO, VEC, A> kernel template.
#define PIT_DISPATCH_A(KERNEL, GEO, VEC, a, ...)                       \
  switch (a) {                                                          \
    case 1: KERNEL<GEO, VEC, 1> __VA_ARGS__; break;                     \
    case 2: KERNEL<GEO, VEC, 2> __VA_ARGS__; break;                     \
    default: KERNEL<GEO, VEC, 4> __VA_ARGS__; break;                    \
  }
#define PIT_DISPATCH_VEC(KERNEL, GEO, vec, a, ...)                      \
  if ((vec) == 4) {                                                     \
    PIT_DISPATCH_A(KERNEL, GEO, 4, a, __VA_ARGS__)                      \
  } else {                                                              \
    PIT_DISPATCH_A(KERNEL, GEO, 1, a, __VA_ARGS__)                      \
  }
#define PIT_DISPATCH(KERNEL, geo, vec, a, ...)                                              \
  switch (geo) {                                                                             \
    case pit::GEO_EUCLID1: PIT_DISPATCH_VEC(KERNEL, pit::GEO_EUCLID1, vec, a, __VA_ARGS__) break;     \
    case pit::GEO_EUCLID2: PIT_DISPATCH_VEC(KERNEL, pit::GEO_EUCLID2, vec, a, __VA_ARGS__) break;     \
    case pit::GEO_PERIODIC1: PIT_DISPATCH_VEC(KERNEL, pit::GEO_PERIODIC1, vec, a, __VA_ARGS__) break; \
    default: PIT_DISPATCH_VEC(KERNEL, pit::GEO_PERIODIC2, vec, a, __VA_ARGS__) break;                 \
  }

}  // namespace

extern "C" {

int pit_abi_version(void) { return PIT_ABI_VERSION; }
const char* pit_last_error(void) { return g_error; }
uint64_t pit_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int pit_quantile_ranks(double q, int32_t m, int32_t* k_lo, int32_t* k_hi, float* w) {
  if (m < 1 || !k_lo || !k_hi || !w) return fail(PIT_ERR_ARG, "bad arguments");
  if (!(q >= 0.0 && q <= 1.0)) return fail(PIT_ERR_ARG, "quantile() q must be in [0, 1], got %g", q);
  // ATen: ranks = q_tensor(fp32) * (m - 1); floor / ceil / lerp weight, all in fp32.
  volatile float rank = (float)q * (float)(m - 1);
  const float lo = floorf(rank);
  *k_lo = (int32_t)lo;
  *k_hi = (int32_t)ceilf(rank);
  *w = rank - lo;
  return PIT_OK;
}

size_t pit_workspace_bytes(const pit_problem_t* p) {
  if (check_problem(p) != PIT_OK) return 0;
  const int64_t rows_total = (int64_t)(p->mesh_batched ? p->batch : 1) * p->n_out;
  const Shape f = make_shape(p, rows_total * p->n_head, p->n_in, true);
  size_t fwd = f.n_split > 1 ? (size_t)f.items * f.width * sizeof(float) : 0;
  size_t bwd = (size_t)rows_total * p->n_head * 3 * sizeof(float);
  size_t need = fwd > bwd ? fwd : bwd;
  return need + 256;
}

int pit_rowstat(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                int32_t k_lo, int32_t k_hi, float* v_min, float* v_lo, float* v_hi, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !v_min || !v_lo || !v_hi) return fail(PIT_ERR_ARG, "null pointer");
  if (p->variant != PIT_EUCLID && !period) return fail(PIT_ERR_ARG, "periodic variant needs the period pointer");
  if (k_lo < 0 || k_hi < k_lo || k_hi > k_lo + 1 || k_hi >= p->n_in)
    return fail(PIT_ERR_ARG, "ranks out of range: k_lo=%d k_hi=%d M=%d", k_lo, k_hi, p->n_in);
  pit::RowstatParams R{};
  R.mesh_out = mesh_out;
  R.mesh_in = mesh_in;
  R.period = p->variant == PIT_EUCLID ? nullptr : period;
  R.v_min = v_min;
  R.v_lo = v_lo;
  R.v_hi = v_hi;
  R.N = p->n_out;
  R.M = p->n_in;
  R.sd = p->space_dim;
  R.mesh_batched = p->mesh_batched;
  R.rows_total = (p->mesh_batched ? p->batch : 1) * p->n_out;
  R.k_lo = k_lo;
  R.k_hi = k_hi;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int geo = geo_of(p);
#define ROWSTAT_WARP(GEO, Rn) pit::rowstat_warp_kernel<GEO, Rn><<<(R.rows_total + 3) / 4, 128, 0, st>>>(R)
#define ROWSTAT_GEO(GEO)                                                               \
  if (R.M <= 128) ROWSTAT_WARP(GEO, 4);                                                \
  else if (R.M <= 256) ROWSTAT_WARP(GEO, 8);                                           \
  else if (R.M <= 512) ROWSTAT_WARP(GEO, 16);                                          \
  else if (R.M <= 1024) ROWSTAT_WARP(GEO, 32);                                         \
  else pit::rowstat_block_kernel<GEO><<<R.rows_total, pit::ROWSTAT_BLOCK, 0, st>>>(R)
  switch (geo) {
    case pit::GEO_EUCLID1: ROWSTAT_GEO(pit::GEO_EUCLID1); break;
    case pit::GEO_EUCLID2: ROWSTAT_GEO(pit::GEO_EUCLID2); break;
    case pit::GEO_PERIODIC1: ROWSTAT_GEO(pit::GEO_PERIODIC1); break;
    default: ROWSTAT_GEO(pit::GEO_PERIODIC2); break;
  }
#undef ROWSTAT_GEO
#undef ROWSTAT_WARP
  PIT_LAUNCHED();
  return PIT_OK;
}

int pit_posatt_forward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                       const float* values, const float* scale, const pit_rowstat_t* stat, float* out, int64_t ld_out,
                       int64_t col_off, int32_t copy_values, float* rowsum, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !values || !scale || !out || !rowsum) return fail(PIT_ERR_ARG, "null pointer");
  if (int rc = check_stat(p, stat, period)) return rc;
  if (ld_out < col_off + (int64_t)p->n_head * p->dim || col_off < 0) return fail(PIT_ERR_ARG, "ld_out/col_off do not fit H*D");
  if (copy_values && (p->n_out != p->n_in || col_off < p->dim)) return fail(PIT_ERR_ARG, "copy_values needs N == M and col_off >= D");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows_total = (int64_t)(p->mesh_batched ? p->batch : 1) * p->n_out;
  Shape s = make_shape(p, rows_total * p->n_head, p->n_in, true);
  if (s.vec == 4 && (!aligned16(values) || !aligned16(out) || (ld_out % 4) || (col_off % 4))) {
    return fail(PIT_ERR_ARG, "values/out must be 16-byte aligned with ld_out, col_off multiples of 4 when D %% 4 == 0");
  }
  pit::AttnParams P = base_params(p, mesh_out, mesh_in, period, values, scale, stat);
  P.out = out;
  P.ld_out = ld_out;
  P.col_off = col_off;
  P.rowsum = rowsum;
  P.split_len = s.split_len;
  if (s.n_split > 1) {
    const size_t need = (size_t)s.items * s.width * sizeof(float);
    if (!workspace || workspace_bytes < need) return fail(PIT_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    P.partial = static_cast<float*>(workspace);
    PIT_CUDA(cudaMemsetAsync(P.partial, 0, need, st));
    PIT_CUDA(cudaMemsetAsync(rowsum, 0, (size_t)s.items * sizeof(float), st));
  }
  if (copy_values) {  // first D columns of the concat output (pit.py:44)
    PIT_CUDA(cudaMemcpy2DAsync(out, (size_t)ld_out * sizeof(float), values, (size_t)p->dim * sizeof(float),
                               (size_t)p->dim * sizeof(float), (size_t)p->batch * p->n_in, cudaMemcpyDeviceToDevice, st));
  }
  const dim3 grid((unsigned)((s.items + pit::WARPS_PER_BLOCK - 1) / pit::WARPS_PER_BLOCK), s.chunks, s.n_split);
  const dim3 block(pit::WARPS_PER_BLOCK * 32);
  PIT_DISPATCH(pit::posatt_fwd_kernel, geo_of(p), s.vec, s.a, <<<grid, block, 0, st>>>(P));
  PIT_LAUNCHED();
  if (s.n_split > 1) {
    const int64_t total = s.items * s.width;
    pit::posatt_fwd_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P);
    PIT_LAUNCHED();
  }
  return PIT_OK;
}

int pit_posatt_backward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                        const float* values, const float* scale, const pit_rowstat_t* stat, const float* rowsum,
                        const float* d_out, int64_t ld_out, int64_t col_off, int32_t accumulate_concat, float* d_values,
                        float* d_scale_rows, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !values || !scale || !rowsum || !d_out) return fail(PIT_ERR_ARG, "null pointer");
  if (int rc = check_stat(p, stat, period)) return rc;
  if (ld_out < col_off + (int64_t)p->n_head * p->dim || col_off < 0) return fail(PIT_ERR_ARG, "ld_out/col_off do not fit H*D");
  if (accumulate_concat && (p->n_out != p->n_in || col_off < p->dim))
    return fail(PIT_ERR_ARG, "accumulate_concat needs N == M and col_off >= D");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows_total = (int64_t)(p->mesh_batched ? p->batch : 1) * p->n_out;
  const int64_t cols_total = (int64_t)(p->mesh_batched ? p->batch : 1) * p->n_in;
  const int geo = geo_of(p);
  const bool vec4 = p->dim % 4 == 0;
  if (vec4 && (!aligned16(values) || !aligned16(d_out) || (d_values && !aligned16(d_values)) || (ld_out % 4) || (col_off % 4)))
    return fail(PIT_ERR_ARG, "values/d_out/d_values must be 16-byte aligned with ld_out, col_off multiples of 4 when D %% 4 == 0");
  pit::AttnParams P = base_params(p, mesh_out, mesh_in, period, values, scale, stat);
  P.rowsum = const_cast<float*>(rowsum);
  P.d_out = d_out;
  P.ld_out = ld_out;
  P.col_off = col_off;
  P.d_values = d_values;
  P.add_concat = accumulate_concat;
  const dim3 block(pit::WARPS_PER_BLOCK * 32);

  if (d_scale_rows) {
    Shape s = make_shape(p, rows_total * p->n_head, p->n_in, true);
    const size_t need = (size_t)s.items * 3 * sizeof(float);
    if (!workspace || workspace_bytes < need) return fail(PIT_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    P.dscale_terms = static_cast<float*>(workspace);
    P.split_len = s.split_len;
    PIT_CUDA(cudaMemsetAsync(P.dscale_terms, 0, need, st));
    const dim3 grid((unsigned)((s.items + pit::WARPS_PER_BLOCK - 1) / pit::WARPS_PER_BLOCK), s.chunks, s.n_split);
    PIT_DISPATCH(pit::posatt_dscale_kernel, geo, s.vec, s.a, <<<grid, block, 0, st>>>(P));
    PIT_LAUNCHED();
    pit::posatt_dscale_finalize_kernel<<<(unsigned)((s.items + 255) / 256), 256, 0, st>>>(P, d_scale_rows);
    PIT_LAUNCHED();
  }
  if (d_values) {
    Shape s = make_shape(p, cols_total, p->n_out * p->n_head, true);
    // the sweep runs over rows i for every head, so the split is expressed in rows
    if (s.n_split > 1) {
      int64_t n = s.n_split;
      int len = (int)((p->n_out + n - 1) / n);
      len = (len + 127) / 128 * 128;
      s.split_len = len;
      s.n_split = (p->n_out + len - 1) / len;
    } else {
      s.split_len = (p->n_out + 31) / 32 * 32;
    }
    P.split_len = s.split_len;
    if (s.n_split > 1) PIT_CUDA(cudaMemsetAsync(d_values, 0, (size_t)p->batch * p->n_in * p->dim * sizeof(float), st));
    const dim3 grid((unsigned)((s.items + pit::WARPS_PER_BLOCK - 1) / pit::WARPS_PER_BLOCK), s.chunks, s.n_split);
    PIT_DISPATCH(pit::posatt_dvalues_kernel, geo, s.vec, s.a, <<<grid, block, 0, st>>>(P));
    PIT_LAUNCHED();
  }
  return PIT_OK;
}

}  // extern "C"
