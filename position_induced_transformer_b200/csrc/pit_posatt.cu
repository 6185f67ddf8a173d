// libpit_posatt.so -- C ABI (include/pit_posatt.h) over the sm_100a position-attention kernels.
// Host side only validates, picks a launch shape and enqueues; it never allocates or syncs.
#include "../../include/pit_posatt.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "launchers.h"

namespace {

using pit::launch::TallPlan;
using pit::launch::WidePlan;
namespace launch = pit::launch;

thread_local char g_error[512] = "";
std::atomic<int> g_dense_precision{0};   // pit_set_dense_precision: PIT_DENSE_FP32 / TF32 / BF16
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

// Device attributes are cached per device index (a process may drive GPUs of different kinds).
int device_attr(cudaDeviceAttr attr, int slot, int fallback) {
  static int cache[2][64];   // zero-initialised; [slot][device]
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fallback;
  if (cache[slot][dev] == 0) {
    int v = 0;
    cache[slot][dev] = (cudaDeviceGetAttribute(&v, attr, dev) == cudaSuccess && v > 0) ? v : fallback;
  }
  return cache[slot][dev];
}

int sm_count() { return device_attr(cudaDevAttrMultiProcessorCount, 0, 148 /* B200 */); }

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

pit::TailPlanDev tail_plan_view(const pit_tail_plan_t* plan, int tiles_per_cta, int round);


// ---------------------------------------------------------------------------------------------
// "tall" kernels (shared meshes, M <= 1024, H <= 2): eligibility, launch shape, dispatch
// ---------------------------------------------------------------------------------------------
int max_smem_optin() { return device_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, 1, 227 * 1024); }

bool tall_eligible(const pit_problem_t* p) {
  return !p->mesh_batched && p->n_in <= pit::TALL_MAX_M && p->dim % 4 == 0 && p->n_head <= pit::TALL_MAX_H;
}

int cpl_of(int m) { return m <= 256 ? 8 : 32; }  // instantiated column-per-lane counts

// forward: one warp per row, 32*l4 float4 lanes per warp pass, `chunks` passes over blockIdx.y
TallPlan plan_tall_fwd(const pit_problem_t* p) {
  TallPlan c{};
  if (!tall_eligible(p)) return c;
  c.lanes4 = p->batch * p->dim / 4;
  c.l4 = c.lanes4 >= 128 ? 4 : (c.lanes4 >= 64 ? 2 : 1);
  c.chunks = (c.lanes4 + 32 * c.l4 - 1) / (32 * c.l4);
  c.cpl = cpl_of(p->n_in);
  c.smem = (size_t)pit::TALL_WARPS * c.cpl * 32 * 16;
  const int64_t warps_wanted = (int64_t)sm_count() * 24;
  int64_t rows = (p->n_out + warps_wanted - 1) / warps_wanted;
  if (rows < 1) rows = 1;
  c.rows_per_unit = (int)rows;
  const int64_t warps = (p->n_out + rows - 1) / rows;
  c.grid = (int)((warps + pit::TALL_WARPS - 1) / pit::TALL_WARPS);
  c.ok = true;
  return c;
}

// backward: CTA-cooperative, thread owns l4 float4 lanes, lanes4 <= 4 * 128
TallPlan plan_tall_bwd(const pit_problem_t* p, bool with_values) {
  TallPlan c{};
  if (!tall_eligible(p)) return c;
  c.lanes4 = p->batch * p->dim / 4;
  if (c.lanes4 > 4 * pit::TALL_THREADS) return c;
  c.l4 = (c.lanes4 + pit::TALL_THREADS - 1) / pit::TALL_THREADS;
  if (c.l4 == 3) c.l4 = 4;
  c.cpl = cpl_of(p->n_in);
  const int budget = max_smem_optin();
  c.n_slots = 0;
  if (with_values) {
    const int per_slot = c.lanes4 * 16;
    c.n_slots = 32 * 1024 / per_slot;
    if (c.n_slots > 64) c.n_slots = 64;
    if (c.n_slots > p->n_in) c.n_slots = p->n_in;
    if (c.n_slots < 4) return c;
  }
  c.smem = pit::tall_bwd_smem_bytes(c.cpl, p->n_in, c.lanes4, c.n_slots);
  if (c.smem > (size_t)budget - 1024) return c;
  int per_sm = (int)((size_t)budget / (c.smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int target = sm_count() * per_sm;
  int rows = (p->n_out + target - 1) / target;
  rows = (rows + pit::TALL_WARPS - 1) / pit::TALL_WARPS * pit::TALL_WARPS;
  c.rows_per_unit = rows;
  c.grid = (p->n_out + rows - 1) / rows;
  c.chunks = 1;
  c.ok = true;
  return c;
}

pit::TallParams tall_params(const pit_problem_t* p, const TallPlan& c, const float* mesh_out, const float* mesh_in,
                            const float* period, const float* values, const float* scale, const pit_rowstat_t* st) {
  pit::TallParams P{};
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
  P.lanes4 = c.lanes4;
  P.rows_per_unit = c.rows_per_unit;
  P.n_slots = c.n_slots;
  return P;
}

// ---------------------------------------------------------------------------------------------
// "wide" kernels (shared meshes, few rows, huge column set, narrow values): the local encoder
// ---------------------------------------------------------------------------------------------
WidePlan plan_wide(const pit_problem_t* p, const pit_rowstat_t* st) {
  WidePlan w{};
  const int width = p->batch * p->dim;
  if (p->mesh_batched || !st->masked || p->n_head > pit::WIDE_MAX_H || width > pit::WIDE_MAX_WIDTH) return w;
  if (p->n_in < 4096 || p->n_in < 8 * p->n_out) return w;
  if ((int64_t)p->batch * p->n_in * p->dim >= (1ll << 31)) return w;
  w.smem = ((size_t)p->n_out * pit::wide_row_words(p->n_head) + 2 * (size_t)width) * sizeof(float);
  if (w.smem + (size_t)p->n_out * p->n_head * 32 * sizeof(float) > 160 * 1024) return w;
  const int64_t warps = ((int64_t)p->n_in + 32 * pit::WIDE_CPL - 1) / (32 * pit::WIDE_CPL);
  const int64_t groups = (warps + pit::WIDE_WARPS - 1) / pit::WIDE_WARPS;  // 512 columns each
  const int64_t per_cta = (groups + 4 * sm_count() - 1) / (4 * sm_count());  // grid-stride: equal shares, <= 4 CTAs per SM
  w.grid = (int)((groups + per_cta - 1) / per_cta);
  w.ok = true;
  return w;
}

// Encoder-side ("columns") tile plan: usable when it describes this stage's column set
bool column_plan_ok(const pit_problem_t* p, const pit_tail_plan_t* plan) {
  return plan && plan->rec && plan->tile_off && plan->tile_cnt && plan->cand && plan->d2 && p->n_out <= pit::TALL_MAX_M &&
         plan->n_tiles == (p->n_in + pit::TP_ROWS - 1) / pit::TP_ROWS;
}
int column_plan_grid(const pit_tail_plan_t* plan) {
  const int want = (plan->n_tiles + pit::WP_WARPS - 1) / pit::WP_WARPS;   // one tile per warp ...
  // ... up to three CTAs per SM: every CTA first builds its row (and gradient) tables, so fewer CTAs walking more tiles each
  // beat one tile per warp (Darcy-421, 5 539 tiles: 693 CTAs 75 / 60 us backward / forward, 444 CTAs 69 / 59 us, 148 CTAs 77 / 72 us)
  int cap = sm_count() * 3;
  if (const char* e = getenv("PIT_WIDE_GRID")) cap = atoi(e) > 0 ? atoi(e) : cap;   // tuning hook
  return want < cap ? want : cap;
}

pit::WideParams wide_params(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                            const float* values, const float* scale, const pit_rowstat_t* st) {
  pit::WideParams P{};
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
  P.width = p->batch * p->dim;
  return P;
}

// ---------------------------------------------------------------------------------------------
// dense (global) stages on tcgen05: eligibility, tile width, dispatch
// ---------------------------------------------------------------------------------------------
bool dense_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PIT_DENSE_TCGEN05");  // debugging switch: "0" keeps global stages on the SIMT kernels
    cached = (e && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

bool dense_eligible(const pit_problem_t* p, const pit_rowstat_t* st) {
  return dense_enabled() && !st->masked && p->dim % 4 == 0;
}

pit::DenseParams dense_params(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                              const float* scale, const pit_rowstat_t* st, int mode) {
  pit::DenseParams P{};
  P.operand_prec = g_dense_precision.load(std::memory_order_relaxed);
  const bool dv = mode == pit::DENSE_DVALUES;
  P.mesh_own = dv ? mesh_in : mesh_out;
  P.mesh_red = dv ? mesh_out : mesh_in;
  P.period = p->variant == PIT_EUCLID ? nullptr : period;
  P.scale = scale;
  P.v_min = st->v_min;
  P.n_own = dv ? p->n_in : p->n_out;
  P.n_red = dv ? p->n_out : p->n_in;
  P.N = p->n_out;
  P.M = p->n_in;
  P.B = p->batch;
  P.H = p->n_head;
  P.D = p->dim;
  P.sd = p->space_dim;
  P.mesh_batched = p->mesh_batched;
  P.width = p->mesh_batched ? p->dim : p->batch * p->dim;
  return P;
}

// Value gradient of a small stage: its K loop runs over the rows of every head; when the grid is far from filling the chip the
// heads go to separate CTAs instead (half the serial K blocks each) and meet in d_values with REDs.
bool dense_split_heads(const pit_problem_t* p) {
  if (p->n_head < 2) return false;
  const int64_t width = p->mesh_batched ? p->dim : (int64_t)p->batch * p->dim;
  const int64_t ctas = (int64_t)((p->n_in + pit::DENSE_ROWS - 1) / pit::DENSE_ROWS) * ((width + 63) / 64) * (p->mesh_batched ? p->batch : 1);
  return ctas * p->n_head <= sm_count();
}

dim3 dense_grid(int mode, const pit_problem_t* p, const pit::DenseParams& P, int* nv_out) {
  const int row_tiles = (P.n_own + pit::DENSE_ROWS - 1) / pit::DENSE_ROWS;
  const int z = mode == pit::DENSE_DVALUES ? (p->mesh_batched ? p->batch : 1) * (P.split_heads ? p->n_head : 1)
                                           : p->n_head * (p->mesh_batched ? p->batch : 1);
  // Column-tile width by a measured cost model.  One CTA per SM, so a launch takes ceil(CTAs / SMs) waves; the time of a CTA
  // grows affinely with its tile width, ~ (52 + nv) units (fits the B200 timings of the elasticity, NACA and cylinder shapes
  // at 64 / 128 / 256 columns, and does not depend on the operand precision: the kernel is bound by its producer warps, not
  // by the tensor pipe); the scale-gradient mode generates a second operand and holds two accumulators (<= 128 columns each).
  // E.g. NACA's value gradient: 120 CTAs of 128 columns in one wave beat 240 CTAs of 64 columns in two.
  const bool two_acc = mode == pit::DENSE_DSCALE;
  int best_nv = 64;
  double best_cost = 1e300;
  for (int nv = (two_acc ? 128 : 256); nv >= 64; nv /= 2) {
    if (nv > 64 && nv / 2 >= P.width) continue;                       // a narrower tile already covers every column
    const int64_t ctas = (int64_t)row_tiles * z * ((P.width + nv - 1) / nv);
    const int64_t waves = (ctas + sm_count() - 1) / sm_count();
    const double cost = (double)waves * (two_acc ? 83.0 + 2.0 * nv : 52.0 + nv);
    if (cost < best_cost) best_cost = cost, best_nv = nv;
  }
  *nv_out = best_nv;
  return dim3(row_tiles, (P.width + best_nv - 1) / best_nv, z);
}

cudaError_t dense_launch(int mode, int geo, const pit_problem_t* p, const pit::DenseParams& P, cudaStream_t st) {
  int nv = 0;
  const dim3 grid = dense_grid(mode, p, P, &nv);
  return launch::dense(mode, geo, nv, grid, P, st);
}

// ---------------------------------------------------------------------------------------------
// fused decoder tail (shared meshes, M <= 1024, H <= 2, hidden width a power of two in [32, 512], out_dim <= 4)
// ---------------------------------------------------------------------------------------------
bool tail_eligible(const pit_problem_t* p, int out_dim) {
  const int c = p->dim;
  return tall_eligible(p) && c >= 32 && c <= 512 && (c & (c - 1)) == 0 && out_dim >= 1 && out_dim <= pit::TAIL_MAX_OUT &&
         (int64_t)p->batch * p->n_in * p->n_head * c < (1ll << 31);
}

TallPlan plan_tail_fwd(const pit_problem_t* p) {
  TallPlan c = plan_tall_fwd(p);  // same row partition; lanes4 = B*C/4 (dim holds the hidden width)
  const int k_per_b = p->dim / 4 > 32 ? p->dim / 128 : 1;
  if (c.ok && c.l4 % k_per_b != 0) c.l4 = 4;
  if (c.ok) c.chunks = (c.lanes4 + 32 * c.l4 - 1) / (32 * c.l4);
  if (c.ok) c.smem += (size_t)p->dim * (1 + pit::TAIL_MAX_OUT) * sizeof(float);  // b1 and W2 behind the entry segments
  return c;
}

TallPlan plan_tail_bwd(const pit_problem_t* p, int out_dim) {
  TallPlan c{};
  if (!tall_eligible(p)) return c;
  c.lanes4 = p->batch * p->dim / 4;
  if (c.lanes4 > 4 * pit::TALL_THREADS) return c;
  c.l4 = (c.lanes4 + pit::TALL_THREADS - 1) / pit::TALL_THREADS;
  if (c.l4 == 3) c.l4 = 4;
  c.cpl = cpl_of(p->n_in);
  const int budget = max_smem_optin();
  const int per_slot = p->n_head * c.lanes4 * 16;
  c.n_slots = 32 * 1024 / per_slot;
  if (c.n_slots > 64) c.n_slots = 64;
  if (c.n_slots > p->n_in) c.n_slots = p->n_in;
  if (c.n_slots < 4) return c;
  c.smem = pit::tail_bwd_smem_bytes(c.cpl, p->n_in, c.lanes4, p->n_head, c.n_slots, p->dim, out_dim);
  if (c.smem > (size_t)budget - 1024) return c;
  int per_sm = (int)((size_t)budget / (c.smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int target = sm_count() * per_sm;
  int rows = (p->n_out + target - 1) / target;
  rows = (rows + pit::TALL_WARPS - 1) / pit::TALL_WARPS * pit::TALL_WARPS;
  c.rows_per_unit = rows;
  c.grid = (p->n_out + rows - 1) / rows;
  c.chunks = 1;
  c.ok = true;
  return c;
}

// Tensor-core variant (decoder_tail_mma.cuh): hidden width 32, 64 or 128 and B*C a multiple of 128.
bool tail_mma_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("PIT_TAIL_MMA");  // debugging switch: "0" keeps the decoder tail on the SIMT gather kernels
    cached = (e && e[0] == '0') ? 0 : 1;
  }
  return cached == 1;
}

bool tail_mma_eligible(const pit_problem_t* p, int out_dim) {
  const int c = p->dim;
  return tail_mma_enabled() && tail_eligible(p, out_dim) && (c == 32 || c == 64) && ((int64_t)p->batch * c) % pit::TM_CHUNK == 0;
}

// rows_per_unit counts 16-row tiles per CTA; one 64-column chunk per warp, at most 8 warps.  Equal shares for
// `ctas_per_sm` resident CTAs per SM (a last, partly filled round costs less than an unbalanced wave).
void plan_tail_mma_grid(const pit_problem_t* p, TallPlan& c, int ctas_per_sm) {
  const int chunks = p->batch * p->dim / pit::TM_CHUNK;
  c.threads = chunks <= 4 ? 128 : 256;
  const int tiles = (p->n_out + pit::TM_ROWS - 1) / pit::TM_ROWS;
  const int target = sm_count() * ctas_per_sm;
  const int per_cta = (tiles + target - 1) / target;
  c.rows_per_unit = per_cta;
  c.grid = (tiles + per_cta - 1) / per_cta;
  c.cpl = cpl_of(p->n_in);
}

TallPlan plan_tail_mma_fwd(const pit_problem_t* p, int out_dim) {
  TallPlan c{};
  if (!tail_mma_eligible(p, out_dim)) return c;
  plan_tail_mma_grid(p, c, 3);  // 80 registers per thread: three CTAs of 8 warps per SM
  c.smem = pit::tm_fwd_smem_bytes(p->n_head, p->n_in, p->dim, out_dim, c.threads);
  if (c.smem > (size_t)max_smem_optin() - 1024) return c;
  c.ok = true;
  return c;
}

TallPlan plan_tail_mma_bwd(const pit_problem_t* p, int out_dim) {
  TallPlan c{};
  if (!tail_mma_eligible(p, out_dim)) return c;
  plan_tail_mma_grid(p, c, 2);
  const int W = p->batch * p->dim;
  const size_t fixed = pit::tm_bwd_smem_bytes(p->n_head, p->n_in, W, p->dim, out_dim, 0, c.threads);
  const size_t per_slot = (size_t)p->n_head * W * 4 + 2;
  // two CTAs per SM if that leaves room for a useful slot set, else one
  for (int per_sm = 2; per_sm >= 1; --per_sm) {
    const size_t budget = ((size_t)max_smem_optin() + 1024) / per_sm - 2048;
    if (budget <= fixed) continue;
    int n = (int)((budget - fixed) / per_slot);
    if (n > 64) n = 64;
    if (n > p->n_in) n = p->n_in;
    if (n >= 12 || n == p->n_in) {
      c.n_slots = n;
      break;
    }
  }
  if (c.n_slots == 0) return c;
  c.smem = pit::tm_bwd_smem_bytes(p->n_head, p->n_in, W, p->dim, out_dim, c.n_slots, c.threads);
  if (c.smem > (size_t)max_smem_optin() - 1024) return c;
  c.ok = true;
  return c;
}

// Tile-plan variant (decoder_tail_plan.cuh): needs a plan from pit_tail_plan_*; hidden width 32 or 64.
bool tail_plan_eligible(const pit_problem_t* p, int out_dim, const pit_tail_plan_t* plan) {
  const int c = p->dim;
  return plan && plan->rec && plan->tile_off && plan->tile_cnt && plan->cand && plan->d2 && plan->n_tiles == (p->n_out + pit::TP_ROWS - 1) / pit::TP_ROWS &&
         tail_mma_enabled() && tail_eligible(p, out_dim) && (c == 32 || c == 64);
}

pit::TailPlanDev tail_plan_view(const pit_tail_plan_t* plan, int tiles_per_cta, int round) {
  pit::TailPlanDev V{};
  V.rec = static_cast<const float4*>(plan->rec);
  V.tile_off = plan->tile_off;
  V.tile_cnt = plan->tile_cnt;
  V.cand = plan->cand;
  V.d2 = plan->d2;
  V.n_tiles = plan->n_tiles;
  V.tiles_per_cta = tiles_per_cta;
  V.round = round;
  return V;
}

// One warp per sample (at most 8 warps); equal tile shares for `ctas_per_sm` resident CTAs per SM.
void plan_tail_plan_grid(const pit_problem_t* p, const pit_tail_plan_t* plan, TallPlan& c, int ctas_per_sm, int round) {
  const int warps = p->batch < pit::TP_MAX_WARPS ? p->batch : pit::TP_MAX_WARPS;
  c.threads = 32 * warps;
  const int target = sm_count() * ctas_per_sm;
  (void)round;  // equal shares matter more than whole rounds: a partly filled last round only idles the preparing warps
  const int per_cta = (plan->n_tiles + target - 1) / target;
  c.rows_per_unit = per_cta;
  c.grid = (plan->n_tiles + per_cta - 1) / per_cta;
}

TallPlan plan_tail_plan_fwd(const pit_problem_t* p, int out_dim, const pit_tail_plan_t* plan) {
  TallPlan c{};
  if (!tail_plan_eligible(p, out_dim, plan)) return c;
  const int warps = p->batch < pit::TP_MAX_WARPS ? p->batch : pit::TP_MAX_WARPS;
  c.l4 = warps < 4 ? warps : 4;  // tiles per round (kept in l4): 4 keeps three CTAs per SM at ~37 KB each and leaves most of the 228 KB to L1
  plan_tail_plan_grid(p, plan, c, pit::TP_FWD_CTAS, c.l4);
  c.smem = pit::tp_fwd_smem_bytes(p->n_head, p->n_in, p->dim, out_dim, c.l4);
  if (c.smem > (size_t)max_smem_optin() - 1024) return c;
  c.ok = true;
  return c;
}

TallPlan plan_tail_plan_bwd(const pit_problem_t* p, int out_dim, const pit_tail_plan_t* plan) {
  TallPlan c{};
  if (!tail_plan_eligible(p, out_dim, plan)) return c;
  // one CTA per SM, one warp per (sample, 32-column chunk), at most 16: a single set of tiles and slots per SM leaves
  // ~96 KB of the 228 KB to L1, which is what keeps the gathered Y rows of the current tiles on chip
  const int units = p->batch * (p->dim / pit::TP_CHUNK);
  const int warps = units < pit::TP_BWD_MAX_WARPS ? units : pit::TP_BWD_MAX_WARPS;
  c.threads = 32 * warps;
  const int target = sm_count();
  c.rows_per_unit = (plan->n_tiles + target - 1) / target;
  c.grid = (plan->n_tiles + c.rows_per_unit - 1) / c.rows_per_unit;
  const int W = p->batch * p->dim;
  const size_t fixed = pit::tp_bwd_smem_bytes(p->n_head, p->n_in, W, p->dim, out_dim, 0);
  const size_t per_slot = (size_t)p->n_head * W * 4 + 4;
  const size_t budget = 128 * 1024;
  int n = budget > fixed ? (int)((budget - fixed) / per_slot) : 0;
  if (n < 8) n = (int)(((size_t)max_smem_optin() - 2048 - fixed) / per_slot);  // wide batches: whatever fits
  if (n > 32) n = 32;
  if (n > p->n_in) n = p->n_in;
  if (n < 1) return c;
  c.n_slots = n;
  c.smem = pit::tp_bwd_smem_bytes(p->n_head, p->n_in, W, p->dim, out_dim, c.n_slots);
  if (c.smem > (size_t)max_smem_optin() - 1024) return c;
  c.ok = true;
  return c;
}

pit::TailParams tail_params(const pit_problem_t* p, const TallPlan& c, const float* mesh_out, const float* mesh_in,
                            const float* period, const float* y, const float* scale, const pit_rowstat_t* st,
                            const float* b1, const float* w2, const float* b2, int out_dim) {
  pit::TailParams P{};
  P.mesh_out = mesh_out;
  P.mesh_in = mesh_in;
  P.period = p->variant == PIT_EUCLID ? nullptr : period;
  P.y = y;
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
  P.C = p->dim;
  P.O = out_dim;
  P.sd = p->space_dim;
  P.lanes4 = c.lanes4;
  P.rows_per_unit = c.rows_per_unit;
  P.n_slots = c.n_slots;
  P.b1 = b1;
  P.w2 = w2;
  P.b2 = b2;
  return P;
}

// scale map of pit.py:48 and its derivative (libdevice sinf / tanf, no fast-math: same functions torch's kernels call)
__global__ void head_scale_fwd_kernel(const float* __restrict__ lmda, float* __restrict__ scale, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c = (float)(0.25 * 3.141592653589793 * (1 - 1e-7));
  scale[i] = tanf(__fmul_rn(c, __fadd_rn(1.0f, sinf(lmda[i]))));
}
__global__ void head_scale_bwd_kernel(const float* __restrict__ lmda, const float* __restrict__ scale,
                                      const float* __restrict__ d_scale, float* __restrict__ d_lmda, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c = (float)(0.25 * 3.141592653589793 * (1 - 1e-7));
  const float s = scale[i];
  d_lmda[i] = d_scale[i] * fmaf(s, s, 1.0f) * c * cosf(lmda[i]);
}

}  // namespace

extern "C" {

int pit_abi_version(void) { return PIT_ABI_VERSION; }
int pit_set_dense_precision(int32_t precision) {
  if (precision < PIT_DENSE_FP32 || precision > PIT_DENSE_BF16) return fail(PIT_ERR_ARG, "dense precision must be PIT_DENSE_FP32, _TF32 or _BF16");
  g_dense_precision.store(precision, std::memory_order_relaxed);
  return PIT_OK;
}
int pit_get_dense_precision(void) { return g_dense_precision.load(std::memory_order_relaxed); }
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
  size_t fwd = (f.n_split > 1 || (!p->mesh_batched && f.width <= pit::WIDE_MAX_WIDTH)) ? (size_t)f.items * f.width * sizeof(float) : 0;
  size_t bwd = (size_t)rows_total * p->n_head * 4 * sizeof(float);
  size_t need = fwd > bwd ? fwd : bwd;
  return need + 256;
}

static int rowstat_impl(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period, int32_t k_lo, int32_t k_hi,
                        float* v_min, float* v_lo, float* v_hi, int16_t* nbr_idx, float* nbr_d2, int32_t* nbr_cnt, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !v_min || !v_lo || !v_hi) return fail(PIT_ERR_ARG, "null pointer");
  if (p->variant != PIT_EUCLID && !period) return fail(PIT_ERR_ARG, "periodic variant needs the period pointer");
  if (k_lo < 0 || k_hi < k_lo || k_hi > k_lo + 1 || k_hi >= p->n_in)
    return fail(PIT_ERR_ARG, "ranks out of range: k_lo=%d k_hi=%d M=%d", k_lo, k_hi, p->n_in);
  if (nbr_idx && (p->n_in > 1024 || !nbr_d2 || !nbr_cnt)) return fail(PIT_ERR_ARG, "neighbour lists need M <= 1024 and all three arrays");
  pit::RowstatParams R{};
  R.mesh_out = mesh_out;
  R.mesh_in = mesh_in;
  R.period = p->variant == PIT_EUCLID ? nullptr : period;
  R.v_min = v_min;
  R.v_lo = v_lo;
  R.v_hi = v_hi;
  R.nbr_idx = nbr_idx;
  R.nbr_d2 = nbr_d2;
  R.nbr_cnt = nbr_cnt;
  R.N = p->n_out;
  R.M = p->n_in;
  R.sd = p->space_dim;
  R.mesh_batched = p->mesh_batched;
  R.rows_total = (p->mesh_batched ? p->batch : 1) * p->n_out;
  R.k_lo = k_lo;
  R.k_hi = k_hi;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PIT_CUDA(launch::rowstat(geo_of(p), R, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_rowstat(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                int32_t k_lo, int32_t k_hi, float* v_min, float* v_lo, float* v_hi, void* stream) {
  return rowstat_impl(p, mesh_out, mesh_in, period, k_lo, k_hi, v_min, v_lo, v_hi, nullptr, nullptr, nullptr, stream);
}

int pit_rowstat_lists(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period, int32_t k_lo, int32_t k_hi,
                      float* v_min, float* v_lo, float* v_hi, int16_t* nbr_idx, float* nbr_d2, int32_t* nbr_cnt, void* stream) {
  if (!nbr_idx) return fail(PIT_ERR_ARG, "null pointer");
  return rowstat_impl(p, mesh_out, mesh_in, period, k_lo, k_hi, v_min, v_lo, v_hi, nbr_idx, nbr_d2, nbr_cnt, stream);
}

static bool sample_tile_eligible(const pit_problem_t* p, const pit_rowstat_t* stat, int64_t ld_out, int64_t col_off, bool backward) {
  if (!p->mesh_batched || !stat->masked || p->n_in > 1024 || p->n_in < 1 || p->n_head > 2) return false;
  if (p->dim % 4 != 0 || p->dim > 256 || (ld_out % 4) || (col_off % 4)) return false;
  if (!stat->nbr_idx || !stat->nbr_d2 || !stat->nbr_cnt) return false;          // driven by the neighbour lists of pit_rowstat_lists
  // The tiles pay off when consecutive rows share their neighbours -- an upsampling stage over an ordered mesh (NACA decoder:
  // 11 271 rows on 728 columns).  Point clouds in arbitrary order (elasticity, N = M) touch most columns from any 64 rows: the
  // slots overflow and the generic warp-per-row kernels are faster (measured 2x).
  if ((int64_t)p->n_out < 4 * (int64_t)p->n_in) return false;
  if (p->batch > 65535) return false;
  return launch::sample_tile_slots(p->n_in, p->dim, backward, max_smem_optin()) >= 8;
}

static pit::SampleTileParams sample_tile_params(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                                         const float* values, const float* scale, const pit_rowstat_t* stat, bool backward) {
  pit::SampleTileParams T = {};
  (void)mesh_out, (void)mesh_in, (void)period;      // the distances come with the neighbour lists
  T.values = values, T.scale = scale;
  T.v_min = stat->v_min, T.v_lo = stat->v_lo, T.v_hi = stat->v_hi, T.weight = stat->weight;
  T.nbr_idx = stat->nbr_idx, T.nbr_d2 = stat->nbr_d2, T.nbr_cnt = stat->nbr_cnt;
  T.B = p->batch, T.H = p->n_head, T.N = p->n_out, T.M = p->n_in, T.D = p->dim, T.sd = p->space_dim;
  T.n_slots = launch::sample_tile_slots(p->n_in, p->dim, backward, max_smem_optin());
  return T;
}

int pit_posatt_forward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                       const float* values, const float* scale, const pit_rowstat_t* stat, float* out, int64_t ld_out,
                       int64_t col_off, int32_t copy_values, float* rowsum, void* workspace, size_t workspace_bytes,
                       const pit_tail_plan_t* column_plan, void* stream) {
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
  if (copy_values) {  // first D columns of the concat output (pit.py:44)
    PIT_CUDA(cudaMemcpy2DAsync(out, (size_t)ld_out * sizeof(float), values, (size_t)p->dim * sizeof(float),
                               (size_t)p->dim * sizeof(float), (size_t)p->batch * p->n_in, cudaMemcpyDeviceToDevice, st));
  }
  if (!copy_values && aligned16(values) && aligned16(out) && sample_tile_eligible(p, stat, ld_out, col_off, false)) {
    pit::SampleTileParams T = sample_tile_params(p, mesh_out, mesh_in, period, values, scale, stat, false);
    T.out = out, T.ld_out = ld_out, T.col_off = col_off, T.rowsum = rowsum;
    PIT_CUDA(launch::sample_tile(false, T, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PIT_OK;
  }
  if (dense_eligible(p, stat)) {
    pit::DenseParams Dn = dense_params(p, mesh_out, mesh_in, period, scale, stat, pit::DENSE_FWD);
    Dn.b_src = values;
    Dn.b_kstride = p->dim;
    Dn.b_bstride = (int64_t)p->n_in * p->dim;
    Dn.out = out;
    Dn.ld_out = ld_out;
    Dn.col_off = col_off;
    Dn.rowsum_out = rowsum;
    PIT_CUDA(dense_launch(pit::DENSE_FWD, geo_of(p), p, Dn, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PIT_OK;
  }
  const TallPlan plan = plan_tall_fwd(p);
  if (plan.ok) {
    pit::TallParams C = tall_params(p, plan, mesh_out, mesh_in, period, values, scale, stat);
    C.out = out;
    C.ld_out = ld_out;
    C.col_off = col_off;
    C.rowsum = rowsum;
    PIT_CUDA(launch::tall_forward(geo_of(p), plan, C, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PIT_OK;
  }
  pit::AttnParams P = base_params(p, mesh_out, mesh_in, period, values, scale, stat);
  P.out = out;
  P.ld_out = ld_out;
  P.col_off = col_off;
  P.rowsum = rowsum;
  P.split_len = s.split_len;
  const WidePlan wide = plan_wide(p, stat);
  if (wide.ok) {
    const size_t need = (size_t)s.items * s.width * sizeof(float);
    if (!workspace || workspace_bytes < need) return fail(PIT_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    P.partial = static_cast<float*>(workspace);
    PIT_CUDA(cudaMemsetAsync(P.partial, 0, need, st));
    PIT_CUDA(cudaMemsetAsync(rowsum, 0, (size_t)s.items * sizeof(float), st));
    pit::WideParams W = wide_params(p, mesh_out, mesh_in, period, values, scale, stat);
    W.partial = P.partial;
    W.rowsum = rowsum;
    if (column_plan_ok(p, column_plan))
      PIT_CUDA(launch::wide_plan_forward(column_plan_grid(column_plan), W, tail_plan_view(column_plan, 0, 0), st));
    else
      PIT_CUDA(launch::wide_forward(geo_of(p), wide, W, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const int64_t total = s.items * s.width;
    PIT_CUDA(launch::local_forward_finalize(total, P, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PIT_OK;
  }
  if (s.n_split > 1) {
    const size_t need = (size_t)s.items * s.width * sizeof(float);
    if (!workspace || workspace_bytes < need) return fail(PIT_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    P.partial = static_cast<float*>(workspace);
    PIT_CUDA(cudaMemsetAsync(P.partial, 0, need, st));
    PIT_CUDA(cudaMemsetAsync(rowsum, 0, (size_t)s.items * sizeof(float), st));
  }
  const dim3 grid((unsigned)((s.items + pit::WARPS_PER_BLOCK - 1) / pit::WARPS_PER_BLOCK), s.chunks, s.n_split);
  PIT_CUDA(launch::local_forward(geo_of(p), s.vec, s.a, grid, P, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (s.n_split > 1) {
    const int64_t total = s.items * s.width;
    PIT_CUDA(launch::local_forward_finalize(total, P, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  return PIT_OK;
}

int pit_posatt_backward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                        const float* values, const float* scale, const pit_rowstat_t* stat, const float* rowsum,
                        const float* d_out, int64_t ld_out, int64_t col_off, int32_t accumulate_concat, float* d_values,
                        float* d_scale, void* workspace, size_t workspace_bytes, const pit_tail_plan_t* column_plan, void* stream) {
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

  // Tall path (shared meshes, M <= 1024, H <= 2): one pass over d_out gives the scale gradient and, for a
  // masked cross stage, the value gradient too.  A dense self stage keeps the value gradient on the
  // column-owner kernel below (every column is touched by every row, slots would not help).
  bool values_done = d_values == nullptr, scale_done = d_scale == nullptr;
  if (!accumulate_concat && (d_values || d_scale) && aligned16(values) && aligned16(d_out) && (!d_values || aligned16(d_values)) &&
      sample_tile_eligible(p, stat, ld_out, col_off, true)) {
    // per-sample meshes, masked: one pass per (sample, 64-row tile) gives both gradients
    pit::SampleTileParams T = sample_tile_params(p, mesh_out, mesh_in, period, values, scale, stat, true);
    T.d_out = d_out, T.ld_out = ld_out, T.col_off = col_off, T.d_values = d_values, T.d_scale = d_scale;
    if (d_values) PIT_CUDA(cudaMemsetAsync(d_values, 0, (size_t)p->batch * p->n_in * p->dim * sizeof(float), st));
    if (d_scale) PIT_CUDA(cudaMemsetAsync(d_scale, 0, (size_t)p->n_head * sizeof(float), st));
    PIT_CUDA(launch::sample_tile(true, T, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PIT_OK;
  }
  if (dense_eligible(p, stat)) {
    pit::DenseParams Ds{}, Dv{};
    if (d_scale) {
      Ds = dense_params(p, mesh_out, mesh_in, period, scale, stat, pit::DENSE_DSCALE);
      Ds.rowsum = rowsum;
      Ds.b_src = values;
      Ds.b_kstride = p->dim;
      Ds.b_bstride = (int64_t)p->n_in * p->dim;
      Ds.d_out = d_out;
      Ds.ld_out = ld_out;
      Ds.col_off = col_off;
      Ds.d_scale = d_scale;
      PIT_CUDA(cudaMemsetAsync(d_scale, 0, (size_t)p->n_head * sizeof(float), st));
    }
    if (d_values) {
      Dv = dense_params(p, mesh_out, mesh_in, period, scale, stat, pit::DENSE_DVALUES);
      Dv.rowsum = rowsum;
      Dv.b_src = d_out;
      Dv.b_kstride = ld_out;
      Dv.b_bstride = (int64_t)p->n_out * ld_out;
      Dv.b_hstride = p->dim;
      Dv.b_off = col_off;
      Dv.d_out = d_out;
      Dv.ld_out = ld_out;
      Dv.col_off = col_off;
      Dv.d_values = d_values;
      Dv.add_concat = accumulate_concat;
      Dv.split_heads = dense_split_heads(p) ? 1 : 0;
      if (Dv.split_heads) PIT_CUDA(cudaMemsetAsync(d_values, 0, (size_t)p->batch * p->n_in * p->dim * sizeof(float), st));
    }
    int nv_s = 0, nv_v = 0;
    const dim3 gs = d_scale ? dense_grid(pit::DENSE_DSCALE, p, Ds, &nv_s) : dim3(0, 0, 0);
    const dim3 gv = d_values ? dense_grid(pit::DENSE_DVALUES, p, Dv, &nv_v) : dim3(0, 0, 0);
    // small stage: neither grid fills the chip -> both gradient modes side by side in one launch
    if (d_scale && d_values && nv_s == 64 && nv_v == 64 &&
        (int64_t)gs.x * gs.y * gs.z + (int64_t)gv.x * gv.y * gv.z <= sm_count()) {
      PIT_CUDA(launch::dense_bwd_pair(geo, Ds, gs, Dv, gv, st));
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
      if (d_scale) {
        PIT_CUDA(launch::dense(pit::DENSE_DSCALE, geo, nv_s, gs, Ds, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
      }
      if (d_values) {
        PIT_CUDA(launch::dense(pit::DENSE_DVALUES, geo, nv_v, gv, Dv, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
      }
    }
    scale_done = values_done = true;
  }
  if (!(values_done && scale_done)) {
    const bool fuse_values = d_values && stat->masked && !accumulate_concat;
    const TallPlan plan = plan_tall_bwd(p, fuse_values);
    if (plan.ok && (d_scale || fuse_values)) {
      pit::TallParams C = tall_params(p, plan, mesh_out, mesh_in, period, values, scale, stat);
      C.rowsum = const_cast<float*>(rowsum);
      C.d_out = d_out;
      C.ld_out = ld_out;
      C.col_off = col_off;
      C.d_values = fuse_values ? d_values : nullptr;
      C.d_scale = d_scale;
      if (d_scale) PIT_CUDA(cudaMemsetAsync(d_scale, 0, (size_t)p->n_head * sizeof(float), st));
      if (fuse_values) PIT_CUDA(cudaMemsetAsync(d_values, 0, (size_t)p->batch * p->n_in * p->dim * sizeof(float), st));
      PIT_CUDA(launch::tall_backward(geo, plan, C, fuse_values, st));
      g_launches.fetch_add(1, std::memory_order_relaxed);
      scale_done = true;
      if (fuse_values) values_done = true;
    }
  }

  pit::AttnParams P = base_params(p, mesh_out, mesh_in, period, values, scale, stat);
  P.rowsum = const_cast<float*>(rowsum);
  P.d_out = d_out;
  P.ld_out = ld_out;
  P.col_off = col_off;
  P.d_values = d_values;
  P.add_concat = accumulate_concat;

  if (!scale_done) {
    Shape s = make_shape(p, rows_total * p->n_head, p->n_in, true);
    const size_t need = (size_t)s.items * 4 * sizeof(float);
    if (!workspace || workspace_bytes < need) return fail(PIT_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
    P.dscale_terms = static_cast<float*>(workspace);
    float* rows = P.dscale_terms + (size_t)s.items * 3;
    P.split_len = s.split_len;
    PIT_CUDA(cudaMemsetAsync(P.dscale_terms, 0, (size_t)s.items * 3 * sizeof(float), st));
    const WidePlan wide = plan_wide(p, stat);
    if (wide.ok) {
      pit::WideParams W = wide_params(p, mesh_out, mesh_in, period, values, scale, stat);
      W.d_out = d_out;
      W.ld_out = ld_out;
      W.col_off = col_off;
      W.dscale_terms = P.dscale_terms;
      if (column_plan_ok(p, column_plan))
        PIT_CUDA(launch::wide_plan_dscale(column_plan_grid(column_plan), W, tail_plan_view(column_plan, 0, 0), st));
      else
        PIT_CUDA(launch::wide_dscale(geo, wide, W, st));
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
      const dim3 grid((unsigned)((s.items + pit::WARPS_PER_BLOCK - 1) / pit::WARPS_PER_BLOCK), s.chunks, s.n_split);
      PIT_CUDA(launch::local_dscale(geo, s.vec, s.a, grid, P, st));
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    PIT_CUDA(launch::local_dscale_finalize(s.items, P, rows, st));
    PIT_CUDA(launch::reduce_scale_rows(rows, rows_total, p->n_head, d_scale, st));
    g_launches.fetch_add(2, std::memory_order_relaxed);
  }
  if (!values_done) {
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
    PIT_CUDA(launch::local_dvalues(geo, s.vec, s.a, grid, P, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  return PIT_OK;
}


int pit_head_scale_forward(const float* lmda, float* scale, int32_t n, void* stream) {
  if (!lmda || !scale || n < 1) return fail(PIT_ERR_ARG, "head scale: bad arguments");
  head_scale_fwd_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(lmda, scale, n);
  PIT_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_head_scale_backward(const float* lmda, const float* scale, const float* d_scale, float* d_lmda, int32_t n, void* stream) {
  if (!lmda || !scale || !d_scale || !d_lmda || n < 1) return fail(PIT_ERR_ARG, "head scale: bad arguments");
  head_scale_bwd_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(lmda, scale, d_scale, d_lmda, n);
  PIT_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_rel_lp_supported(int32_t batch, int64_t length, int32_t out_dim, int32_t p) {
  return batch >= 1 && batch <= 65535 && length >= 1 && out_dim >= 1 && out_dim <= pit::LOSS_MAX_OUT && (p == 1 || p == 2) ? 1 : 0;
}

namespace {
// blocks along one sample: enough to fill the chip together with the batch, total thread count a multiple of out_dim
int rel_lp_grid(int batch, int64_t length, int out_dim) {
  const int64_t n = length * out_dim;
  int64_t blocks = (n + pit::LOSS_THREADS * 8 - 1) / (pit::LOSS_THREADS * 8);
  const int64_t cap = ((int64_t)sm_count() * 8 + batch - 1) / batch;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  while ((blocks * pit::LOSS_THREADS) % out_dim != 0) ++blocks;   // out_dim 3: a multiple of three blocks
  return (int)blocks;
}
}  // namespace

int pit_rel_lp_forward(const float* truth, const float* pred, int32_t batch, int64_t length, int32_t out_dim, int32_t p,
                       float* norms, float* loss, void* stream) {
  if (!truth || !pred || !norms || !loss || !pit_rel_lp_supported(batch, length, out_dim, p)) return fail(PIT_ERR_ARG, "rel_lp: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pit::LossParams P{};
  P.truth = truth, P.pred = pred, P.sums = norms, P.loss = loss, P.L = length, P.B = batch, P.O = out_dim, P.p = p;
  PIT_CUDA(cudaMemsetAsync(norms, 0, (size_t)batch * out_dim * 2 * sizeof(float), st));
  PIT_CUDA(launch::rel_lp(0, P, rel_lp_grid(batch, length, out_dim), st));
  PIT_CUDA(launch::rel_lp(1, P, 1, st));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_rel_lp_backward(const float* truth, const float* pred, const float* norms, const float* d_loss, int32_t batch,
                        int64_t length, int32_t out_dim, int32_t p, float* d_pred, void* stream) {
  if (!truth || !pred || !norms || !d_loss || !d_pred || !pit_rel_lp_supported(batch, length, out_dim, p))
    return fail(PIT_ERR_ARG, "rel_lp: bad arguments");
  pit::LossParams P{};
  P.truth = truth, P.pred = pred, P.sums = const_cast<float*>(norms), P.d_loss = d_loss, P.d_pred = d_pred, P.L = length, P.B = batch,
  P.O = out_dim, P.p = p;
  PIT_CUDA(launch::rel_lp(2, P, rel_lp_grid(batch, length, out_dim), static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_bias_act_supported(int64_t rows, int32_t cols) {
  return rows >= 1 && cols >= 4 && cols % 4 == 0 && pit::EPI_THREADS % (cols / 4) == 0 ? 1 : 0;
}

namespace {
int bias_act_grid(int64_t rows, int cols4) {
  const int64_t blocks = (rows * cols4 + pit::EPI_THREADS - 1) / pit::EPI_THREADS;
  const int64_t cap = (int64_t)sm_count() * 4;
  return (int)(blocks < cap ? blocks : cap);
}
}  // namespace

int pit_bias_act_forward(const float* z, const float* bias, float* out, int64_t rows, int32_t cols, int32_t apply_gelu, void* stream) {
  if (!z || !bias || !out || !pit_bias_act_supported(rows, cols)) return fail(PIT_ERR_ARG, "bias_act: bad arguments (cols must be a multiple of 4 dividing 1024)");
  if (!aligned16(z) || !aligned16(bias) || !aligned16(out)) return fail(PIT_ERR_ARG, "bias_act: pointers must be 16-byte aligned");
  pit::EpiParams P{};
  P.z = z, P.bias = bias, P.out = out, P.rows = rows, P.cols4 = cols / 4, P.gelu = apply_gelu;
  PIT_CUDA(launch::bias_act(false, P, bias_act_grid(rows, P.cols4), static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_bias_act_backward(const float* z, const float* bias, const float* d_out, float* d_z, float* d_bias, int64_t rows,
                          int32_t cols, int32_t apply_gelu, void* stream) {
  if (!z || !bias || !d_out || !d_z || !d_bias || !pit_bias_act_supported(rows, cols))
    return fail(PIT_ERR_ARG, "bias_act: bad arguments (cols must be a multiple of 4 dividing 1024)");
  if (!aligned16(z) || !aligned16(bias) || !aligned16(d_out) || !aligned16(d_z) || !aligned16(d_bias))
    return fail(PIT_ERR_ARG, "bias_act: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pit::EpiParams P{};
  P.z = z, P.bias = bias, P.d_out = d_out, P.out = d_z, P.d_bias = d_bias, P.rows = rows, P.cols4 = cols / 4, P.gelu = apply_gelu;
  PIT_CUDA(cudaMemsetAsync(d_bias, 0, (size_t)cols * sizeof(float), st));
  PIT_CUDA(launch::bias_act(true, P, bias_act_grid(rows, P.cols4), st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

namespace {
int check_tail_plan_args(const pit_problem_t* p, int side, const float* mesh_out, const float* mesh_in, const float* period,
                         const pit_rowstat_t* stat, void* workspace, size_t workspace_bytes) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !workspace) return fail(PIT_ERR_ARG, "null pointer");
  if (side != PIT_PLAN_ROWS && side != PIT_PLAN_COLUMNS) return fail(PIT_ERR_ARG, "tile plan: unknown side %d", side);
  if (int rc = check_stat(p, stat, period)) return rc;
  const int small = side == PIT_PLAN_ROWS ? p->n_in : p->n_out;
  if (p->mesh_batched || small > pit::TALL_MAX_M)
    return fail(PIT_ERR_ARG, "tile plan: needs shared meshes with at most %d points on the candidate side", pit::TALL_MAX_M);
  const size_t need = launch::tail_plan_workspace_bytes(side == PIT_PLAN_ROWS ? p->n_out : p->n_in);
  if (workspace_bytes < need) return fail(PIT_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
  return PIT_OK;
}

pit::PlanBuildParams plan_build_params(const pit_problem_t* p, int side, const float* mesh_out, const float* mesh_in, const float* period,
                                       const pit_rowstat_t* stat) {
  pit::PlanBuildParams B{};
  B.transposed = side == PIT_PLAN_COLUMNS;
  B.mesh_out = B.transposed ? mesh_in : mesh_out;   // tiled side
  B.mesh_in = B.transposed ? mesh_out : mesh_in;    // candidate side
  B.period = p->variant == PIT_EUCLID ? nullptr : period;
  B.v_min = stat->v_min;
  B.v_lo = stat->v_lo;
  B.v_hi = stat->v_hi;
  B.masked = stat->masked;
  B.N = B.transposed ? p->n_in : p->n_out;
  B.M = B.transposed ? p->n_out : p->n_in;
  B.sd = p->space_dim;
  B.n_tiles = (B.N + pit::TP_ROWS - 1) / pit::TP_ROWS;
  return B;
}
}  // namespace

size_t pit_tail_plan_workspace_bytes(const pit_problem_t* p, int32_t side) {
  if (check_problem(p) != PIT_OK) return 0;
  return launch::tail_plan_workspace_bytes(side == PIT_PLAN_COLUMNS ? p->n_in : p->n_out);
}

int pit_tail_plan_rows(const pit_problem_t* p, int32_t side, const float* mesh_out, const float* mesh_in, const float* period,
                       const pit_rowstat_t* stat, int32_t* tile_off, int32_t* tile_cnt, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (int rc = check_tail_plan_args(p, side, mesh_out, mesh_in, period, stat, workspace, workspace_bytes)) return rc;
  if (!tile_off || !tile_cnt) return fail(PIT_ERR_ARG, "null pointer");
  const pit::PlanBuildParams B0 = plan_build_params(p, side, mesh_out, mesh_in, period, stat);
  PIT_CUDA(launch::tail_plan_rows(geo_of(p), cpl_of(B0.M), B0, tile_off, tile_cnt, workspace,
                                  static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_tail_plan_fill(const pit_problem_t* p, int32_t side, const float* mesh_out, const float* mesh_in, const float* period,
                       const pit_rowstat_t* stat, const int32_t* tile_off, void* rec, int16_t* cand, float* d2,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_tail_plan_args(p, side, mesh_out, mesh_in, period, stat, workspace, workspace_bytes)) return rc;
  if (!tile_off || !rec || !cand || !d2) return fail(PIT_ERR_ARG, "null pointer");
  if (!aligned16(rec)) return fail(PIT_ERR_ARG, "tail plan: rec must be 16-byte aligned");
  pit::PlanBuildParams B = plan_build_params(p, side, mesh_out, mesh_in, period, stat);
  B.tile_off = tile_off;
  B.rec = static_cast<float4*>(rec);
  B.cand = cand;
  B.d2 = d2;
  PIT_CUDA(launch::tail_plan_fill(geo_of(p), cpl_of(B.M), B, workspace, static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_decoder_tail_supported(const pit_problem_t* p, int32_t out_dim) {
  if (check_problem(p) != PIT_OK) return 0;
  if (!tail_eligible(p, out_dim)) return 0;
  return plan_tail_fwd(p).ok && plan_tail_bwd(p, out_dim).ok ? 1 : 0;
}

int pit_decoder_tail_forward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                             const float* y, const float* scale, const pit_rowstat_t* stat, const float* b1, const float* w2,
                             const float* b2, int32_t out_dim, float* out, float* rowsum, const pit_tail_plan_t* tile_plan,
                             void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !y || !scale || !b1 || !w2 || !b2 || !out || !rowsum) return fail(PIT_ERR_ARG, "null pointer");
  if (int rc = check_stat(p, stat, period)) return rc;
  if (!tail_eligible(p, out_dim)) return fail(PIT_ERR_ARG, "decoder tail: unsupported configuration (see pit_decoder_tail_supported)");
  if (!aligned16(y) || !aligned16(b1) || !aligned16(w2)) return fail(PIT_ERR_ARG, "decoder tail: y, b1, w2 must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TallPlan tiled = plan_tail_plan_fwd(p, out_dim, tile_plan);
  if (tiled.ok) {
    pit::TailParams P = tail_params(p, tiled, mesh_out, mesh_in, period, y, scale, stat, b1, w2, b2, out_dim);
    P.out = out;
    P.rowsum = rowsum;
    PIT_CUDA(launch::tail_plan_forward(geo_of(p), tiled, P, tail_plan_view(tile_plan, tiled.rows_per_unit, tiled.l4), st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return PIT_OK;
  }
  const TallPlan mma = plan_tail_mma_fwd(p, out_dim);
  const TallPlan plan = mma.ok ? mma : plan_tail_fwd(p);
  if (!plan.ok) return fail(PIT_ERR_ARG, "decoder tail: no launch plan");
  pit::TailParams P = tail_params(p, plan, mesh_out, mesh_in, period, y, scale, stat, b1, w2, b2, out_dim);
  P.out = out;
  P.rowsum = rowsum;
  if (mma.ok)
    PIT_CUDA(launch::tail_mma_forward(geo_of(p), plan, P, st));
  else
    PIT_CUDA(launch::tail_forward(geo_of(p), plan, P, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_decoder_tail_backward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                              const float* y, const float* scale, const pit_rowstat_t* stat, const float* b1, const float* w2,
                              const float* b2, int32_t out_dim, const float* rowsum, const float* d_out, float* d_y,
                              float* d_scale, float* d_b1, float* d_w2, float* d_b2, const pit_tail_plan_t* tile_plan,
                              void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !y || !scale || !b1 || !w2 || !b2 || !rowsum || !d_out || !d_y || !d_scale || !d_b1 || !d_w2 || !d_b2)
    return fail(PIT_ERR_ARG, "null pointer");
  if (int rc = check_stat(p, stat, period)) return rc;
  if (!tail_eligible(p, out_dim)) return fail(PIT_ERR_ARG, "decoder tail: unsupported configuration (see pit_decoder_tail_supported)");
  if (!aligned16(y) || !aligned16(b1) || !aligned16(w2) || !aligned16(d_y) || !aligned16(d_b1) || !aligned16(d_w2))
    return fail(PIT_ERR_ARG, "decoder tail: y, b1, w2 and their gradients must be 16-byte aligned");
  const TallPlan tiled = plan_tail_plan_bwd(p, out_dim, tile_plan);
  const TallPlan mma = tiled.ok ? TallPlan{} : plan_tail_mma_bwd(p, out_dim);
  const TallPlan plan = tiled.ok ? tiled : (mma.ok ? mma : plan_tail_bwd(p, out_dim));
  if (!plan.ok) return fail(PIT_ERR_ARG, "decoder tail: no launch plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pit::TailParams P = tail_params(p, plan, mesh_out, mesh_in, period, y, scale, stat, b1, w2, b2, out_dim);
  P.rowsum = const_cast<float*>(rowsum);
  P.d_out = d_out;
  P.d_y = d_y;
  P.d_scale = d_scale;
  P.d_b1 = d_b1;
  P.d_w2 = d_w2;
  P.d_b2 = d_b2;
  const size_t c = (size_t)p->dim;
  {
    // gradient buffers that sit back to back in d_y, d_b1, d_w2, d_b2, d_scale order (each padded to a multiple of four floats, as
    // the Python host side allocates them) are cleared with one memset node instead of five
    auto pad = [](size_t n) { return (n + 3) / 4 * 4; };
    const size_t n_y = (size_t)p->batch * p->n_in * p->n_head * c, n_w2 = (size_t)out_dim * c;
    if (d_b1 == d_y + pad(n_y) && d_w2 == d_b1 + pad(c) && d_b2 == d_w2 + pad(n_w2) && d_scale == d_b2 + pad((size_t)out_dim)) {
      PIT_CUDA(cudaMemsetAsync(d_y, 0, (pad(n_y) + pad(c) + pad(n_w2) + pad((size_t)out_dim) + (size_t)p->n_head) * sizeof(float), st));
    } else {
      PIT_CUDA(cudaMemsetAsync(d_y, 0, n_y * sizeof(float), st));
      PIT_CUDA(cudaMemsetAsync(d_scale, 0, (size_t)p->n_head * sizeof(float), st));
      PIT_CUDA(cudaMemsetAsync(d_b1, 0, c * sizeof(float), st));
      PIT_CUDA(cudaMemsetAsync(d_w2, 0, n_w2 * sizeof(float), st));
      PIT_CUDA(cudaMemsetAsync(d_b2, 0, (size_t)out_dim * sizeof(float), st));
    }
  }
  if (tiled.ok)
    PIT_CUDA(launch::tail_plan_backward(geo_of(p), plan, P, tail_plan_view(tile_plan, plan.rows_per_unit, pit::TP_BWD_ROUND), st));
  else if (mma.ok)
    PIT_CUDA(launch::tail_mma_backward(geo_of(p), plan, P, st));
  else
    PIT_CUDA(launch::tail_backward(geo_of(p), plan, P, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// fused processor (pit.py:114-122)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
bool processor_eligible(const pit_problem_t* p, int32_t n_blocks) {
  const int tr = launch::PROC_TILE_ROWS;
  if (p->mesh_batched || p->n_out != p->n_in) return false;
  if (n_blocks < 1 || n_blocks > pit::PB_MAX_BLOCKS) return false;
  if (p->n_head < 1 || p->n_head > 2 || (p->dim != 32 && p->dim != 64)) return false;
  if (p->n_out % tr != 0 || p->n_out / tr < 1 || p->n_out / tr > 8) return false;   // portable cluster size
  if (p->batch < 1 || p->batch > 65535) return false;
  return launch::processor_smem_bytes(p->dim, p->n_head, p->n_out) <= (size_t)max_smem_optin();
}

pit::ProcParams processor_params(const pit_problem_t* p, int32_t n_blocks, const float* mesh, const float* period, const float* x0,
                                 const float* scale, const pit_processor_block_t* blocks, float* saved) {
  pit::ProcParams P = {};
  P.geo = geo_of(p);
  P.sd = p->space_dim;
  P.B = p->batch;
  P.N = p->n_out;
  P.n_blocks = n_blocks;
  P.mesh = mesh;
  P.period = period;
  P.x0 = x0;
  P.scale = scale;
  P.saved = saved;
  for (int k = 0; k < n_blocks; ++k) P.w[k] = pit::ProcWeights{blocks[k].w1, blocks[k].b1, blocks[k].w2, blocks[k].b2};
  return P;
}

int check_processor(const pit_problem_t* p, int32_t n_blocks, const float* mesh, const float* period, const float* x0, const float* scale,
                    const pit_processor_block_t* blocks, const float* saved) {
  if (int rc = check_problem(p)) return rc;
  if (!processor_eligible(p, n_blocks)) return fail(PIT_ERR_ARG, "processor: unsupported configuration (see pit_processor_supported)");
  if (!mesh || !x0 || !scale || !blocks || !saved) return fail(PIT_ERR_ARG, "null pointer");
  if (p->variant != PIT_EUCLID && !period) return fail(PIT_ERR_ARG, "periodic variant needs the wrap length");
  if (!aligned16(x0) || !aligned16(saved)) return fail(PIT_ERR_ARG, "processor: x0 and saved must be 16-byte aligned");
  for (int k = 0; k < n_blocks; ++k) {
    if (!blocks[k].w1 || !blocks[k].b1 || !blocks[k].w2 || !blocks[k].b2) return fail(PIT_ERR_ARG, "processor: null weight pointer");
    if ((reinterpret_cast<uintptr_t>(blocks[k].b1) | reinterpret_cast<uintptr_t>(blocks[k].b2)) & 7u)
      return fail(PIT_ERR_ARG, "processor: biases must be 8-byte aligned");
  }
  return PIT_OK;
}
}  // namespace

int pit_processor_supported(const pit_problem_t* p, int32_t n_blocks) {
  if (check_problem(p) != PIT_OK) return 0;
  return processor_eligible(p, n_blocks) ? 1 : 0;
}

size_t pit_processor_saved_floats(const pit_problem_t* p, int32_t n_blocks) {
  if (check_problem(p) != PIT_OK || n_blocks < 1) return 0;
  return (size_t)pit::proc_saved_layout(p->batch, p->n_out, p->n_head, p->dim).stride * (size_t)n_blocks;
}

size_t pit_processor_grad_floats(const pit_problem_t* p, int32_t n_blocks) {
  if (check_problem(p) != PIT_OK || n_blocks < 1) return 0;
  const size_t d = (size_t)p->dim, cw = d * (1 + (size_t)p->n_head);
  return (size_t)n_blocks * (d * cw + d + d * d + d + (size_t)p->n_head);
}

size_t pit_processor_scratch_floats(const pit_problem_t* p) {
  if (check_problem(p) != PIT_OK) return 0;
  return 2 * (size_t)p->batch * p->n_out * p->n_head * p->dim;
}

int pit_processor_forward(const pit_problem_t* p, int32_t n_blocks, const float* mesh, const float* period, const float* x0,
                          const float* scale, const pit_processor_block_t* blocks, int32_t linear_3xtf32, float* saved, float* out,
                          void* stream) {
  if (int rc = check_processor(p, n_blocks, mesh, period, x0, scale, blocks, saved)) return rc;
  if (!out) return fail(PIT_ERR_ARG, "null pointer");
  pit::ProcParams P = processor_params(p, n_blocks, mesh, period, x0, scale, blocks, saved);
  P.out = out;
  PIT_CUDA(launch::processor(false, p->dim, p->n_head, linear_3xtf32 != 0, P, static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_processor_backward(const pit_problem_t* p, int32_t n_blocks, const float* mesh, const float* period, const float* x0,
                           const float* scale, const pit_processor_block_t* blocks, int32_t linear_3xtf32, const float* saved,
                           const float* d_out, float* d_x0, float* grads, float* scratch, void* stream) {
  if (int rc = check_processor(p, n_blocks, mesh, period, x0, scale, blocks, saved)) return rc;
  if (!d_out || !d_x0 || !grads || !scratch) return fail(PIT_ERR_ARG, "null pointer");
  if (!aligned16(scratch) || (reinterpret_cast<uintptr_t>(grads) & 7u) || (reinterpret_cast<uintptr_t>(d_x0) & 7u))
    return fail(PIT_ERR_ARG, "processor: scratch must be 16-byte, grads and d_x0 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pit::ProcParams P = processor_params(p, n_blocks, mesh, period, x0, scale, blocks, const_cast<float*>(saved));
  P.d_out = d_out;
  P.d_x0 = d_x0;
  P.scratch = scratch;
  const size_t d = (size_t)p->dim, cw = d * (1 + (size_t)p->n_head);
  float* q = grads;
  for (int k = 0; k < n_blocks; ++k) {
    P.g[k].d_w1 = q, q += d * cw;
    P.g[k].d_b1 = q, q += d;
    P.g[k].d_w2 = q, q += d * d;
    P.g[k].d_b2 = q, q += d;
  }
  P.d_scale = q;
  PIT_CUDA(cudaMemsetAsync(grads, 0, pit_processor_grad_floats(p, n_blocks) * sizeof(float), st));
  PIT_CUDA(launch::processor(true, p->dim, p->n_head, linear_3xtf32 != 0, P, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// coordinate gradients
// ---------------------------------------------------------------------------------------------------------------------
int pit_posatt_backward_coords(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                               const float* values, const float* scale, const pit_rowstat_t* stat, const float* rowsum,
                               const float* d_out, int64_t ld_out, int64_t col_off, float* d_mesh_out, float* d_mesh_in,
                               float* d_period, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (!mesh_out || !mesh_in || !values || !scale || !rowsum || !d_out || !d_mesh_out || !d_mesh_in) return fail(PIT_ERR_ARG, "null pointer");
  if (int rc = check_stat(p, stat, period)) return rc;
  if (p->variant != PIT_EUCLID && !d_period) return fail(PIT_ERR_ARG, "periodic variant: d_period is required");
  if ((int64_t)p->batch * p->n_out > (int64_t)2147483647 * pit::CG_WARPS) return fail(PIT_ERR_ARG, "too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pit::CoordGradParams P = {};
  P.mesh_out = mesh_out, P.mesh_in = mesh_in, P.period = period, P.values = values, P.scale = scale;
  P.v_min = stat->v_min, P.v_lo = stat->v_lo, P.v_hi = stat->v_hi, P.rowsum = rowsum, P.d_out = d_out;
  P.weight = stat->weight, P.masked = stat->masked;
  P.B = p->batch, P.H = p->n_head, P.N = p->n_out, P.M = p->n_in, P.D = p->dim, P.sd = p->space_dim, P.mesh_batched = p->mesh_batched;
  P.ld_out = ld_out, P.col_off = col_off;
  P.d_mesh_out = d_mesh_out, P.d_mesh_in = d_mesh_in, P.d_period = p->variant == PIT_EUCLID ? nullptr : d_period;
  const size_t copies = p->mesh_batched ? (size_t)p->batch : 1;
  PIT_CUDA(cudaMemsetAsync(d_mesh_out, 0, copies * p->n_out * p->space_dim * sizeof(float), st));
  if (d_mesh_in != d_mesh_out) PIT_CUDA(cudaMemsetAsync(d_mesh_in, 0, copies * p->n_in * p->space_dim * sizeof(float), st));
  if (d_period) PIT_CUDA(cudaMemsetAsync(d_period, 0, sizeof(float), st));
  PIT_CUDA(launch::coord_gradient(geo_of(p), P, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// gradient all-reduce over peer memory + Adam
// ---------------------------------------------------------------------------------------------------------------------
size_t pit_allreduce_adam_region_floats(int64_t total) {
  if (total < 1) return 0;
  const int64_t stride = (total + 3) / 4 * 4;
  return (size_t)(pit::ARA_FLAG_WORDS + 2 * stride);
}

int pit_allreduce_adam(const pit_allreduce_adam_t* a, void* stream) {
  if (!a) return fail(PIT_ERR_ARG, "null pointer");
  if (a->world < 1 || a->world > pit::ARA_MAX_WORLD || a->rank < 0 || a->rank >= a->world) return fail(PIT_ERR_ARG, "allreduce_adam: bad world / rank");
  if (a->n_tensors < 1 || a->n_tensors > pit::ARA_MAX_TENSORS) return fail(PIT_ERR_ARG, "allreduce_adam: 1..%d gradient tensors", pit::ARA_MAX_TENSORS);
  if (!a->param || !a->exp_avg || !a->exp_avg_sq || !a->step || !a->sync || !a->lr) return fail(PIT_ERR_ARG, "null pointer");
  pit::AllReduceAdamParams P = {};
  P.world = a->world, P.rank = a->rank, P.n_tensors = a->n_tensors;
  int64_t total = 0;
  for (int k = 0; k < a->n_tensors; ++k) {
    if (a->numel[k] < 0) return fail(PIT_ERR_ARG, "allreduce_adam: negative tensor size");
    P.grad[k] = a->grad[k];
    P.numel[k] = a->numel[k];
    total += ((int64_t)a->numel[k] + 3) / 4 * 4;
  }
  if (total < 1) return fail(PIT_ERR_ARG, "allreduce_adam: no elements");
  P.total = total;
  P.stride = (total + 3) / 4 * 4;
  for (int r = 0; r < a->world; ++r) {
    if (a->world > 1 && !a->region[r]) return fail(PIT_ERR_ARG, "allreduce_adam: null peer region");
    P.region[r] = a->region[r];
  }
  P.param = a->param, P.exp_avg = a->exp_avg, P.exp_avg_sq = a->exp_avg_sq, P.step = a->step, P.arrive = a->sync, P.lr = a->lr;
  P.beta1 = a->beta1, P.beta2 = a->beta2, P.eps = a->eps;
  PIT_CUDA(launch::allreduce_adam(P, launch::allreduce_adam_grid(total, sm_count(), a->world > 1), static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// fused narrow-input MLP (the encoder lift)
// ---------------------------------------------------------------------------------------------------------------------
int pit_mlp_fused_supported(int64_t rows, int32_t in_dim, int32_t hid_dim, int32_t out_dim) {
  return rows >= 1 && rows <= ((int64_t)1 << 31) - 64 && in_dim >= 1 && in_dim <= pit::MF_MAX_IN && hid_dim == out_dim && (hid_dim == 32 || hid_dim == 64) ? 1 : 0;
}

int pit_mlp_fused_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t rows, int32_t in_dim,
                          int32_t hid_dim, int32_t act_out, int32_t linear_3xtf32, float* z1, float* z2, float* out, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !z1 || !z2 || !out) return fail(PIT_ERR_ARG, "null pointer");
  if (!pit_mlp_fused_supported(rows, in_dim, hid_dim, hid_dim)) return fail(PIT_ERR_ARG, "mlp_fused: unsupported shape (see pit_mlp_fused_supported)");
  if ((reinterpret_cast<uintptr_t>(b2) | reinterpret_cast<uintptr_t>(z2) | reinterpret_cast<uintptr_t>(out)) & 7u)
    return fail(PIT_ERR_ARG, "mlp_fused: b2, z2, out must be 8-byte aligned");
  pit::MlpFusedParams P = {};
  P.x = x, P.w1 = w1, P.b1 = b1, P.w2 = w2, P.b2 = b2, P.z1 = z1, P.z2 = z2, P.out = out;
  P.R = rows, P.K = in_dim, P.act_out = act_out;
  PIT_CUDA(launch::mlp_fused(false, hid_dim, linear_3xtf32 != 0, P, static_cast<cudaStream_t>(stream)));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

int pit_mlp_fused_backward(const float* x, const float* w1, const float* w2, const float* z1, const float* z2, const float* d_out, int64_t rows,
                           int32_t in_dim, int32_t hid_dim, int32_t act_out, int32_t linear_3xtf32, float* d_x, float* grads, void* stream) {
  if (!x || !w1 || !w2 || !z1 || !z2 || !d_out || !grads) return fail(PIT_ERR_ARG, "null pointer");
  if (!pit_mlp_fused_supported(rows, in_dim, hid_dim, hid_dim)) return fail(PIT_ERR_ARG, "mlp_fused: unsupported shape (see pit_mlp_fused_supported)");
  if ((reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(z1)) & 7u) return fail(PIT_ERR_ARG, "mlp_fused: grads and z1 must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t d = (size_t)hid_dim, k = (size_t)in_dim;
  pit::MlpFusedParams P = {};
  P.x = x, P.w1 = w1, P.w2 = w2, P.z1 = const_cast<float*>(z1), P.z2 = const_cast<float*>(z2), P.d_out = d_out, P.d_x = d_x;
  P.R = rows, P.K = in_dim, P.act_out = act_out;
  // grads: d_w2 [D,D] | d_b2 [D] | d_b1 [D] | d_w1 [D,K]  (the 8-byte aligned pieces first)
  P.d_w2 = grads, P.d_b2 = grads + d * d, P.d_b1 = grads + d * d + d, P.d_w1 = grads + d * d + 2 * d;
  PIT_CUDA(cudaMemsetAsync(grads, 0, (d * d + 2 * d + d * k) * sizeof(float), st));
  PIT_CUDA(launch::mlp_fused(true, hid_dim, linear_3xtf32 != 0, P, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PIT_OK;
}

}  // extern "C"
