/*
 * pit_posatt.h -- C ABI of the B200 (sm_100a) position-attention library, libpit_posatt.so.
 *
 * Drop-in boundary for the hot path of the Position-induced Transformer reference
 * (junfeng-chen/position_induced_transformer, file pit.py).  Every entry point replaces a
 * piece of reference Python that runs on materialised N x M tensors:
 *
 *   pit_quantile_ranks      torch.quantile rank arithmetic        pit.py:49, 136, 197, 255
 *   pit_rowstat             the row sort inside torch.quantile    pit.py:49, 136, 197, 255
 *   pit_posatt_forward      dist2att + convolution (+ concat)     pit.py:46-57, 133-144, 190-200, 247-258, 37-44
 *   pit_posatt_backward     what autograd replays for the above   (no explicit code in the reference)
 *   pit_posatt_backward_coords  ... with respect to the meshes    pit.py:47, 134, 191-195, 248-253 (dist2att is differentiable in them)
 *   pit_head_scale*         the scale map tan(c*(1+sin(lmda)))    pit.py:48, 135, 196, 254
 *   pit_bias_act*           bias + GELU epilogues of the MLPs     pit.py:21-26, 111, 121
 *   pit_rel_lp*             the training loss RelLpNorm           utils.py:60-98
 *   pit_decoder_tail*       pit.decoder = up + de MLP, fused      pit.py:124-127, 21-26
 *   pit_tail_plan*          lambda-independent part of the mask   pit.py:136 (re-sorted every step there)
 *   pit_processor*          pit.processor, all blocks, fused      pit.py:114-122, 37-44, 21-26
 *   pit_mlp_fused*          the encoder lift en_layer (+ GELU)     pit.py:110-111, 21-26
 *   pit_allreduce_adam      optimizer.step() of the scripts + the gradient SUM of a data-parallel run   train_darcy.py:115, 131
 *
 * Conventions
 *   - all tensors are fp32, contiguous, row-major, resident on the CURRENT CUDA device;
 *   - pointers are device pointers unless a parameter is documented as host;
 *   - nothing is allocated, freed or synchronised inside the library: outputs and the
 *     workspace are caller-owned; kernels are enqueued on `stream` (a cudaStream_t);
 *   - every function returns PIT_OK (0) or a negative PIT_ERR_* code; pit_last_error()
 *     returns a thread-local message for the last failure;
 *   - there is no CPU fallback: without a CUDA device every launch returns PIT_ERR_CUDA.
 */
#ifndef PIT_POSATT_H_
#define PIT_POSATT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIT_ABI_VERSION 11

#define PIT_OK 0
#define PIT_ERR_ARG (-1)       /* bad shape / null pointer / unsupported configuration */
#define PIT_ERR_CUDA (-2)      /* CUDA runtime reported an error (launch, no device, ...) */
#define PIT_ERR_WORKSPACE (-3) /* caller-provided workspace too small */

/* Distance variants (reference class families). */
#define PIT_EUCLID 0     /* posatt / posatt_fixed            pit.py:47, 134  */
#define PIT_PERIODIC1D 1 /* posatt_periodic1d                pit.py:191-195  */
#define PIT_PERIODIC2D 2 /* posatt_periodic2d                pit.py:248-253  */

/* Problem description shared by all entry points. */
typedef struct pit_problem {
  int32_t variant;      /* PIT_EUCLID | PIT_PERIODIC1D | PIT_PERIODIC2D                          */
  int32_t space_dim;    /* coordinates per point: 1 or 2                                          */
  int32_t mesh_batched; /* 0: meshes are (N,sd)/(M,sd), shared by the batch (posatt_fixed, ...)   */
                        /* 1: meshes are (B,N,sd)/(B,M,sd), one per sample   (posatt)             */
  int32_t batch;        /* B: samples in `values`                                                 */
  int32_t n_head;       /* H                                                                      */
  int32_t n_out;        /* N: rows   = points of mesh_out                                         */
  int32_t n_in;         /* M: columns = points of mesh_in                                         */
  int32_t dim;          /* D: value features per point                                            */
} pit_problem_t;

/* Row statistics in squared-distance space (head independent); each array has
 * (mesh_batched ? B : 1) * N entries. v_lo / v_hi may be NULL for a global stage. */
typedef struct pit_rowstat {
  const float* v_min; /* smallest d2 of the row                                     */
  const float* v_lo;  /* k_lo-th smallest d2 (0-based)                              */
  const float* v_hi;  /* k_hi-th smallest d2                                        */
  float weight;       /* interpolation weight w of torch.quantile, in [0,1)        */
  int32_t masked;     /* 0: locality >= 1, every column kept; 1: apply the quantile mask */
  int32_t rank_hi;    /* k_hi of pit_quantile_ranks (the 0-based rank of v_hi), or 0 if not known                       */
  /* Neighbour lists written by pit_rowstat_lists (all three NULL if not built): for every row the columns with d2 <= v_hi -- a  */
  /* superset of what any head keeps -- as up to 32 {column, d2} entries in ascending column order.  With them the masked stages */
  /* over per-sample meshes never sweep the N x M pairs again.                                                                  */
  const int16_t* nbr_idx; /* [(B),N,32]                                                       */
  const float* nbr_d2;    /* [(B),N,32]                                                       */
  const int32_t* nbr_cnt; /* [(B),N] true number of such columns (> 32: the list is incomplete and the stage must not use it) */
} pit_rowstat_t;

/* Tile plan of a mesh pair (described with pit_tail_plan_rows / pit_tail_plan_fill below). */
#define PIT_PLAN_ROWS 0    /* tiles of mesh_out rows against mesh_in candidates (M <= 1024): decoder stages              */
#define PIT_PLAN_COLUMNS 1 /* tiles of mesh_in columns against mesh_out candidates (N <= 1024): local encoder stages     */

typedef struct pit_tail_plan {
  const void* rec;
  const int32_t* tile_off;
  const int32_t* tile_cnt;
  const int16_t* cand;
  const float* d2;
  int32_t n_tiles;
} pit_tail_plan_t;

/* Operand precision of the GLOBAL stages (locality >= 1: the dense contraction of pit.py:54-57 on tcgen05; accumulation is fp32
 * in every mode; masked stages are always exact fp32).  Process-wide, default PIT_DENSE_FP32.
 *   PIT_DENSE_FP32  3xTF32 products: parity with the fp32 reference (rel-Linf ~1e-6)
 *   PIT_DENSE_TF32  single TF32 products: what torch runs the reference's einsum as under pit.py:2 ('high'); bound 1e-3
 *   PIT_DENSE_BF16  operands rounded to BF16, fp32 accumulation; bound 5e-3
 * The reduced modes apply to the forward product and to the value gradient; the scale gradient (a difference of large sums)
 * always multiplies 3xTF32. */
#define PIT_DENSE_FP32 0
#define PIT_DENSE_TF32 1
#define PIT_DENSE_BF16 2
int pit_set_dense_precision(int32_t precision);
int pit_get_dense_precision(void);

/* Library identification. */
int pit_abi_version(void);
const char* pit_last_error(void);
/* Number of kernel launches issued by this process through the library so far. */
uint64_t pit_launch_count(void);

/* Host-side: rank arithmetic of torch.quantile(x, q, dim=-1) with linear interpolation
 * on a row of m entries: rank = fp32(q) * fp32(m-1); k_lo = floor, k_hi = ceil, w = rank - k_lo. */
int pit_quantile_ranks(double q, int32_t m, int32_t* k_lo, int32_t* k_hi, float* w);

/* Bytes of scratch pit_rowstat / pit_posatt_forward / pit_posatt_backward may need for `p`. */
size_t pit_workspace_bytes(const pit_problem_t* p);

/* Per-row order statistics of the squared distances.
 *   mesh_out  [(B),N,sd]   mesh_in [(B),M,sd]
 *   period    device pointer to the wrap length l (periodic variants), else NULL
 *   v_min, v_lo, v_hi  outputs, [(B),N] each                                              */
int pit_rowstat(const pit_problem_t* p, const float* mesh_out, const float* mesh_in,
                const float* period, int32_t k_lo, int32_t k_hi,
                float* v_min, float* v_lo, float* v_hi, void* stream);

/* pit_rowstat plus the neighbour lists described at pit_rowstat_t (M <= 1024); the lists come out of the sweep the order
 * statistics need anyway. */
int pit_rowstat_lists(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period, int32_t k_lo, int32_t k_hi,
                      float* v_min, float* v_lo, float* v_hi, int16_t* nbr_idx, float* nbr_d2, int32_t* nbr_cnt, void* stream);

/* Fused position-attention forward.
 *   values [B,M,D]; scale [H] = tan(c*(1+sin(lmda))) (pit.py:48), computed by the caller;
 *   out: row (b,n) starts at out + ((size_t)b*N + n)*ld_out + col_off and receives H*D floats,
 *        head-major (h*D + d) -- ld_out = H*D, col_off = 0 for a cross stage;
 *        ld_out = (1+H)*D, col_off = D for the self stage with concat (the caller, or
 *        copy_values != 0, fills the first D columns with `values`, which requires N == M);
 *   rowsum [(B),H,N]: sum of unnormalised weights exp(s*v_min - s*d2) of each row (saved for backward).
 *   column_plan: optional PIT_PLAN_COLUMNS tile plan of the mesh pair (NULL: none), used by masked stages with few rows and a
 *        huge column set (the local encoder); the same plan must then be passed to pit_posatt_backward.  */
int pit_posatt_forward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in,
                       const float* period, const float* values, const float* scale,
                       const pit_rowstat_t* stat, float* out, int64_t ld_out, int64_t col_off,
                       int32_t copy_values, float* rowsum, void* workspace, size_t workspace_bytes,
                       const pit_tail_plan_t* column_plan, void* stream);

/* Fused backward.  d_out uses the same (ld_out, col_off) addressing as `out`.
 *   d_values [B,M,D]  (may be NULL: not computed).  If accumulate_concat != 0 the first D columns
 *                     of d_out (the concat pass-through, pit.py:44) are added into d_values.
 *   d_scale [H]       dL/ds_h, summed over rows (and over the batch); overwritten. May be NULL. */
int pit_posatt_backward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in,
                        const float* period, const float* values, const float* scale,
                        const pit_rowstat_t* stat, const float* rowsum, const float* d_out,
                        int64_t ld_out, int64_t col_off, int32_t accumulate_concat,
                        float* d_values, float* d_scale, void* workspace, size_t workspace_bytes,
                        const pit_tail_plan_t* column_plan, void* stream);

/* Gradient with respect to the mesh coordinates (`dX`): what autograd gives the reference when a mesh requires grad (no script
 * does; a learnable latent mesh would).  Same inputs as pit_posatt_backward; outputs, all overwritten:
 *   d_mesh_out [(B),N,sd], d_mesh_in [(B),M,sd]   (summed over the batch for shared meshes; the two may NOT alias)
 *   d_period   [1]  gradient with respect to the wrap length l of the periodic variants (NULL for PIT_EUCLID); l is a function
 *              of mesh_in (pit.py:191-192, 248-250) and the caller chains d_period through that expression.
 * No gradient flows through the quantile mask, as in the reference.  A correctness path: one warp per (sample, row). */
int pit_posatt_backward_coords(const pit_problem_t* p, const float* mesh_out, const float* mesh_in, const float* period,
                               const float* values, const float* scale, const pit_rowstat_t* stat, const float* rowsum,
                               const float* d_out, int64_t ld_out, int64_t col_off, float* d_mesh_out, float* d_mesh_in,
                               float* d_period, void* stream);

/* Per-head scale map of pit.py:48:  scale[i] = tan(c * (1 + sin(lmda[i]))),  c = fp32(0.25*pi*(1-1e-7)),
 * each operation rounded to fp32 separately as the reference's chain of torch ops does (sin, add, mul, tan:
 * four launches there, one here), and its derivative
 *   d_lmda[i] = d_scale[i] * c * cos(lmda[i]) * (1 + scale[i]^2)
 * (seven launches of autograd there).  n = number of heads; all pointers device. */
int pit_head_scale_forward(const float* lmda, float* scale, int32_t n, void* stream);
int pit_head_scale_backward(const float* lmda, const float* scale, const float* d_scale, float* d_lmda, int32_t n,
                            void* stream);

/* Epilogue of the Linear layers of kaiming_mlp (pit.py:21-26) and of the GELUs around them (pit.py:111, 121):
 *   out[r,c] = act(z[r,c] + bias[c]),   act = exact (erf) GELU if apply_gelu else identity,
 * one coalesced 128-bit pass, and its backward in one pass as well:
 *   d_z[r,c] = d_out[r,c] * act'(z[r,c] + bias[c]),   d_bias[c] = sum_r d_z[r,c]   (overwritten)
 * -- what autograd does with a GELU-backward kernel plus a separate column reduction.  z is the bias-free GEMM output
 * [rows, cols]; cols must be a multiple of 4 that divides 1024 (every hidden width of the reference). */
int pit_bias_act_supported(int64_t rows, int32_t cols);
int pit_bias_act_forward(const float* z, const float* bias, float* out, int64_t rows, int32_t cols, int32_t apply_gelu,
                         void* stream);
int pit_bias_act_backward(const float* z, const float* bias, const float* d_out, float* d_z, float* d_bias, int64_t rows,
                          int32_t cols, int32_t apply_gelu, void* stream);

/* Relative Lp error of the reference's utils.py:60-98 (RelLpNorm):
 *   loss = sum_b (1/O) sum_o ||truth[b,:,o] - pred[b,:,o]||_p / ||truth[b,:,o]||_p,   p = 1 or 2, O <= 4,
 * truth / pred [B, L, O].  norms [B, O, 2] receives (||e||_p, ||truth||_p) and is consumed by the backward, which
 * writes d_pred = d_loss * d loss / d pred (d_loss: device scalar). */
int pit_rel_lp_supported(int32_t batch, int64_t length, int32_t out_dim, int32_t p);
int pit_rel_lp_forward(const float* truth, const float* pred, int32_t batch, int64_t length, int32_t out_dim, int32_t p,
                       float* norms, float* loss, void* stream);
int pit_rel_lp_backward(const float* truth, const float* pred, const float* norms, const float* d_loss, int32_t batch,
                        int64_t length, int32_t out_dim, int32_t p, float* d_pred, void* stream);

/* Tile plan of a decoder stage over shared meshes -- everything about its locality mask that does not depend on lmda.
 * The reference re-derives the mask from a full row sort in every step (torch.quantile, pit.py:136) although the
 * meshes never change (train_darcy.py:88-96, 128).  The plan sorts the N rows by their candidate set {j : d2(row, j) <=
 * v_hi(row) (1 + 1e-6)} (a superset of what any head can keep) and cuts the sorted order into tiles of 32 rows:
 *   rec      [n_tiles*32] x float4  {v_min, v_lo, v_hi, row index (int32 bit pattern; -1 = padding of the last tile)}
 *   tile_off [n_tiles+1] int32      offsets of the tiles' candidate lists in `cand`: multiples of 8, so that every list
 *                                   (and its lines of d2) starts on a 16-byte boundary; n_cand = tile_off[n_tiles]
 *   tile_cnt [n_tiles] int32        number of candidates of each tile
 *   cand     [n_cand] int16         candidate columns, ascending inside a tile (then zero padding up to the next offset)
 *   d2       [n_cand*32] float      squared distance (reference rounding order, pit.py:134 / 193-195 / 251-253) of tile
 *                                   row r to candidate k of tile t at d2[(tile_off[t] + k)*32 + r]
 * n_tiles = ceil(N / 32).  Build it once per (mesh_out, mesh_in, locality):
 *   1. pit_tail_plan_rows   sorts the rows and writes tile_off / tile_cnt; the caller reads tile_off[n_tiles] (one int32, the only
 *                           device->host read of the library's protocol) and allocates `cand` and `d2`;
 *   2. pit_tail_plan_fill   writes cand, d2, rec.  Same workspace as step 1, untouched in between.
 * Requires shared meshes with M <= 1024.  Pass the plan to pit_decoder_tail_forward / _backward (NULL: no plan, the
 * kernels scan the latent mesh themselves).
 * The same structure with the roles of the meshes exchanged (side = PIT_PLAN_COLUMNS: the mesh_in columns are sorted by the
 * set of mesh_out rows that can see them; `cand` then holds row indices, rec = {0, 0, 0, column}, d2 the distance of tile
 * column c to candidate row k) serves the local ENCODER stage (few rows, a huge column set, pit.py:109 at Darcy-421:
 * 256 x 177 241): pass it to pit_posatt_forward / _backward as `column_plan`. */

size_t pit_tail_plan_workspace_bytes(const pit_problem_t* p, int32_t side);
int pit_tail_plan_rows(const pit_problem_t* p, int32_t side, const float* mesh_out, const float* mesh_in, const float* period,
                       const pit_rowstat_t* stat, int32_t* tile_off, int32_t* tile_cnt, void* workspace, size_t workspace_bytes,
                       void* stream);
int pit_tail_plan_fill(const pit_problem_t* p, int32_t side, const float* mesh_out, const float* mesh_in, const float* period,
                       const pit_rowstat_t* stat, const int32_t* tile_off, void* rec, int16_t* cand, float* d2,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Fused decoder tail: pit.decoder (pit.py:124-127) = cross position-attention `up` + kaiming_mlp `de`
 * (pit.py:21-26), for shared meshes with M <= 1024, H <= 2, hidden width C a power of two in [32, 512], out_dim <= 4.
 * problem->dim is the hidden width C.  The first Linear is pushed through the (linear) attention by the caller:
 *   y [B,M,H,C]  with  y[b,j,h,:] = W1[:, h*D:(h+1)*D] @ U[b,j,:]        (W1 = de.mlp1.weight, U = latent features)
 *   out[b,n,o]  = b2[o] + sum_c W2[o,c] * gelu(b1[c] + sum_h sum_j A_h[n,j] * y[b,j,h,c])      (exact erf GELU)
 * so nothing of size N x H*D or N x C is ever written to memory.  rowsum (2*H*ceil32(N) floats, opaque: row sums and,
 * with a tile plan, the rows' mean squared distances in tile order) is saved for the backward -- pass the same plan --, which
 * returns d_y [B,M,H,C] (dW1 and dU follow from it through the caller's GEMM), d_scale [H], d_b1 [C], d_w2 [O,C],
 * d_b2 [O]; every gradient buffer is overwritten. */
int pit_decoder_tail_supported(const pit_problem_t* p, int32_t out_dim);
int pit_decoder_tail_forward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in,
                             const float* period, const float* y, const float* scale,
                             const pit_rowstat_t* stat, const float* b1, const float* w2, const float* b2,
                             int32_t out_dim, float* out, float* rowsum, const pit_tail_plan_t* plan, void* stream);
int pit_decoder_tail_backward(const pit_problem_t* p, const float* mesh_out, const float* mesh_in,
                              const float* period, const float* y, const float* scale,
                              const pit_rowstat_t* stat, const float* b1, const float* w2, const float* b2,
                              int32_t out_dim, const float* rowsum, const float* d_out, float* d_y,
                              float* d_scale, float* d_b1, float* d_w2, float* d_b2, const pit_tail_plan_t* plan,
                              void* stream);

/* Fused processor: pit.processor (pit.py:114-122) for a shared latent mesh -- n_blocks x [ self position-attention with
 * locality 1 and concat (pit.py:37-44) + kaiming_mlp (pit.py:21-26) + GELU (pit.py:121) ] -- as ONE launch per direction.
 * problem: mesh_batched = 0, n_out = n_in = N latent points (a multiple of 32, at most 256), dim = hidden width D (32 or 64),
 * n_head <= 2; n_blocks <= 8.  One thread-block cluster per sample; nothing N x N is written to memory.
 *   mesh   [N,sd]           period: device scalar (periodic variants) or NULL
 *   x0     [B,N,D]          input of the first block
 *   scale  [n_blocks,H]     tan(c*(1+sin(lmda))) of every block's attention layer (pit_head_scale_forward)
 *   blocks host array of n_blocks entries: mlp1.weight [D,(1+H)D], mlp1.bias [D], mlp2.weight [D,D], mlp2.bias [D] (device)
 *   linear_3xtf32  != 0: the Linear products run as 3xTF32 (fp32 parity, torch 'highest'); 0: single TF32 products, what torch
 *                  runs them as under set_float32_matmul_precision('high') (pit.py:2).  Attention products are always 3xTF32.
 *   saved  pit_processor_saved_floats() floats, written by the forward and consumed by the backward (opaque)
 *   out    [B,N,D]          output of the last block
 * Backward: d_out [B,N,D] -> d_x0 [B,N,D]; `grads` (pit_processor_grad_floats() floats, overwritten) receives per block
 * d_mlp1.weight [D,(1+H)D] | d_mlp1.bias [D] | d_mlp2.weight [D,D] | d_mlp2.bias [D], and after the last block d_scale
 * [n_blocks,H]; scratch: pit_processor_scratch_floats() floats. */
typedef struct pit_processor_block {
  const float* w1;
  const float* b1;
  const float* w2;
  const float* b2;
} pit_processor_block_t;
int pit_processor_supported(const pit_problem_t* p, int32_t n_blocks);
size_t pit_processor_saved_floats(const pit_problem_t* p, int32_t n_blocks);
size_t pit_processor_grad_floats(const pit_problem_t* p, int32_t n_blocks);
size_t pit_processor_scratch_floats(const pit_problem_t* p);
int pit_processor_forward(const pit_problem_t* p, int32_t n_blocks, const float* mesh, const float* period, const float* x0,
                          const float* scale, const pit_processor_block_t* blocks, int32_t linear_3xtf32, float* saved, float* out,
                          void* stream);
int pit_processor_backward(const pit_problem_t* p, int32_t n_blocks, const float* mesh, const float* period, const float* x0,
                           const float* scale, const pit_processor_block_t* blocks, int32_t linear_3xtf32, const float* saved,
                           const float* d_out, float* d_x0, float* grads, float* scratch, void* stream);

/* Data-parallel optimizer step in one launch: SUM of the flat gradient over `world` ranks through peer-mapped ("symmetric")
 * memory on NVLink, fused with torch.optim.Adam's update (no weight decay, no amsgrad: train_darcy.py:115).  The reference
 * has no multi-GPU path; with world = 1 this is optimizer.step() alone (train_darcy.py:131).
 *   grad[k], numel[k]  this rank's gradient tensors in parameter order (device pointers; NULL = all zeros), k < n_tensors <= 64
 *   region[r]          rank r's symmetric region as mapped into THIS process (r < world <= 16; unused for world = 1):
 *                      pit_allreduce_adam_region_floats(total) floats, zero-filled once before the first step on every rank
 *                      (64 flag words, then two gradient buckets that alternate with the parity of *step)
 *   param, exp_avg, exp_avg_sq   flat fp32 buffers of `total` = sum of ceil4(numel[k]) elements: tensor k starts at the sum of the
 *                      rounded sizes before it, i.e. on a 16-byte boundary (parameters are views into `param`)
 *   step               device int32: number of updates done so far (incremented by the kernel; bias corrections use step + 1)
 *   sync               three device uint32 words, zero-initialised once: two monotone CTA counters and an error flag that is
 *                      set to 1 if a peer's flag did not arrive within 2 s (the kernel then carries on instead of hanging)
 *   lr                 device float (so that a scheduler can change it under a replayed CUDA graph)
 * Every rank must call this once per step, in the same order; the kernel's CTAs wait for each other and for the peers. */
typedef struct pit_allreduce_adam {
  int32_t world, rank, n_tensors;
  const float* grad[64];
  int32_t numel[64];
  float* region[16];
  float* param;
  float* exp_avg;
  float* exp_avg_sq;
  int32_t* step;
  uint32_t* sync;
  const float* lr;
  float beta1, beta2, eps;
} pit_allreduce_adam_t;
size_t pit_allreduce_adam_region_floats(int64_t total);
int pit_allreduce_adam(const pit_allreduce_adam_t* a, void* stream);

/* A whole kaiming_mlp with a narrow input (pit.py:21-26; the encoder lift `en_layer` of the shared-mesh models, pit.py:110-111):
 *   out = act(W2 gelu(W1 x + b1) + b2),  act = exact GELU if act_out else identity,
 * x [rows, in_dim] with in_dim <= 32, W1 [hid, in_dim], W2 [hid, hid], hid in {32, 64}; one launch per direction.  z1, z2
 * [rows, hid] receive the pre-activations for the backward.  linear_3xtf32: as pit_processor_forward (the first product always
 * runs on the fp32 pipe).  Backward: d_out [rows, hid] -> d_x [rows, in_dim] (may be NULL) and `grads`, overwritten:
 *   d_W2 [hid, hid] | d_b2 [hid] | d_b1 [hid] | d_W1 [hid, in_dim]. */
int pit_mlp_fused_supported(int64_t rows, int32_t in_dim, int32_t hid_dim, int32_t out_dim);
int pit_mlp_fused_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t rows, int32_t in_dim,
                          int32_t hid_dim, int32_t act_out, int32_t linear_3xtf32, float* z1, float* z2, float* out, void* stream);
int pit_mlp_fused_backward(const float* x, const float* w1, const float* w2, const float* z1, const float* z2, const float* d_out, int64_t rows,
                           int32_t in_dim, int32_t hid_dim, int32_t act_out, int32_t linear_3xtf32, float* d_x, float* grads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIT_POSATT_H_ */
