// Host-side launch layer: one function per kernel family, each defined in its own translation unit so that the
// template instantiations compile in parallel (see build.py).  The C ABI (pit_posatt.cu) only plans and calls these.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "allreduce_adam.cuh"
#include "coord_gradient.cuh"
#include "decoder_tail.cuh"
#include "decoder_tail_mma.cuh"
#include "decoder_tail_plan.cuh"
#include "encoder_plan.cuh"
#include "dense_attention.cuh"
#include "local_attention.cuh"
#include "mlp_epilogue.cuh"
#include "mlp_fused.cuh"
#include "processor_block.cuh"
#include "rel_lp_loss.cuh"
#include "rowstat.cuh"
#include "sample_tile.cuh"
#include "tall_attention.cuh"
#include "wide_attention.cuh"

namespace pit {
namespace launch {

struct TallPlan {
  bool ok;
  int cpl, l4, lanes4, chunks, rows_per_unit, grid, n_slots, threads;
  size_t smem;
};

struct WidePlan {
  bool ok;
  int grid;
  size_t smem;
};

// tu_local.cu
cudaError_t rowstat(int geo, const RowstatParams& R, cudaStream_t st);
cudaError_t local_forward(int geo, int vec, int a, dim3 grid, const AttnParams& P, cudaStream_t st);
cudaError_t local_dscale(int geo, int vec, int a, dim3 grid, const AttnParams& P, cudaStream_t st);
cudaError_t local_dvalues(int geo, int vec, int a, dim3 grid, const AttnParams& P, cudaStream_t st);
cudaError_t local_forward_finalize(int64_t total, const AttnParams& P, cudaStream_t st);
cudaError_t local_dscale_finalize(int64_t items, const AttnParams& P, float* rows, cudaStream_t st);
cudaError_t reduce_scale_rows(const float* rows, int64_t n_rows, int H, float* d_scale, cudaStream_t st);
// tu_tall_fwd.cu / tu_tall_bwd.cu
cudaError_t tall_forward(int geo, const TallPlan& plan, const TallParams& P, cudaStream_t st);
cudaError_t tall_backward(int geo, const TallPlan& plan, const TallParams& P, bool with_values, cudaStream_t st);
// tu_tail_fwd.cu / tu_tail_bwd.cu
cudaError_t tail_forward(int geo, const TallPlan& plan, const TailParams& P, cudaStream_t st);
cudaError_t tail_backward(int geo, const TallPlan& plan, const TailParams& P, cudaStream_t st);
// tu_tail_mma_fwd.cu / tu_tail_mma_bwd.cu  (plan.rows_per_unit counts 16-row tiles per CTA)
cudaError_t tail_mma_forward(int geo, const TallPlan& plan, const TailParams& P, cudaStream_t st);
cudaError_t tail_mma_backward(int geo, const TallPlan& plan, const TailParams& P, cudaStream_t st);
// tu_tail_plan_build.cu / tu_tail_plan_fwd.cu / tu_tail_plan_bwd.cu  (plan.rows_per_unit is unused: the tile plan carries tiles_per_cta)
size_t tail_plan_workspace_bytes(int N);
cudaError_t tail_plan_rows(int geo, int cpl, PlanBuildParams B, int32_t* tile_off, int32_t* tile_cnt, void* ws, cudaStream_t st);
cudaError_t tail_plan_fill(int geo, int cpl, PlanBuildParams B, void* ws, cudaStream_t st);
cudaError_t tail_plan_forward(int geo, const TallPlan& plan, const TailParams& P, const TailPlanDev& V, cudaStream_t st);
cudaError_t tail_plan_backward(int geo, const TallPlan& plan, const TailParams& P, const TailPlanDev& V, cudaStream_t st);
// tu_wide.cu
int wide_pad(int width);
cudaError_t wide_forward(int geo, const WidePlan& w, const WideParams& P, cudaStream_t st);
cudaError_t wide_dscale(int geo, const WidePlan& w, const WideParams& P, cudaStream_t st);
// tu_wide_plan.cu: the encoder walk over a cached (transposed) tile plan
cudaError_t wide_plan_forward(int grid, const WideParams& P, const TailPlanDev& V, cudaStream_t st);
cudaError_t wide_plan_dscale(int grid, const WideParams& P, const TailPlanDev& V, cudaStream_t st);
// tu_loss.cu  (stage 0: partial sums, 1: finalize, 2: backward)
cudaError_t rel_lp(int stage, const LossParams& P, int grid_x, cudaStream_t st);
// tu_mlp_epilogue.cu
cudaError_t bias_act(bool backward, const EpiParams& P, int grid, cudaStream_t st);
// tu_dense.cu  (mode: DENSE_FWD / DENSE_DSCALE / DENSE_DVALUES; nv: 64, 128 or 256 value columns per tile)
cudaError_t dense(int mode, int geo, int nv, dim3 grid, const DenseParams& P, cudaStream_t st);
// both gradient modes of a small stage (64-column tiles) in one launch
cudaError_t dense_bwd_pair(int geo, const DenseParams& Ps, dim3 gs, const DenseParams& Pv, dim3 gv, cudaStream_t st);

// tu_allreduce_adam.cu: gradient all-reduce over peer memory fused with the Adam update
int allreduce_adam_grid(int64_t total, int sms, bool multi);
cudaError_t allreduce_adam(const AllReduceAdamParams& P, int grid, cudaStream_t st);
// tu_coord_gradient.cu: gradient with respect to the mesh coordinates
cudaError_t coord_gradient(int geo, const CoordGradParams& P, cudaStream_t st);
// tu_mlp_fused.cu: Linear -> GELU -> Linear (-> GELU) with a narrow input in one launch per direction
cudaError_t mlp_fused(bool backward, int D, bool lin3, const MlpFusedParams& P, cudaStream_t st);
// tu_sample_tile.cu: masked stages over per-sample meshes, tiled by (sample, 64 rows)
int sample_tile_slots(int M, int D, bool backward, int smem_optin);
cudaError_t sample_tile(bool backward, const SampleTileParams& P, cudaStream_t st);
// tu_processor.cu: the whole processor (n_blocks x [self attention + concat + MLP + GELU]) in one cluster launch per direction
constexpr int PROC_TILE_ROWS = 32;  // latent rows per CTA; the cluster of a sample has N / 32 <= 8 CTAs
size_t processor_smem_bytes(int D, int H, int N);
cudaError_t processor(bool backward, int D, int H, bool lin3, const ProcParams& P, cudaStream_t st);

}  // namespace launch
}  // namespace pit
