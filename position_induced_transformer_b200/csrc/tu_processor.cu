// Instantiations of the fused processor kernels (processor_block.cuh): cluster launch, one cluster per sample.
#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

namespace {
template <typename K>
cudaError_t launch_cluster(K kernel, int tiles, int batch, size_t smem, const ProcParams& P, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, batch, 1);
  cfg.blockDim = dim3(PB_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tiles;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, P);
}

template <int D, int NH>
cudaError_t processor_dn(bool backward, bool lin3, const ProcParams& P, cudaStream_t st) {
  constexpr int TR = PROC_TILE_ROWS;
  const int tiles = P.N / TR;
  if (!backward) {
    const size_t smem = proc_fwd_smem_floats<D, NH, TR>(P.N) * sizeof(float);
    return lin3 ? launch_cluster(processor_fwd_kernel<D, NH, TR, true>, tiles, P.B, smem, P, st)
                : launch_cluster(processor_fwd_kernel<D, NH, TR, false>, tiles, P.B, smem, P, st);
  }
  const size_t smem = proc_bwd_smem_floats<D, NH, TR>(P.N) * sizeof(float);
  return lin3 ? launch_cluster(processor_bwd_kernel<D, NH, TR, true>, tiles, P.B, smem, P, st)
              : launch_cluster(processor_bwd_kernel<D, NH, TR, false>, tiles, P.B, smem, P, st);
}
}  // namespace

size_t processor_smem_bytes(int D, int H, int N) {
  constexpr int TR = PROC_TILE_ROWS;
  if (D == 32) return (H == 1 ? proc_bwd_smem_floats<32, 1, TR>(N) : proc_bwd_smem_floats<32, 2, TR>(N)) * sizeof(float);
  return (H == 1 ? proc_bwd_smem_floats<64, 1, TR>(N) : proc_bwd_smem_floats<64, 2, TR>(N)) * sizeof(float);
}

cudaError_t processor(bool backward, int D, int H, bool lin3, const ProcParams& P, cudaStream_t st) {
  if (D == 32) return H == 1 ? processor_dn<32, 1>(backward, lin3, P, st) : processor_dn<32, 2>(backward, lin3, P, st);
  return H == 1 ? processor_dn<64, 1>(backward, lin3, P, st) : processor_dn<64, 2>(backward, lin3, P, st);
}

}  // namespace launch
}  // namespace pit
