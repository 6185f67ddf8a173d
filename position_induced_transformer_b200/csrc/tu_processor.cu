// Instantiations of the fused processor kernels (processor_block.cuh): cluster launch, one cluster per sample.
#include <stdio.h>
#include <stdlib.h>

#include "dispatch.h"
#include "launchers.h"

namespace pit {
namespace launch {

namespace {
template <typename K>
cudaError_t launch_cluster(K kernel, int tiles, int batch, size_t smem, const ProcParams& P, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (tiles > 8) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
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

// Rows per CTA: 16 (a cluster of N/16 <= 16 CTAs, a non-portable size) when every cluster of the launch can be resident at once,
// which halves the serial work of a CTA; else 32 (N/32 <= 8 CTAs per cluster, always launchable).
template <typename K>
int clusters_resident(K kernel, int tiles, size_t smem) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  if (tiles > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, 1, 1);
  cfg.blockDim = dim3(PB_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tiles;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

template <int D, int NH, int TR>
cudaError_t processor_tr(bool backward, bool lin3, const ProcParams& P, cudaStream_t st) {
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

template <int D, int NH>
cudaError_t processor_dn(bool backward, bool lin3, const ProcParams& P, cudaStream_t st) {
  // the choice is a function of (N, B) and of the device: cached per shape
  static int cached_n = 0, cached_b = 0, cached_tr = 0;
  int tr = PROC_TILE_ROWS;
  if (P.N == cached_n && P.B == cached_b) {
    tr = cached_tr;
  } else {
    if (P.N % 16 == 0 && P.N / 16 <= 16 && P.N / 16 > 1) {
      const size_t smem = proc_bwd_smem_floats<D, NH, 16>(P.N) * sizeof(float);
      const int fit = lin3 ? clusters_resident(processor_bwd_kernel<D, NH, 16, true>, P.N / 16, smem)
                           : clusters_resident(processor_bwd_kernel<D, NH, 16, false>, P.N / 16, smem);
      const size_t smem_f = proc_fwd_smem_floats<D, NH, 16>(P.N) * sizeof(float);
      const int fit_f = lin3 ? clusters_resident(processor_fwd_kernel<D, NH, 16, true>, P.N / 16, smem_f)
                             : clusters_resident(processor_fwd_kernel<D, NH, 16, false>, P.N / 16, smem_f);
      if (fit >= P.B && fit_f >= P.B) tr = 16;
      if (getenv("PIT_DEBUG")) {
        const size_t s32 = proc_bwd_smem_floats<D, NH, PROC_TILE_ROWS>(P.N) * sizeof(float);
        fprintf(stderr, "[pit] processor N=%d B=%d: resident clusters of %d CTAs (16 rows): bwd %d fwd %d; of %d CTAs (32 rows): bwd %d -> %d rows per CTA\n",
                P.N, P.B, P.N / 16, fit, fit_f, P.N / PROC_TILE_ROWS,
                clusters_resident(processor_bwd_kernel<D, NH, PROC_TILE_ROWS, false>, P.N / PROC_TILE_ROWS, s32), tr);
      }
    }
    cached_n = P.N, cached_b = P.B, cached_tr = tr;
  }
  return tr == 16 ? processor_tr<D, NH, 16>(backward, lin3, P, st) : processor_tr<D, NH, PROC_TILE_ROWS>(backward, lin3, P, st);
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
