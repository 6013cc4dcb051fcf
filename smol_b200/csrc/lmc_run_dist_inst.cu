// Instantiations of the fused MC kernel for distance processors (Metropolis flip / swap, one warp per walker).
#include "lmc_kernels.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int USHER>
static int launch_dist_one(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_run_kernel<32, KONE, 0, USHER, false, true>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

int launch_run_dist(const DevModel& m, const RunArgs& a, bool kone, int usher, const LaunchCfg& lc) {
  if (usher == LMC_USHER_FLIP)
    return kone ? launch_dist_one<true, LMC_USHER_FLIP>(m, a, lc) : launch_dist_one<false, LMC_USHER_FLIP>(m, a, lc);
  if (usher == LMC_USHER_SWAP)
    return kone ? launch_dist_one<true, LMC_USHER_SWAP>(m, a, lc) : launch_dist_one<false, LMC_USHER_SWAP>(m, a, lc);
  return -2;
}

}  // namespace lmc
