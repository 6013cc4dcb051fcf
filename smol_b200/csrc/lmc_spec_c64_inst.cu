// Instantiations of the speculative kernel over compact environment words.
#include "lmc_spec_c64.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int USHER, int EB>
static int launch_c64_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_spec_c64_kernel<KONE, USHER, EB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

template <bool KONE, int USHER>
static int launch_c64_b(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  if (m.c64B == 1) return launch_c64_k<KONE, USHER, 1>(m, a, lc);
  if (m.c64B == 2) return launch_c64_k<KONE, USHER, 2>(m, a, lc);
  return -2;
}

int launch_spec_c64(const DevModel& m, const RunArgs& a, bool kone, int usher, const LaunchCfg& lc) {
  if (usher == LMC_USHER_FLIP) return kone ? launch_c64_b<true, LMC_USHER_FLIP>(m, a, lc) : launch_c64_b<false, LMC_USHER_FLIP>(m, a, lc);
  if (usher == LMC_USHER_SWAP) return kone ? launch_c64_b<true, LMC_USHER_SWAP>(m, a, lc) : launch_c64_b<false, LMC_USHER_SWAP>(m, a, lc);
  return -2;
}

}  // namespace lmc
