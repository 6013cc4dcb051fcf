// Instantiations of the speculative-batch Metropolis kernel.
#include "lmc_spec.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int USHER, int SG, bool EWF, int MAXT, int MINB>
static int launch_spec_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_spec_kernel<KONE, USHER, SG, (SG == 1 && USHER == LMC_USHER_SWAP), EWF, MAXT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

template <bool KONE, int USHER>
static int launch_spec_one(const DevModel& m, const RunArgs& a, int sg, const LaunchCfg& lc) {
  if (sg == 1) return launch_spec_k<KONE, USHER, 1, false, 128, 7>(m, a, lc);
  if (sg == 2) return launch_spec_k<KONE, USHER, 2, false, 128, 7>(m, a, lc);
  return launch_spec_k<KONE, USHER, 4, false, 128, 7>(m, a, lc);
}

// plain variants: no Ewald term, blocks of 128 threads; sg = lanes per speculated step
int launch_spec(const DevModel& m, const RunArgs& a, bool kone, int usher, int sg, const LaunchCfg& lc) {
  if (usher == LMC_USHER_FLIP)
    return kone ? launch_spec_one<true, LMC_USHER_FLIP>(m, a, sg, lc) : launch_spec_one<false, LMC_USHER_FLIP>(m, a, sg, lc);
  if (usher == LMC_USHER_SWAP)
    return kone ? launch_spec_one<true, LMC_USHER_SWAP>(m, a, sg, lc) : launch_spec_one<false, LMC_USHER_SWAP>(m, a, sg, lc);
  return -2;
}

}  // namespace lmc
