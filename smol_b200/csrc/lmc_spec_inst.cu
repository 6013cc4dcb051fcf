// Instantiations of the speculative-batch Metropolis kernel.
#include "lmc_spec.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int USHER, int SG, bool LISTS, bool EWF, int MAXT, int MINB>
static int launch_spec_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_spec_kernel<KONE, USHER, SG, LISTS, EWF, MAXT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

template <bool KONE, int USHER>
static int launch_spec_one(const DevModel& m, const RunArgs& a, int sg, bool lists, const LaunchCfg& lc) {
  constexpr bool SW = USHER == LMC_USHER_SWAP;   // position lists serve the swap usher only
  if (sg == 1) return launch_spec_k<KONE, USHER, 1, SW, false, 128, 7>(m, a, lc);
  if (sg == 2) return launch_spec_k<KONE, USHER, 2, false, false, 128, 7>(m, a, lc);
  if (SW && lists) return launch_spec_k<KONE, USHER, 4, SW, false, 128, 7>(m, a, lc);
  return launch_spec_k<KONE, USHER, 4, false, false, 128, 7>(m, a, lc);
}

// plain variants: no Ewald term, blocks of 128 threads; sg = lanes per speculated step, lists = swap partner
// from sorted position lists (always with sg == 1)
int launch_spec(const DevModel& m, const RunArgs& a, bool kone, int usher, int sg, bool lists, const LaunchCfg& lc) {
  if (usher == LMC_USHER_FLIP)
    return kone ? launch_spec_one<true, LMC_USHER_FLIP>(m, a, sg, lists, lc) : launch_spec_one<false, LMC_USHER_FLIP>(m, a, sg, lists, lc);
  if (usher == LMC_USHER_SWAP)
    return kone ? launch_spec_one<true, LMC_USHER_SWAP>(m, a, sg, lists, lc) : launch_spec_one<false, LMC_USHER_SWAP>(m, a, sg, lists, lc);
  return -2;
}

}  // namespace lmc
