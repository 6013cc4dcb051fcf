// Instantiations of the speculative-batch Metropolis kernel: Ewald potential cache and wide-block variants.
#include "lmc_spec.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int USHER, int SG, bool EWF, int MAXT, int MINB>
static int launch_spec_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_spec_kernel<KONE, USHER, SG, false, EWF, MAXT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

template <bool KONE, int USHER>
static int launch_spec_x1(const DevModel& m, const RunArgs& a, bool ewf, bool wide, const LaunchCfg& lc) {
  if (wide) return ewf ? launch_spec_k<KONE, USHER, 4, true, 448, 2>(m, a, lc) : launch_spec_k<KONE, USHER, 4, false, 448, 2>(m, a, lc);
  return ewf ? launch_spec_k<KONE, USHER, 4, true, 128, 7>(m, a, lc) : -2;
}

// Ewald through the potential cache (ewf) and / or blocks of 448 threads (wide); four lanes per step
int launch_spec_x(const DevModel& m, const RunArgs& a, bool kone, int usher, bool ewf, bool wide, const LaunchCfg& lc) {
  if (usher == LMC_USHER_FLIP)
    return kone ? launch_spec_x1<true, LMC_USHER_FLIP>(m, a, ewf, wide, lc) : launch_spec_x1<false, LMC_USHER_FLIP>(m, a, ewf, wide, lc);
  if (usher == LMC_USHER_SWAP)
    return kone ? launch_spec_x1<true, LMC_USHER_SWAP>(m, a, ewf, wide, lc) : launch_spec_x1<false, LMC_USHER_SWAP>(m, a, ewf, wide, lc);
  return -2;
}

}  // namespace lmc
