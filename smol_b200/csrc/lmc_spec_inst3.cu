// Instantiations of the speculative-batch Metropolis kernel: environment-word variants (ENV), four lanes per step.
#include "lmc_spec.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int USHER, bool EWF, int MAXT, int MINB, int EB>
static int launch_spec_kb(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_spec_kernel<KONE, USHER, 4, false, EWF, MAXT, MINB, EB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

template <bool KONE, int USHER, bool EWF, int MAXT, int MINB>
static int launch_spec_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  if (m.envB == 1) return launch_spec_kb<KONE, USHER, EWF, MAXT, MINB, 1>(m, a, lc);
  if (m.envB == 2) return launch_spec_kb<KONE, USHER, EWF, MAXT, MINB, 2>(m, a, lc);
  return -2;
}

template <bool KONE, int USHER>
static int launch_spec_e1(const DevModel& m, const RunArgs& a, bool ewf, bool wide, const LaunchCfg& lc) {
  if (wide) return ewf ? launch_spec_k<KONE, USHER, true, 448, 2>(m, a, lc) : launch_spec_k<KONE, USHER, false, 448, 2>(m, a, lc);
  return ewf ? launch_spec_k<KONE, USHER, true, 128, 7>(m, a, lc) : launch_spec_k<KONE, USHER, false, 128, 7>(m, a, lc);
}

// environment words (a.env); Ewald through the potential cache (ewf), blocks of 448 threads (wide)
int launch_spec_env(const DevModel& m, const RunArgs& a, bool kone, int usher, bool ewf, bool wide, const LaunchCfg& lc) {
  if (usher == LMC_USHER_FLIP)
    return kone ? launch_spec_e1<true, LMC_USHER_FLIP>(m, a, ewf, wide, lc) : launch_spec_e1<false, LMC_USHER_FLIP>(m, a, ewf, wide, lc);
  if (usher == LMC_USHER_SWAP)
    return kone ? launch_spec_e1<true, LMC_USHER_SWAP>(m, a, ewf, wide, lc) : launch_spec_e1<false, LMC_USHER_SWAP>(m, a, ewf, wide, lc);
  return -2;
}

}  // namespace lmc
