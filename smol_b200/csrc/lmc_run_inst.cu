// Instantiations of the fused MC kernel for ONE group size (compile with -DLMC_G=4|8|16|32).
#include "lmc_kernels.cuh"
#include "lmc_launch.h"

#if !defined(LMC_G) || !defined(LMC_WL)
#error "compile with -DLMC_G=<group size> -DLMC_WL=<0|1>"
#endif
#define LMC_CAT2(a, b) a##b
#define LMC_CAT(a, b) LMC_CAT2(a, b)

namespace lmc {

template <bool KONE, bool EWALD, int USHER>
static int launch_one(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_run_kernel<LMC_G, KONE, EWALD, USHER, (LMC_WL != 0)>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

template <bool KONE, bool EWALD>
static int launch_usher(const DevModel& m, const RunArgs& a, int usher, const LaunchCfg& lc) {
  switch (usher) {
    case LMC_USHER_FLIP: return launch_one<KONE, EWALD, LMC_USHER_FLIP>(m, a, lc);
    case LMC_USHER_SWAP: return launch_one<KONE, EWALD, LMC_USHER_SWAP>(m, a, lc);
#if LMC_G == 8 || LMC_G == 32
    case LMC_USHER_TABLEFLIP: return launch_one<KONE, EWALD, LMC_USHER_TABLEFLIP>(m, a, lc);
#endif
    default: return -2;
  }
}

#if LMC_WL
#define LMC_FN LMC_CAT(launch_run_wl_g, LMC_G)
#else
#define LMC_FN LMC_CAT(launch_run_g, LMC_G)
#endif
int LMC_FN(const DevModel& m, const RunArgs& a, bool kone, bool ewald, int usher,
                                 const LaunchCfg& lc) {
  if (kone) return ewald ? launch_usher<true, true>(m, a, usher, lc) : launch_usher<true, false>(m, a, usher, lc);
  return ewald ? launch_usher<false, true>(m, a, usher, lc) : launch_usher<false, false>(m, a, usher, lc);
}

}  // namespace lmc
