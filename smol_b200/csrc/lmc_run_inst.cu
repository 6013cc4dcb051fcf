// Instantiations of the fused MC kernel for ONE group size (compile with -DLMC_G=4|8|16|32).
#include "lmc_kernels.cuh"
#include "lmc_launch.h"

#if !defined(LMC_G) || !defined(LMC_WL)
#error "compile with -DLMC_G=<group size> -DLMC_WL=<0|1>"
#endif
#define LMC_CAT2(a, b) a##b
#define LMC_CAT(a, b) LMC_CAT2(a, b)

namespace lmc {

template <bool KONE, int EWALD, int USHER>
static int launch_one(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_run_kernel<LMC_G, KONE, EWALD, USHER, (LMC_WL != 0)>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

template <bool KONE, int EWALD>
static int launch_usher(const DevModel& m, const RunArgs& a, int usher, const LaunchCfg& lc) {
  switch (usher) {
    case LMC_USHER_FLIP: return launch_one<KONE, EWALD, LMC_USHER_FLIP>(m, a, lc);
    case LMC_USHER_SWAP: return launch_one<KONE, EWALD, LMC_USHER_SWAP>(m, a, lc);
#if LMC_G == 8 || LMC_G == 32
    case LMC_USHER_TABLEFLIP: return launch_one<KONE, EWALD, LMC_USHER_TABLEFLIP>(m, a, lc);
#endif
#if LMC_G == 32
    case LMC_USHER_COMPOSITE: return launch_one<KONE, EWALD, LMC_USHER_COMPOSITE>(m, a, lc);
    case LMC_USHER_MULTISTEP: return launch_one<KONE, EWALD, LMC_USHER_MULTISTEP>(m, a, lc);
#endif
    default: return -2;
  }
}

#if LMC_WL
#define LMC_FN LMC_CAT(launch_run_wl_g, LMC_G)
#else
#define LMC_FN LMC_CAT(launch_run_g, LMC_G)
#endif
// ewald: 0 none, 1 gathered matrix rows, 2 potential cache (RunArgs::ew_field)
int LMC_FN(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher,
                                 const LaunchCfg& lc) {
  if (kone) return ewald == 2 ? launch_usher<true, 2>(m, a, usher, lc)
                              : (ewald ? launch_usher<true, 1>(m, a, usher, lc) : launch_usher<true, 0>(m, a, usher, lc));
  return ewald == 2 ? launch_usher<false, 2>(m, a, usher, lc)
                    : (ewald ? launch_usher<false, 1>(m, a, usher, lc) : launch_usher<false, 0>(m, a, usher, lc));
}

}  // namespace lmc
