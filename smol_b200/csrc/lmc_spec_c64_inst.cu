// Instantiations of the speculative kernel over compact environment words.
#include "lmc_spec_c64.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int USHER, int EB, bool P2, int NPAIR>
static int launch_c64_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_spec_c64_kernel<KONE, USHER, EB, P2, NPAIR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

// three record pairs per lane (24 merged records per site: the FCC and rocksalt cluster sets) as straight-line code,
// any other count through the loop
template <bool KONE, int USHER, int EB, bool P2>
static int launch_c64_n(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  if (m.c64NRL == 6) return launch_c64_k<KONE, USHER, EB, P2, 3>(m, a, lc);
  return launch_c64_k<KONE, USHER, EB, P2, 0>(m, a, lc);
}

template <bool KONE, int USHER>
static int launch_c64_b(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  if (m.c64B == 1) return launch_c64_n<KONE, USHER, 1, true>(m, a, lc);   // two codes, one bit each
  if (m.c64B == 2) return m.spNC == 4 ? launch_c64_n<KONE, USHER, 2, true>(m, a, lc) : launch_c64_n<KONE, USHER, 2, false>(m, a, lc);
  return -2;
}

int launch_spec_c64(const DevModel& m, const RunArgs& a, bool kone, int usher, const LaunchCfg& lc) {
  if (usher == LMC_USHER_FLIP) return kone ? launch_c64_b<true, LMC_USHER_FLIP>(m, a, lc) : launch_c64_b<false, LMC_USHER_FLIP>(m, a, lc);
  if (usher == LMC_USHER_SWAP) return kone ? launch_c64_b<true, LMC_USHER_SWAP>(m, a, lc) : launch_c64_b<false, LMC_USHER_SWAP>(m, a, lc);
  return -2;
}

}  // namespace lmc
