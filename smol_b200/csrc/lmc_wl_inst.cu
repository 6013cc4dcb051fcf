// Instantiations of the warp-specialised Wang-Landau kernel (lmc_wl.cuh).
#include "lmc_wl.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, int NE>
static int launch_wl2_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_wl2_kernel<KONE, NE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

// ne: decision warps per walker (1, or 3 = depth-2 speculation)
int launch_wl2(const DevModel& m, const RunArgs& a, bool kone, int ne, const LaunchCfg& lc) {
  if (ne == 3) return kone ? launch_wl2_k<true, 3>(m, a, lc) : launch_wl2_k<false, 3>(m, a, lc);
  return kone ? launch_wl2_k<true, 1>(m, a, lc) : launch_wl2_k<false, 1>(m, a, lc);
}

// merged-record variant (cluster-decomposition models with the per-feature table): one decision warp, depth-2 speculation
template <int NQ8>
static int launch_wl3_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_wl3_kernel<NQ8>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}
int launch_wl3(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  switch (m.spNQ / 8) {
    case 1: return launch_wl3_k<1>(m, a, lc);
    case 2: return launch_wl3_k<2>(m, a, lc);
    case 3: return launch_wl3_k<3>(m, a, lc);
    case 4: return launch_wl3_k<4>(m, a, lc);
    default: return launch_wl3_k<0>(m, a, lc);
  }
}
size_t wl3_smem_bytes(const DevModel& m, int num_bins) {
  const Wl3Layout L = wl3_layout(m.F, m.spNQ, num_bins);
  return (((size_t)m.blob_bytes + 15) & ~size_t(15)) + (size_t)m.Npad + (size_t)L.total;
}

size_t wl2_smem_bytes(const DevModel& m, int num_bins, int ne) {
  const Wl2Layout L = wl2_layout(m.F, m.Rstride, m.Sstride, num_bins, m.kone != 0, ne);
  return (((size_t)m.off_dtab + 15) & ~size_t(15)) + (size_t)m.Npad + (size_t)L.total;
}

}  // namespace lmc
