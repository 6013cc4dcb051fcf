// Instantiations of the speculative table-flip kernel.
#include "lmc_spec_tf.cuh"
#include "lmc_launch.h"

namespace lmc {

template <bool KONE, bool EWF>
static int launch_tf_k(const DevModel& m, const RunArgs& a, const LaunchCfg& lc) {
  auto kern = lmc_spec_tf_kernel<KONE, EWF>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lc.smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<lc.grid, lc.threads, lc.smem, lc.stream>>>(m, a);
  return (int)cudaGetLastError();
}

int launch_spec_tf(const DevModel& m, const RunArgs& a, bool kone, bool ewf, const LaunchCfg& lc) {
  if (kone) return ewf ? launch_tf_k<true, true>(m, a, lc) : launch_tf_k<true, false>(m, a, lc);
  return ewf ? launch_tf_k<false, true>(m, a, lc) : launch_tf_k<false, false>(m, a, lc);
}

}  // namespace lmc
