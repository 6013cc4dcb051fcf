// Per-group-size launchers (each compiled in its own translation unit: lmc_run_inst.cu -DLMC_G=n).
#pragma once
#include <cuda_runtime.h>
#include "lmc_model.cuh"

namespace lmc {
struct LaunchCfg {
  int grid, threads;
  size_t smem;
  cudaStream_t stream;
};
// ewald: 0 none, 1 gathered matrix rows, 2 potential cache
// return 0 on success, a cudaError_t value on failure, -2 if the combination is not instantiated
int launch_run_g4(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
int launch_run_g8(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
int launch_run_g16(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
int launch_run_g32(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
// speculative-batch Metropolis kernel (flip / swap)
// sg = lanes per speculated step (1, 2 or 4; 1 uses sorted position lists for swaps)
int launch_spec(const DevModel& m, const RunArgs& a, bool kone, int usher, int sg, bool lists, const LaunchCfg& lc);
// Ewald potential cache (ewf) and / or 448-thread blocks (wide), four lanes per step
int launch_spec_x(const DevModel& m, const RunArgs& a, bool kone, int usher, bool ewf, bool wide, const LaunchCfg& lc);
// environment-word variants (RunArgs.env), four lanes per step
int launch_spec_env(const DevModel& m, const RunArgs& a, bool kone, int usher, bool ewf, bool wide, const LaunchCfg& lc);
// compact environment words in shared memory (lmc_spec_c64.cuh): Metropolis flip / swap, four walkers per block
int launch_spec_c64(const DevModel& m, const RunArgs& a, bool kone, int usher, const LaunchCfg& lc);
// speculative table-flip kernel (lmc_spec_tf.cuh): up to 16 walkers per block, ewf = Ewald through the potential cache
int launch_spec_tf(const DevModel& m, const RunArgs& a, bool kone, bool ewf, const LaunchCfg& lc);
// distance processors (Metropolis flip / swap, G = 32)
int launch_run_dist(const DevModel& m, const RunArgs& a, bool kone, int usher, const LaunchCfg& lc);
// Wang-Landau variants
int launch_run_wl_g4(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
int launch_run_wl_g8(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
int launch_run_wl_g16(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
int launch_run_wl_g32(const DevModel& m, const RunArgs& a, bool kone, int ewald, int usher, const LaunchCfg& lc);
// warp-specialised Wang-Landau flip kernel (lmc_wl.cuh): one walker per block, ne decision warps + one bookkeeping warp
int launch_wl2(const DevModel& m, const RunArgs& a, bool kone, int ne, const LaunchCfg& lc);
size_t wl2_smem_bytes(const DevModel& m, int num_bins, int ne);
int launch_wl3(const DevModel& m, const RunArgs& a, const LaunchCfg& lc);
size_t wl3_smem_bytes(const DevModel& m, int num_bins);
}  // namespace lmc
