// Device-side building blocks of the lattice-MC engine (sm_100a).
//
// Walker-cooperative design: a *group* of G lanes (G = 1..32, power of two, inside one warp)
// advances ONE walker.  The walker's occupancy string lives in shared memory as int8 codes for
// the whole launch; the static model tables (record classes, tensor tables) are staged to shared
// memory once per block with a TMA bulk copy (cp.async.bulk + mbarrier).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lmc {

// ------------------------------------------------------------------------------------------
// Philox4x32-10, identical to oracle/lmc_oracle.py:philox4x32_10
// ------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ void mulwide(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  uint64_t p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
}

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0, lo0, hi1, lo1;
    mulwide(0xD2511F53u, c0, hi0, lo0);
    mulwide(0xCD9E8D57u, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

__device__ __forceinline__ uint32_t mulhi32(uint32_t r, uint32_t n) { return __umulhi(r, n); }
__device__ __forceinline__ double u01(uint32_t r) { return ((double)r + 0.5) * 2.3283064365386963e-10; }

// ------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier helpers (PTX ISA 8.x, sm_90+; SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}
// global -> shared bulk copy, completion signalled on the mbarrier (bytes: multiple of 16)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// fire-and-forget reductions (REDG: no return value, no scoreboard wait; inside the big kernels `atomicAdd` with
// an unused result was compiled to ATOMG ... RZ, whose L2 round trip the next loop iteration waited for)
__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(void* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// per-thread asynchronous copies global -> shared (LDGSTS): no destination registers and no scoreboard slot, the
// issuing thread sees the data after cp_async_wait_all()
__device__ __forceinline__ void cp_async_8(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// group (sub-warp) collectives
// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ uint32_t group_mask() {
  if (G == 32) return 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31;
  return (uint32_t)((1ull << G) - 1ull) << (lane & ~(G - 1));
}
template <int G>
__device__ __forceinline__ void group_sync(uint32_t mask) {
  if (G > 1) __syncwarp(mask);
}
template <int G>
__device__ __forceinline__ double group_sum(double v, uint32_t mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ int group_sum_i(int v, uint32_t mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ double group_min(double v, uint32_t mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o));
  return v;
}

}  // namespace lmc
