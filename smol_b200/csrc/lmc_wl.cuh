// Warp-specialised Wang-Landau kernel (sm_100a): flip steps, one walker per thread block.
//
// A Wang-Landau chain is ONE dependent sequence per walker (kernel/wanglandau.py:186-266) and the
// configurations of interest run few walkers (BASELINE config 4: 1024), so the classic kernel
// (lmc_kernels.cuh: one warp does everything of a step) leaves the SMs idle behind a ~500-instruction
// serial chain.  Here the work of a step is split by ROLE, software-pipelined over the warps of a block:
//
//   * decision warp(s) -- the critical path only: proposal, cluster-record evaluation of dH, the
//     Wang-Landau accept test against the entropy, the one-byte occupancy write and the entropy
//     increment of the post-step bin (which the NEXT accept test depends on);
//   * bookkeeping warp -- everything the next decision does not depend on, one batch behind: the
//     feature fold of accepted steps (cluster order of evaluator.pyx:253-263), histogram /
//     occurrences / per-bin feature sums, the flatness check, sample traces -- plus the
//     state-independent work of FUTURE steps: counter-based random numbers (Philox) and asynchronous
//     prefetch (cp.async) of the flip sites' cluster records into a ring of shared-memory slots.
//
// NE = 3 adds depth-2 speculation: while warp E0 evaluates step t, E1 evaluates step t+1 against the
// current occupancy (valid if t is rejected) and E2 against the occupancy with t's flip applied (valid if t
// is accepted), so every batch commits TWO steps whatever the acceptance ratio (Wang-Landau accepts
// ~70 % of its flips, too many for the linear speculation of lmc_spec.cuh).  The chain is the
// sequential chain of the reference step for step: per-record differences, summation orders and
// the accept arithmetic are the classic kernel's.
#pragma once
#include "lmc_kernels.cuh"

namespace lmc {

#define LMC_WL2_NSLOT 8    // record slots (steps in flight between prefetch and fold)
#define LMC_WL2_RING 64    // state-independent proposal data of upcoming steps

struct Wl2Layout {
  int off_feat, off_stash, stash_stride, off_ring, off_recs, rec_bytes, slot_recs, slot_segs, off_wl, off_mail, total;
};
// shared-memory slab of one walker behind its occupancy row (host and device agree through this function)
__host__ __device__ inline Wl2Layout wl2_layout(int F, int Rstride, int Sstride, int nb, bool kone, int NE) {
  Wl2Layout L;
  int o = 0;
  L.off_feat = o; o += (F * 8 + 15) & ~15;
  L.stash_stride = Rstride * (kone ? 8 : 4);
  L.off_stash = o; o += 2 * NE * L.stash_stride;          // double buffered by batch parity
  L.off_ring = o; o += LMC_WL2_RING * 16;
  L.slot_recs = Rstride < 128 ? Rstride : 128;            // the lanes' first four records (the rest comes from L2)
  L.slot_segs = Sstride < 32 ? Sstride : 32;              // the lanes' first segment entry
  L.rec_bytes = L.slot_recs * 8 + L.slot_segs * 16;
  L.off_recs = o; o += LMC_WL2_NSLOT * L.rec_bytes;
  L.off_wl = o; o += (nb * 16 + 15) & ~15;                // [entropy nb][histogram nb]
  L.off_mail = o; o += 2 * 128;
  L.total = (o + 15) & ~15;
  return L;
}

struct __align__(16) Wl2Step {
  double enth;      // enthalpy after the step
  double dmu;       // chemical-work change of the step (accepted steps)
  int flags;        // 1 accepted, 2 post-step bin inside the window
  int bin;
  int src;          // decision warp whose stash holds the step's per-record differences
  int slot;         // record slot of the step (segment entries for the fold)
};
struct __align__(16) Wl2Mail {
  int n;            // steps committed by the batch (1 or 2)
  int flags;        // 1 flatness check due after the last step, 2 sample boundary, 4 last batch of the launch
  int pad[2];
  Wl2Step st[2];
  double xd[6];     // (dH, dmu) of the batch's three candidates (decision warps only)
};
static_assert(sizeof(Wl2Mail) <= 128, "mailbox page");

template <bool KONE, int NE>
__global__ void __launch_bounds__(32 * (NE + 1), 7) lmc_wl2_kernel(const DevModel m, const RunArgs a) {
  static_assert(NE == 1 || NE == 3, "one decision warp, or three (depth-2 speculation)");
  constexpr int G = 32;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ double wl_m_sh;
  const int g = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int w = blockIdx.x;
  const int nb = a.wl.num_bins;
  const Wl2Layout L = wl2_layout(m.F, m.Rstride, m.Sstride, nb, KONE, NE);

  uint8_t* occ = smem + ((m.off_dtab + 15) & ~15);
  unsigned char* priv = occ + m.Npad;
  double* feat = reinterpret_cast<double*>(priv + L.off_feat);
  unsigned char* stash_base = priv + L.off_stash;     // [2][NE][stash_stride]
  uint4* ring = reinterpret_cast<uint4*>(priv + L.off_ring);
  unsigned char* recs = priv + L.off_recs;            // [NSLOT]([slot_recs] uint2, [slot_segs] int4)
  double* wlSs = reinterpret_cast<double*>(priv + L.off_wl);
  long long* wlHs = reinterpret_cast<long long*>(wlSs) + nb;
  unsigned char* mail0 = priv + L.off_mail;           // two mailbox pages, 128 bytes apart, by batch parity

  stage_tables(m, smem, &bar, occ, a.occ + (size_t)w * m.Npad, (uint32_t)m.Npad, (uint32_t)m.off_dtab);
  const SmemTables t = smem_tables(m, smem);

  const unsigned long long seed = a.seeds[w];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t wid = (uint32_t)(a.walker_base + w);

  if (wp == NE) {
    // =========================== bookkeeping warp =====================================================
    double* wlS = a.wl.entropy_dev + (size_t)w * nb;
    long long* wlH = reinterpret_cast<long long*>(a.wl.histogram_dev) + (size_t)w * nb;
    long long* wlO = reinterpret_cast<long long*>(a.wl.occurrences_dev) + (size_t)w * nb;
    double* wlM = a.wl.mean_features_dev + (size_t)w * nb * m.F;
    double wl_m = a.wl.mod_factor_dev[w];
    long long wl_cnt = a.wl.steps_counter_dev[w];
    int upd_rem = (int)(wl_cnt % a.wl.update_period);
    int chk_rem = (int)(wl_cnt % a.wl.check_period);
    for (int f = g; f < m.F; f += G) feat[f] = a.features[(size_t)w * m.F + f];
    for (int q = g; q < nb; q += G) { wlSs[q] = __ldcg(wlS + q); wlHs[q] = __ldcg(wlH + q); }
    unsigned ring_hi = (unsigned)a.step0;       // low word of the first step whose ring entry is not yet filled
    const uint32_t step_hi0 = (uint32_t)(a.step0 >> 32);
    auto fill_ring = [&]() {                    // the next 32 steps, one per lane
      const unsigned long long st_ = (((unsigned long long)step_hi0 << 32) | (unsigned long long)(uint32_t)a.step0) +
                                     (unsigned long long)(ring_hi - (unsigned)a.step0) + (unsigned long long)g;
      const U4 bq = philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1);
      const int sl_ = choose_sublattice(m, bq.x);
      const int j_ = (int)mulhi32(bq.y, (uint32_t)(m.sl_off[sl_ + 1] - m.sl_off[sl_]));
      ring[(unsigned)st_ & (LMC_WL2_RING - 1)] = make_uint4((uint32_t)((sl_ << 24) | j_), (uint32_t)site_of_pos(m, sl_, j_),
                                                            bq.z, __float_as_uint(log_u_float(bq.w)));
      ring_hi += 32u;
      __syncwarp();
    };
    fill_ring();
    fill_ring();
    unsigned fetched_hi = (unsigned)a.step0;    // records of steps before it are resident (or in flight)
    const int rs8 = L.slot_recs;
    auto fetch_one = [&](unsigned st_) {        // this lane's share of the records / segment entries of step st_
      const int site = (int)ring[st_ & (LMC_WL2_RING - 1)].y;
      unsigned char* slot = recs + (st_ & (LMC_WL2_NSLOT - 1)) * L.rec_bytes;
      uint2* br = reinterpret_cast<uint2*>(slot) + g;
      const uint2* rp = m.site_rec + (size_t)site * m.Rstride + g;
      if (g < rs8) cp_async_8(br, rp);
      if (g + 32 < rs8) cp_async_8(br + 32, rp + 32);
      if (g + 64 < rs8) cp_async_8(br + 64, rp + 64);
      if (g + 96 < rs8) cp_async_8(br + 96, rp + 96);
      if (g < L.slot_segs) cp_async_16(reinterpret_cast<int4*>(slot + rs8 * 8) + g, m.site_seg + (size_t)site * m.Sstride + g);
    };
    for (; fetched_hi != (unsigned)a.step0 + LMC_WL2_NSLOT; ++fetched_hi) fetch_one(fetched_hi);
    cp_async_commit();
    cp_async_wait_all();
    if (g == 0) wl_m_sh = wl_m;
    __syncthreads();                            // P0: ring, records, features and the WL arrays are staged

    const bool wl_sum = a.wl.reserved != 0;
    unsigned step = (unsigned)a.step0;          // low word of the step index (ring / slot addressing only)
    long long sidx = 0;
    int nacc = 0;
    bool accepted = true;
    double wl_m_traced = wl_m, enth_last = 0.0;
    for (unsigned b = 0;; ++b) {
      __syncthreads();                          // rendezvous b: the batch's mailbox and stashes are published
      const Wl2Mail* mb = reinterpret_cast<const Wl2Mail*>(mail0 + (b & 1u) * 128);
      const int n = mb->n, mflags = mb->flags;
      const unsigned char* stash_b = stash_base + (b & 1u) * NE * L.stash_stride;
#pragma unroll
      for (int i = 0; i < (NE == 3 ? 2 : 1); ++i) {
        if (i < n) {
          const Wl2Step st = mb->st[i];
          accepted = (st.flags & 1) != 0;
          enth_last = st.enth;
          if (accepted) {
            // MCKernel._do_accept_step + trace accumulation (kernel/base.py:327-343, sampler.py:204-207)
            const unsigned char* slot = recs + st.slot * L.rec_bytes;
            const int4 seg0 = g < L.slot_segs ? reinterpret_cast<const int4*>(slot + rs8 * 8)[g] : make_int4(0, 0, 0, -1);
            const int site = (int)ring[step & (LMC_WL2_RING - 1)].y;
            flip_features<G, KONE>(m, t, site, stash_b + st.src * L.stash_stride, feat, g, seg0);
            if (m.muW && g == 0) feat[m.muF] += st.dmu;
            ++nacc;
            __syncwarp();
          }
          // WangLandau._do_post_step, kernel/wanglandau.py:222-266 (the entropy increment itself is on the
          // decision warp: the next accept test reads it)
          if (st.flags & 2) {
            const int bin = st.bin;
            ++wl_cnt;
            if (++upd_rem == a.wl.update_period) upd_rem = 0;
            if (++chk_rem == a.wl.check_period) chk_rem = 0;
            const bool upd = upd_rem == 0;
            if (wl_sum) {
              // update_period == 1: the running mean (x_n + (n-1) M)/n is sum/n -- fire-and-forget reductions
              double* mrow = wlM + (size_t)bin * m.F;
              if (m.F <= G) { if (g < m.F) red_add_f64(mrow + g, feat[g]); }
              else for (int f = g; f < m.F; f += G) red_add_f64(mrow + f, feat[f]);
              if (g == 0) red_add_u64(wlO + bin, 1ull);
            } else {
              const long long total = __ldcg(wlO + bin);
              const double inv = 1.0 / (double)(total + 1);
              for (int f = g; f < m.F; f += G) {
                double* p = wlM + (size_t)bin * m.F + f;
                __stcg(p, inv * (feat[f] + (double)total * __ldcg(p)));
              }
              __syncwarp();
              if (upd && g == 0) __stcg(wlO + bin, total + 1);
              __syncwarp();
            }
            if (upd && g == 0) wlHs[bin] += 1;
            wl_m_traced = wl_m;   // trace.mod_factor is copied before the flatness check (wanglandau.py:251)
            if (chk_rem == 0) {
              // (the decision warps wait in the extra rendezvous below: entropy and histogram are stable)
              __syncwarp();
              int nvis = 0;
              double hsum = 0.0, hmin = 1e300;
              for (int q = g; q < nb; q += G)
                if (wlSs[q] > 0.0) {
                  const double h = (double)wlHs[q];
                  ++nvis; hsum += h; hmin = fmin(hmin, h);
                }
              nvis = group_sum_i<G>(nvis, FULL);
              hsum = group_sum<G>(hsum, FULL);
              hmin = group_min<G>(hmin, FULL);
              if (nvis >= 2 && hmin > a.wl.flatness * (hsum / (double)nvis)) {
                for (int q = g; q < nb; q += G) wlHs[q] = 0ll;
                wl_m = wl_next_mod_factor(a.wl, wl_m);
                __syncwarp();
              }
            }
          } else {
            wl_m_traced = wl_m;
          }
          ++step;
        }
      }
      // state-independent work of future steps: random numbers, record prefetch (dead slots only)
      if ((int)(ring_hi - step) <= 32) fill_ring();
      if (fetched_hi != step + LMC_WL2_NSLOT) {
        fetch_one(fetched_hi); ++fetched_hi;
        if (NE == 3 && fetched_hi != step + LMC_WL2_NSLOT) { fetch_one(fetched_hi); ++fetched_hi; }
      }
      cp_async_commit();
      if (mflags & 2) {
        // ------------------------------ sample trace ------------------------------------------
        const size_t sw = (size_t)sidx * a.W + w;
        if (a.tr_occ) {
          int8_t* dst = a.tr_occ + sw * m.N;
          if ((m.N & 15) == 0) {
            if (g == 0) {
              fence_proxy_async();
              tma_store_1d(dst, occ, (uint32_t)m.N);
              tma_store_commit();
              tma_store_wait_read();
            }
          } else {
            for (int i = g; i < m.N; i += G) dst[i] = (int8_t)occ[i];
          }
        }
        if (a.tr_feat)
          for (int f = g; f < m.F; f += G) a.tr_feat[sw * m.F + f] = feat[f];
        if (g == 0) {
          if (a.tr_enth) a.tr_enth[sw] = enth_last;
          if (a.tr_acc) a.tr_acc[sw] = accepted ? 1 : 0;
          if (a.tr_nacc) a.tr_nacc[sw] = nacc;
          if (a.wl.trace_mod_factor_dev) a.wl.trace_mod_factor_dev[sw] = wl_m_traced;
        }
        if (a.wl.trace_entropy_dev || a.wl.trace_histogram_dev || a.wl.trace_occurrences_dev || a.wl.trace_mean_features_dev) {
          __syncwarp();
          for (int q = g; q < nb; q += G) {
            const long long oc = __ldcg(wlO + q);
            if (a.wl.trace_entropy_dev) __stcs(a.wl.trace_entropy_dev + sw * nb + q, wlSs[q]);
            if (a.wl.trace_histogram_dev) __stcs(reinterpret_cast<long long*>(a.wl.trace_histogram_dev) + sw * nb + q, wlHs[q]);
            if (a.wl.trace_occurrences_dev) __stcs(reinterpret_cast<long long*>(a.wl.trace_occurrences_dev) + sw * nb + q, oc);
          }
          if (a.wl.trace_mean_features_dev) {
            for (int i = g; i < nb * m.F; i += G) {
              double v = __ldcg(wlM + i);
              if (wl_sum) {
                const long long oc = __ldcg(wlO + i / m.F);
                if (oc != 1) v = oc > 0 ? v / (double)oc : 0.0;
              }
              __stcs(a.wl.trace_mean_features_dev + (sw * nb) * m.F + i, v);
            }
          }
        }
        ++sidx;
        nacc = 0;
      }
      // records fetched in this batch are read two batches from now at the earliest (NSLOT slots of slack): only the
      // groups of EARLIER batches must have landed before the next rendezvous -- the newest stays in flight
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      if (mflags & 3) {
        if (g == 0) wl_m_sh = wl_m;
        __syncthreads();                        // extra rendezvous: the decision warps waited for check / trace
      }
      if (mflags & 4) break;
    }
    // ------------------------------ final state ---------------------------------------------
    for (int f = g; f < m.F; f += G) a.features[(size_t)w * m.F + f] = feat[f];
    for (int q = g; q < nb; q += G) { wlS[q] = wlSs[q]; wlH[q] = wlHs[q]; }
    if (g == 0) {
      a.wl.mod_factor_dev[w] = wl_m;
      a.wl.steps_counter_dev[w] = wl_cnt;
    }
    return;
  }

  // ============================= decision warps =========================================================
  // E0 proposes, evaluates and DECIDES step t (and, NE == 3, step t+1 from the candidates of E1 / E2); E1 / E2 only
  // propose and evaluate their candidate of step t+1 and follow E0's verdict through the mailbox.
  const double nat_mu = m.muW ? t.nat[m.muF] : 0.0;
  const double wl_inv_bin = 1.0 / a.wl.bin_size;
  double wl_m = a.wl.mod_factor_dev[w];
  int upd_rem = 0, chk_rem = 0;
  {
    const long long c0 = a.wl.steps_counter_dev[w];
    upd_rem = (int)(c0 % a.wl.update_period);
    chk_rem = (int)(c0 % a.wl.check_period);
  }
  double enth = a.enthalpy[w];
  __syncthreads();                              // P0
  double cur_fb = exact_floordiv(enth - a.wl.min_enthalpy, a.wl.bin_size);
  double s_cur = (cur_fb >= 0.0 && cur_fb < (double)nb) ? wlSs[(int)cur_fb] : 0.0;
  unsigned long long step = a.step0;
  unsigned b = 0;
  for (long long s = 0; s < a.S; ++s) {
    int it = 0;
    while (it < a.thin) {
      Wl2Mail* mb = reinterpret_cast<Wl2Mail*>(mail0 + (b & 1u) * 128);
      const bool two = NE == 3 && it + 1 < a.thin;     // a batch never crosses a sample boundary
      // Flip.propose_step, mcusher.py:154-170, for step t and (NE == 3) for step t+1 in both worlds: t rejected
      // (codes of the current occupancy) and t accepted (site0 reads as new0) -- BEFORE any byte of this batch is
      // written (E0 commits only after the named barrier below).
      const uint4 rq0 = ring[(unsigned)step & (LMC_WL2_RING - 1)];
      const int sl0 = (int)(rq0.x >> 24), site0 = (int)rq0.y;
      const int cur0 = occ[site0];
      int new0;
      {
        int ci = (int)mulhi32(rq0.z, (uint32_t)(m.sl_ncodes[sl0] - 1));
        if (ci >= m.sl_code_pos[sl0][cur0]) ++ci;
        new0 = m.sl_codes[sl0][ci];
      }
      int site1 = 0, cur1r = 0, cur1a = 0, new1r = 0, new1a = 0;
      float lf1 = 0.0f;
      if (NE == 3) {
        const uint4 rq1 = ring[((unsigned)step + 1u) & (LMC_WL2_RING - 1)];
        const int sl1 = (int)(rq1.x >> 24);
        site1 = (int)rq1.y;
        lf1 = __uint_as_float(rq1.w);
        const int ci1 = (int)mulhi32(rq1.z, (uint32_t)(m.sl_ncodes[sl1] - 1));
        cur1r = occ[site1];
        cur1a = site1 == site0 ? new0 : cur1r;
        new1r = m.sl_codes[sl1][ci1 + (ci1 >= m.sl_code_pos[sl1][cur1r] ? 1 : 0)];
        new1a = m.sl_codes[sl1][ci1 + (ci1 >= m.sl_code_pos[sl1][cur1a] ? 1 : 0)];
      }
      // ------------------------------ evaluate this warp's candidate ---------------------------------
      // E0 -> step t; E1 -> step t+1 on the current occupancy; E2 -> step t+1 with t's flip applied (patched
      // reads, the shared row is not written)
      const bool live = wp == 0 || two;
      const unsigned mystep = (unsigned)step + (wp > 0 ? 1u : 0u);
      const int site = wp == 0 ? site0 : site1;
      const int cur = wp == 0 ? cur0 : (wp == 1 ? cur1r : cur1a);
      const int newc = wp == 0 ? new0 : (wp == 1 ? new1r : new1a);
      unsigned char* stash = stash_base + ((b & 1u) * NE + wp) * L.stash_stride;
      double dH = 0.0, dmu = 0.0;
      if (live) {
        const uint2* slot = reinterpret_cast<const uint2*>(recs + (mystep & (LMC_WL2_NSLOT - 1)) * L.rec_bytes);
        RecChunk pre;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = g + u * G;
          pre.r[u] = r < L.slot_recs ? slot[r] : make_uint2(0u, (uint32_t)m.nCls << 16);
        }
        double acc;
        if (NE == 3 && wp == 2) acc = flip_energy_patched<G, KONE>(m, t, occ, site, cur, newc, stash, g, pre, (uint32_t)site0, (uint32_t)new0);
        else acc = flip_energy<G, KONE>(m, t, occ, site, cur, newc, stash, g, pre);
        dH = group_sum<G>(acc, FULL);
        if (m.muW) {
          dmu = __ldg(m.mu + site * m.muW + newc) - __ldg(m.mu + site * m.muW + cur);
          dH += nat_mu * dmu;
        }
      }
      int ncommit = 1, mflags = 0;
      if (NE == 3) {
        if (wp > 0 && g == 0) { mb->xd[wp * 2] = dH; mb->xd[wp * 2 + 1] = dmu; }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * NE) : "memory");     // decision warps: candidates published
      }
      if (wp == 0) {
        // ------------------------------ resolve: sequential accept tests (E0) ---------------------------
        // WangLandau._accept_step, kernel/wanglandau.py:186-202
        ncommit = 0;
        bool prev_acc = false;
#pragma unroll
        for (int i = 0; i < (NE == 3 ? 2 : 1); ++i) {
          if ((i == 0 || two) && !(mflags & 1)) {
            double dHc = dH, dmuc = dmu;
            if (NE == 3 && i == 1) { dHc = mb->xd[prev_acc ? 4 : 2]; dmuc = mb->xd[prev_acc ? 5 : 3]; }
            const int sitec = i == 0 ? site0 : site1;
            const int newcc = i == 0 ? new0 : (prev_acc ? new1a : new1r);
            const float lfc = i == 0 ? __uint_as_float(rq0.w) : lf1;
            const double e_new = enth + dHc;
            bool acc_ = false;
            double new_fb = cur_fb, s_new = s_cur;
            if (!(e_new < a.wl.min_enthalpy || e_new >= a.wl.max_enthalpy)) {
              new_fb = exact_floordiv_inv(e_new - a.wl.min_enthalpy, a.wl.bin_size, wl_inv_bin);
              s_new = new_fb == cur_fb ? s_cur : ((new_fb >= 0.0 && new_fb < (double)nb) ? wlSs[(int)new_fb] : 0.0);
              const double exponent = s_cur - s_new;
              const int af = accept_fast(exponent, lfc);
              const unsigned long long st_ = step + (unsigned long long)i;
              acc_ = af >= 0 ? (af != 0)
                             : exponent > log(u01(philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1).w));
            }
            if (acc_) {
              if (g == 0) occ[sitec] = (uint8_t)newcc;
              enth += dHc;
              cur_fb = new_fb;
              s_cur = s_new;
            }
            const bool valid = cur_fb >= 0.0 && cur_fb < (double)nb;
            if (valid) {
              if (++upd_rem == a.wl.update_period) upd_rem = 0;
              if (++chk_rem == a.wl.check_period) chk_rem = 0;
              if (upd_rem == 0) {
                s_cur += wl_m;
                if (g == 0) wlSs[(int)cur_fb] = s_cur;
              }
              if (chk_rem == 0) mflags |= 1;     // the flatness check follows this step: the batch ends here
            }
            if (g == 0) {
              Wl2Step o;
              o.enth = enth; o.dmu = dmuc;
              o.flags = (acc_ ? 1 : 0) | (valid ? 2 : 0);
              o.bin = valid ? (int)cur_fb : 0;
              o.src = i == 0 ? 0 : (prev_acc ? 2 : 1);
              o.slot = (int)(((unsigned)step + (unsigned)i) & (LMC_WL2_NSLOT - 1));
              mb->st[i] = o;
            }
            prev_acc = acc_;
            ++ncommit;
          }
        }
        if (it + ncommit == a.thin) {
          mflags |= 2;
          if (s == a.S - 1) mflags |= 4;
          if (g == 0) fence_proxy_async();      // occupancy bytes written by this thread -> the trace's bulk store
        }
        if (g == 0) { mb->n = ncommit; mb->flags = mflags; }
      }
      __syncthreads();                          // rendezvous b
      if (NE == 3 && wp > 0) { ncommit = mb->n; mflags = mb->flags; }
      it += ncommit;
      step += (unsigned long long)ncommit;
      if (mflags & 3) {
        __syncthreads();                        // the bookkeeping warp has finished the check / the trace
        if (wp == 0) wl_m = wl_m_sh;
      }
      ++b;
    }
  }
  // ------------------------------ final state ---------------------------------------------
  if (wp == 0) {
    for (int i = g; i < m.N; i += G) a.occ[(size_t)w * m.Npad + i] = (int8_t)occ[i];
    if (g == 0) a.enthalpy[w] = enth;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// lmc_wl3_kernel: the same division of labour with the MERGED records of the speculative kernel (lmc_api.cu,
// build_spec_tables) for cluster-decomposition models: a record gathers three sites and ONE pre-differenced table
// entry carries every cluster among them (FCC S_fcc: 87 clusters -> 22 records per flip), so a whole candidate fits
// eight lanes and ONE decision warp evaluates the three candidates of a depth-2 speculation at once
// (lanes 0-7: step t; 8-15: step t+1 if t is rejected; 16-23: step t+1 if t is accepted, gathers patched).  The
// bookkeeping warp gets the feature change of an accepted flip from the per-feature form of the same table
// (DevModel::spFtab): rows of the records' table entries, which the decision warp leaves in a small stash.
// ------------------------------------------------------------------------------------------------------------------
struct Wl3Layout {
  int off_feat, off_stash, off_ring, off_recs, rec_bytes, off_wl, off_mail, total;
};
__host__ __device__ inline Wl3Layout wl3_layout(int F, int NQ, int nb) {
  Wl3Layout L;
  int o = 0;
  L.off_feat = o; o += (F * 8 + 15) & ~15;
  L.off_stash = o; o += 2 * 3 * NQ * 4;                   // [batch parity][candidate][record] table entry
  L.off_ring = o; o += LMC_WL2_RING * 16;
  L.rec_bytes = NQ * 8;
  L.off_recs = o; o += LMC_WL2_NSLOT * L.rec_bytes;
  L.off_wl = o; o += (nb * 16 + 15) & ~15;
  L.off_mail = o; o += 2 * 128;
  L.total = (o + 15) & ~15;
  return L;
}

// feature change of an accepted flip: sum of its records' rows of the per-feature table.  Lane = (row group, feature):
// FP = features padded to a power of two, RG = 32 / FP groups; group rgp takes the contiguous rows [rgp, rgp + 1) * NQ / RG,
// the groups' partial sums are combined in a fixed order.  `off` = the records' row offsets (entry * F) left by the
// decision warp.
template <int FP, int NQ8>
__device__ __forceinline__ double wl3_dfeat(const double* __restrict__ bp, const uint32_t* __restrict__ off, int NQ, int g) {
  constexpr int RG = 32 / FP;
  const int rgp = g / FP;
  double p = 0.0;
  if (NQ8 > 0 && (NQ8 * 8) % RG == 0) {
    constexpr int ROWS = NQ8 > 0 ? NQ8 * 8 / RG : 1;
    const uint32_t* o = off + rgp * ROWS;
    double v[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) v[k] = __ldg(bp + o[k]);
#pragma unroll
    for (int k = 0; k < ROWS; ++k) p += v[k];
  } else {
    const int rows = (NQ + RG - 1) / RG;
    for (int k = 0; k < rows; ++k) {
      const int r = rgp * rows + k;
      p += r < NQ ? __ldg(bp + off[r]) : 0.0;
    }
  }
#pragma unroll
  for (int x = FP; x < 32; x <<= 1) p += __shfl_xor_sync(0xffffffffu, p, x);
  return p;
}

// NQ8 = merged records per lane of a candidate (spNQ / 8), 0 = run-time loop.  Three warps: decision (0), features (1:
// feature change of accepted flips, per-bin feature sums), Wang-Landau state (2: histogram, flatness check, sample
// traces, random numbers and record prefetch of future steps).
template <int NQ8>
__global__ void __launch_bounds__(96, 7) lmc_wl3_kernel(const DevModel m, const RunArgs a) {
  constexpr int G = 32;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ double wl_m_sh;
  const int g = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int w = blockIdx.x;
  const int nb = a.wl.num_bins;
  const int NQ = NQ8 ? NQ8 * 8 : m.spNQ;
  const Wl3Layout L = wl3_layout(m.F, NQ, nb);

  uint8_t* occ = smem + ((m.blob_bytes + 15) & ~15);
  unsigned char* priv = occ + m.Npad;
  double* feat = reinterpret_cast<double*>(priv + L.off_feat);
  uint32_t* stash_base = reinterpret_cast<uint32_t*>(priv + L.off_stash);
  uint4* ring = reinterpret_cast<uint4*>(priv + L.off_ring);   // (site, new-code table, float log u, sublattice)
  unsigned char* recs = priv + L.off_recs;
  double* wlSs = reinterpret_cast<double*>(priv + L.off_wl);
  long long* wlHs = reinterpret_cast<long long*>(wlSs) + nb;
  unsigned char* mail0 = priv + L.off_mail;

  stage_tables(m, smem, &bar, occ, a.occ + (size_t)w * m.Npad, (uint32_t)m.Npad, (uint32_t)m.blob_bytes);
  const SmemTables t = smem_tables(m, smem);
  const double* dtab = reinterpret_cast<const double*>(smem + m.off_dtab);

  const unsigned long long seed = a.seeds[w];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t wid = (uint32_t)(a.walker_base + w);
  long long* wlO = reinterpret_cast<long long*>(a.wl.occurrences_dev) + (size_t)w * nb;
  double* wlM = a.wl.mean_features_dev + (size_t)w * nb * m.F;
  const bool wl_sum = a.wl.reserved != 0;

  if (wp == 1) {
    // =========================== feature warp =========================================================
    long long wl_cnt = a.wl.steps_counter_dev[w];
    int upd_rem = (int)(wl_cnt % a.wl.update_period);
    for (int f = g; f < m.F; f += G) feat[f] = a.features[(size_t)w * m.F + f];
    // lane = (row group, feature): FP = features padded to a power of two, 32 / FP rows of the table in flight
    int FP = 8;
    while (FP < m.F) FP <<= 1;
    const int f = g & (FP - 1), rgp = g / FP;
    const bool fl = f < m.F;
    __syncthreads();                            // P0
    long long sidx = 0;
    for (unsigned b = 0;; ++b) {
      __syncthreads();                          // rendezvous b
      const Wl2Mail* mb = reinterpret_cast<const Wl2Mail*>(mail0 + (b & 1u) * 128);
      const int n = mb->n, mflags = mb->flags;
      const uint32_t* stash_b = stash_base + (b & 1u) * 3 * NQ;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i < n) {
          const int2 mt = *reinterpret_cast<const int2*>(&mb->st[i].flags);
          const int flags = mt.x & 3, bin = mt.x >> 2, src = mt.y;
          if (flags & 1) {
            // feature change of the accepted flip = sum of its records' rows of the per-feature table: row group rgp
            // takes rows rgp, rgp + RG, ...; the groups' partial sums are combined in a fixed order
            const uint32_t* off = stash_b + (src & 0xff) * NQ;
            const double* bp = m.spFtab + (size_t)(src >> 8) * m.spL * m.F + (fl ? f : 0);
            const double p = FP == 8 ? wl3_dfeat<8, NQ8>(bp, off, NQ, g)
                                     : (FP == 16 ? wl3_dfeat<16, NQ8>(bp, off, NQ, g) : wl3_dfeat<32, NQ8>(bp, off, NQ, g));
            if (fl && rgp == 0) feat[f] += p;
            if (m.muW && g == 0) feat[m.muF] += mb->st[i].dmu;
            __syncwarp();
          }
          if (flags & 2) {
            // WangLandau._do_post_step (wanglandau.py:230-238): cumulative mean features and occurrences of the bin
            if (++upd_rem == a.wl.update_period) upd_rem = 0;
            if (wl_sum) {
              double* mrow = wlM + (size_t)bin * m.F;
              if (m.F <= G) { if (g < m.F) red_add_f64(mrow + g, feat[g]); }
              else for (int q = g; q < m.F; q += G) red_add_f64(mrow + q, feat[q]);
              if (g == 0) red_add_u64(wlO + bin, 1ull);
            } else {
              const long long total = __ldcg(wlO + bin);
              const double inv = 1.0 / (double)(total + 1);
              for (int q = g; q < m.F; q += G) {
                double* pm = wlM + (size_t)bin * m.F + q;
                __stcg(pm, inv * (feat[q] + (double)total * __ldcg(pm)));
              }
              __syncwarp();
              if (upd_rem == 0 && g == 0) __stcg(wlO + bin, total + 1);
              __syncwarp();
            }
          }
        }
      }
      if (mflags & 2) {
        const size_t sw = (size_t)sidx * a.W + w;
        if (a.tr_feat)
          for (int q = g; q < m.F; q += G) a.tr_feat[sw * m.F + q] = feat[q];
        ++sidx;
      }
      if (mflags & 3) {
        __threadfence();                        // this walker's reductions are ordered before the state warp's trace reads
        __syncthreads();                        // first extra rendezvous: features / sums of the batch are complete
        __syncthreads();                        // second: the state warp has finished the check / the traces
      }
      if (mflags & 4) break;
    }
    for (int q = g; q < m.F; q += G) a.features[(size_t)w * m.F + q] = feat[q];
    return;
  }

  if (wp == 2) {
    // =========================== state warp ===========================================================
    double* wlS = a.wl.entropy_dev + (size_t)w * nb;
    long long* wlH = reinterpret_cast<long long*>(a.wl.histogram_dev) + (size_t)w * nb;
    double wl_m = a.wl.mod_factor_dev[w];
    long long wl_cnt = a.wl.steps_counter_dev[w];
    int upd_rem = (int)(wl_cnt % a.wl.update_period);
    int chk_rem = (int)(wl_cnt % a.wl.check_period);
    for (int q = g; q < nb; q += G) { wlSs[q] = __ldcg(wlS + q); wlHs[q] = __ldcg(wlH + q); }
    unsigned ring_hi = (unsigned)a.step0;
    auto fill_ring = [&]() {                    // the next 32 steps, one per lane
      const unsigned long long st_ = a.step0 + (unsigned long long)(ring_hi - (unsigned)a.step0) + (unsigned long long)g;
      const U4 bq = philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1);
      const int sl_ = choose_sublattice(m, bq.x);
      const int j_ = (int)mulhi32(bq.y, (uint32_t)(m.sl_off[sl_ + 1] - m.sl_off[sl_]));
      // Flip.propose_step (mcusher.py:154-170) for every code the site may hold now: 4 bits per current code
      const int nc = m.sl_ncodes[sl_];
      const int ci0 = (int)mulhi32(bq.z, (uint32_t)(nc - 1));
      uint32_t tab = 0u;
      for (int i = 0; i < nc; ++i) tab |= (uint32_t)m.sl_codes[sl_][ci0 + (ci0 >= i ? 1 : 0)] << (4 * m.sl_codes[sl_][i]);
      ring[(unsigned)st_ & (LMC_WL2_RING - 1)] = make_uint4((uint32_t)site_of_pos(m, sl_, j_), tab,
                                                            __float_as_uint(log_u_float(bq.w)), (uint32_t)sl_);
      ring_hi += 32u;
      __syncwarp();
    };
    fill_ring();
    fill_ring();
    unsigned fetched_hi = (unsigned)a.step0;
    auto fetch_one = [&](unsigned st_) {
      const int site = (int)ring[st_ & (LMC_WL2_RING - 1)].x;
      uint2* slot = reinterpret_cast<uint2*>(recs + (st_ & (LMC_WL2_NSLOT - 1)) * L.rec_bytes);
      const uint2* rp = reinterpret_cast<const uint2*>(m.sp_rec + (size_t)site * m.spSb);
      if (NQ8 && NQ8 <= 4) { if (g < NQ) cp_async_8(slot + g, rp + g); }
      else for (int r = g; r < NQ; r += G) cp_async_8(slot + r, rp + r);
    };
    for (; fetched_hi != (unsigned)a.step0 + LMC_WL2_NSLOT; ++fetched_hi) fetch_one(fetched_hi);
    cp_async_commit();
    cp_async_wait_all();
    if (g == 0) wl_m_sh = wl_m;
    __syncthreads();                            // P0

    unsigned step = (unsigned)a.step0;
    long long sidx = 0;
    int nacc = 0;
    bool accepted = true;
    double wl_m_traced = wl_m, enth_last = 0.0;
    for (unsigned b = 0;; ++b) {
      __syncthreads();                          // rendezvous b
      const Wl2Mail* mb = reinterpret_cast<const Wl2Mail*>(mail0 + (b & 1u) * 128);
      const int n = mb->n, mflags = mb->flags;
      bool check = false;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i < n) {
          const int4 mt = *reinterpret_cast<const int4*>(&mb->st[i].flags);
          const int flags = mt.x & 3, bin = mt.x >> 2;
          accepted = (flags & 1) != 0;
          nacc += flags & 1;
          enth_last = __hiloint2double(mt.w, mt.z);
          wl_m_traced = wl_m;   // trace.mod_factor is copied before the flatness check (wanglandau.py:251)
          if (flags & 2) {
            ++wl_cnt;
            if (++upd_rem == a.wl.update_period) upd_rem = 0;
            if (++chk_rem == a.wl.check_period) chk_rem = 0;
            if (upd_rem == 0 && g == 0) wlHs[bin] += 1;
            check = chk_rem == 0;   // (the decision warp ends the batch at a check step)
          }
          ++step;
        }
      }
      if (check) {
        __syncwarp();
        int nvis = 0;
        double hsum = 0.0, hmin = 1e300;
        for (int q = g; q < nb; q += G)
          if (wlSs[q] > 0.0) {
            const double h = (double)wlHs[q];
            ++nvis; hsum += h; hmin = fmin(hmin, h);
          }
        nvis = group_sum_i<G>(nvis, FULL);
        hsum = group_sum<G>(hsum, FULL);
        hmin = group_min<G>(hmin, FULL);
        if (nvis >= 2 && hmin > a.wl.flatness * (hsum / (double)nvis)) {
          for (int q = g; q < nb; q += G) wlHs[q] = 0ll;
          wl_m = wl_next_mod_factor(a.wl, wl_m);
          __syncwarp();
        }
      }
      // state-independent work of future steps: random numbers, record prefetch (dead slots only)
      if ((int)(ring_hi - step) <= 32) fill_ring();
      if (fetched_hi != step + LMC_WL2_NSLOT) {
        fetch_one(fetched_hi); ++fetched_hi;
        if (fetched_hi != step + LMC_WL2_NSLOT) { fetch_one(fetched_hi); ++fetched_hi; }
      }
      cp_async_commit();
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      if (mflags & 3) {
        if (g == 0) wl_m_sh = wl_m;
        __syncthreads();                        // first extra rendezvous: the feature warp's sums are complete
        if (mflags & 2) {
          // ------------------------------ sample trace ------------------------------------------
          const size_t sw = (size_t)sidx * a.W + w;
          if (a.tr_occ) {
            int8_t* dst = a.tr_occ + sw * m.N;
            if ((m.N & 15) == 0) {
              if (g == 0) {
                fence_proxy_async();
                tma_store_1d(dst, occ, (uint32_t)m.N);
                tma_store_commit();
                tma_store_wait_read();
              }
            } else {
              for (int i = g; i < m.N; i += G) dst[i] = (int8_t)occ[i];
            }
          }
          if (g == 0) {
            if (a.tr_enth) a.tr_enth[sw] = enth_last;
            if (a.tr_acc) a.tr_acc[sw] = accepted ? 1 : 0;
            if (a.tr_nacc) a.tr_nacc[sw] = nacc;
            if (a.wl.trace_mod_factor_dev) a.wl.trace_mod_factor_dev[sw] = wl_m_traced;
          }
          if (a.wl.trace_entropy_dev || a.wl.trace_histogram_dev || a.wl.trace_occurrences_dev || a.wl.trace_mean_features_dev) {
            for (int q = g; q < nb; q += G) {
              const long long oc = __ldcg(wlO + q);
              if (a.wl.trace_entropy_dev) __stcs(a.wl.trace_entropy_dev + sw * nb + q, wlSs[q]);
              if (a.wl.trace_histogram_dev) __stcs(reinterpret_cast<long long*>(a.wl.trace_histogram_dev) + sw * nb + q, wlHs[q]);
              if (a.wl.trace_occurrences_dev) __stcs(reinterpret_cast<long long*>(a.wl.trace_occurrences_dev) + sw * nb + q, oc);
            }
            if (a.wl.trace_mean_features_dev) {
              for (int i = g; i < nb * m.F; i += G) {
                double v = __ldcg(wlM + i);
                if (wl_sum) {
                  const long long oc = __ldcg(wlO + i / m.F);
                  if (oc != 1) v = oc > 0 ? v / (double)oc : 0.0;
                }
                __stcs(a.wl.trace_mean_features_dev + (sw * nb) * m.F + i, v);
              }
            }
          }
          ++sidx;
          nacc = 0;
        }
        __syncthreads();                        // second extra rendezvous: check / traces done
      }
      if (mflags & 4) break;
    }
    for (int q = g; q < nb; q += G) { wlS[q] = wlSs[q]; wlH[q] = wlHs[q]; }
    if (g == 0) {
      a.wl.mod_factor_dev[w] = wl_m;
      a.wl.steps_counter_dev[w] = wl_cnt;
    }
    return;
  }

  // ============================= decision warp ==========================================================
  const double nat_mu = m.muW ? t.nat[m.muF] : 0.0;
  const double wl_inv_bin = 1.0 / a.wl.bin_size;
  double wl_m = a.wl.mod_factor_dev[w];
  // steps until the next entropy update / flatness check (counted over steps that end inside the window)
  int upd_left, chk_left;
  {
    const long long c0 = a.wl.steps_counter_dev[w];
    upd_left = a.wl.update_period - (int)(c0 % a.wl.update_period);
    chk_left = a.wl.check_period - (int)(c0 % a.wl.check_period);
  }
  const bool every_step = a.wl.update_period == 1;
  double enth = a.enthalpy[w];
  if (g == 0) occ[m.N] = 0;                     // pad byte behind the row: the zero code gathered by unused record slots
  __syncthreads();                              // P0
  int cur_bin;                                  // bin of the current enthalpy; -1 while the walker is outside the window
  {
    const double fb = exact_floordiv(enth - a.wl.min_enthalpy, a.wl.bin_size);
    cur_bin = (fb >= 0.0 && fb < (double)nb) ? (int)fb : -1;
  }
  double s_cur = cur_bin >= 0 ? wlSs[cur_bin] : 0.0;
  const int cand = g >> 3, l = g & 7;           // candidate of this lane group (3 = spare lanes), lane inside the group
  const uint32_t NC = (uint32_t)m.spNC;
  unsigned long long step = a.step0;
  unsigned b = 0;
  for (long long s = 0; s < a.S; ++s) {
    int it = 0;
    while (it < a.thin) {
      Wl2Mail* mb = reinterpret_cast<Wl2Mail*>(mail0 + (b & 1u) * 128);
      const bool two = it + 1 < a.thin;         // a batch never crosses a sample boundary
      // proposals of step t and of step t+1 in both worlds (t rejected / t accepted)
      const uint4 rq0 = ring[(unsigned)step & (LMC_WL2_RING - 1)];
      const uint4 rq1 = ring[((unsigned)step + 1u) & (LMC_WL2_RING - 1)];
      const uint32_t site0 = rq0.x, site1 = rq1.x;
      const uint32_t cur0 = occ[site0], cur1r = occ[site1];
      const uint32_t new0 = (rq0.y >> (4u * cur0)) & 15u;
      const uint32_t cur1a = site1 == site0 ? new0 : cur1r;
      const uint32_t new1r = (rq1.y >> (4u * cur1r)) & 15u, new1a = (rq1.y >> (4u * cur1a)) & 15u;
      // ------------------------------ evaluate: three candidates, eight lanes each -------------------
      const uint32_t cur = cand == 0 ? cur0 : (cand == 1 ? cur1r : cur1a);
      const uint32_t newc = cand == 0 ? new0 : (cand == 1 ? new1r : new1a);
      const uint32_t ps = cand == 2 ? site0 : 0xffffffffu, pc = new0;      // candidate 2 sees site0 flipped
      double acc = 0.0;
      {
        // (the spare lanes 24-31 repeat candidate 2 without storing: no divergence)
        const uint2* slot = reinterpret_cast<const uint2*>(recs + (((unsigned)step + (cand > 0 ? 1u : 0u)) & (LMC_WL2_NSLOT - 1)) * L.rec_bytes) + l;
        uint32_t* st_out = stash_base + ((b & 1u) * 3 + (cand < 3 ? cand : 2)) * NQ + l;
        const double* Dn = dtab + (size_t)newc * m.spL + cur;
        double a1 = 0.0, a2 = 0.0;
        const int nq8 = NQ8 ? NQ8 : NQ / 8;
#pragma unroll
        for (int qq = 0; qq < nq8; ++qq) {
          const uint2 v = slot[qq * 8];
          const uint32_t s0 = v.x & 0xffffu, s1 = v.x >> 16, s2 = v.y & 0xffffu;
          const uint32_t o0 = occ[s0], o1 = occ[s1], o2 = occ[s2];
          const uint32_t c0 = s0 == ps ? pc : o0, c1 = s1 == ps ? pc : o1, c2 = s2 == ps ? pc : o2;
          const uint32_t idx = (v.y >> 16) + NC * (c0 + NC * (c1 + NC * c2));
          if (cand < 3) st_out[qq * 8] = (idx + cur) * (uint32_t)m.F;      // row offset in the per-feature table
          const double d = Dn[idx];
          if ((qq % 3) == 0) acc += d; else if ((qq % 3) == 1) a1 += d; else a2 += d;   // (fixed order, three chains)
        }
        acc = (acc + a1) + a2;
      }
      acc += __shfl_xor_sync(FULL, acc, 1);
      acc += __shfl_xor_sync(FULL, acc, 2);
      acc += __shfl_xor_sync(FULL, acc, 4);
      double dH0 = __shfl_sync(FULL, acc, 0), dH1 = __shfl_sync(FULL, acc, 8), dH2 = __shfl_sync(FULL, acc, 16);
      double dmu0 = 0.0, dmu1 = 0.0, dmu2 = 0.0;
      if (m.muW) {
        dmu0 = __ldg(m.mu + site0 * m.muW + new0) - __ldg(m.mu + site0 * m.muW + cur0);
        dmu1 = __ldg(m.mu + site1 * m.muW + new1r) - __ldg(m.mu + site1 * m.muW + cur1r);
        dmu2 = __ldg(m.mu + site1 * m.muW + new1a) - __ldg(m.mu + site1 * m.muW + cur1a);
        dH0 += nat_mu * dmu0; dH1 += nat_mu * dmu1; dH2 += nat_mu * dmu2;
      }
      // ------------------------------ resolve: sequential accept tests ----------------------------------
      // WangLandau._accept_step, kernel/wanglandau.py:186-202
      int ncommit = 0, mflags = 0;
      bool prev_acc = false;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if ((i == 0 || two) && !(mflags & 1)) {
          const double dHc = i == 0 ? dH0 : (prev_acc ? dH2 : dH1);
          const double dmuc = i == 0 ? dmu0 : (prev_acc ? dmu2 : dmu1);
          const uint32_t sitec = i == 0 ? site0 : site1;
          const uint32_t newcc = i == 0 ? new0 : (prev_acc ? new1a : new1r);
          const float lfc = __uint_as_float(i == 0 ? rq0.z : rq1.z);
          const double e_new = enth + dHc;
          bool acc_ = false;
          int new_bin = cur_bin;
          double s_new = s_cur;
          if (!(e_new < a.wl.min_enthalpy || e_new >= a.wl.max_enthalpy)) {
            // (inside the window the bin index is inside [0, nb): nb = ceil((max - min) / bin_size); the clamp only
            // guards the table against a rounding surprise)
            new_bin = min((int)exact_floordiv_inv(e_new - a.wl.min_enthalpy, a.wl.bin_size, wl_inv_bin), nb - 1);
            if (new_bin != cur_bin) s_new = wlSs[new_bin];
            const double exponent = s_cur - s_new;
            const int af = accept_fast(exponent, lfc);
            const unsigned long long st_ = step + (unsigned long long)i;
            acc_ = af >= 0 ? (af != 0)
                           : exponent > log(u01(philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1).w));
          }
          if (acc_) {
            if (g == 0) occ[sitec] = (uint8_t)newcc;
            enth += dHc;
            cur_bin = new_bin;
            s_cur = s_new;
          }
          if (cur_bin >= 0) {
            if (every_step || --upd_left == 0) {
              upd_left = a.wl.update_period;
              s_cur += wl_m;
              if (g == 0) wlSs[cur_bin] = s_cur;
            }
            if (--chk_left == 0) { chk_left = a.wl.check_period; mflags |= 1; }
          }
          if (g == 0) {
            // (enthalpy after the step, flags | bin, stash row | new code << 8): one 16-byte store; dmu beside it
            const int meta = (acc_ ? 1 : 0) | (cur_bin >= 0 ? 2 | (cur_bin << 2) : 0);
            const int src = (i == 0 ? 0 : (prev_acc ? 2 : 1)) | ((int)newcc << 8);
            *reinterpret_cast<int4*>(&mb->st[i].flags) = make_int4(meta, src, __double2loint(enth), __double2hiint(enth));
            if (m.muW) mb->st[i].dmu = dmuc;
          }
          prev_acc = acc_;
          ++ncommit;
        }
      }
      it += ncommit;
      step += (unsigned long long)ncommit;
      if (it == a.thin) {
        mflags |= 2;
        if (s == a.S - 1) mflags |= 4;
        if (g == 0) fence_proxy_async();
      }
      if (g == 0) { mb->n = ncommit; mb->flags = mflags; }
      __syncthreads();                          // rendezvous b
      if (mflags & 3) {
        __syncthreads();
        __syncthreads();
        wl_m = wl_m_sh;
      }
      ++b;
    }
  }
  for (int i = g; i < m.N; i += G) a.occ[(size_t)w * m.Npad + i] = (int8_t)occ[i];
  if (g == 0) a.enthalpy[w] = enth;
}

}  // namespace lmc
