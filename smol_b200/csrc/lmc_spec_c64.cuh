// Speculative-batch Metropolis kernel over COMPACT environment words (sm_100a).
//
// lmc_spec.cuh evaluates a flip with three occupancy byte gathers and one table lookup per merged record; on the binary
// FCC configuration that is 66 shared-memory byte loads per flip at 2.4-way bank conflicts, and the L1 / shared-memory
// data pipe is what bounds the kernel (88 % of its peak).  Here every active site keeps the codes its records gather in
// ONE 64-bit word per walker, in shared memory: a record's three codes are a bit field of the word (fields of records
// that share an end site overlap -- that is how 22 records x 3 bits fit 64, see build_c64_tables in lmc_api.cu), so a
// flip costs one 8-byte load plus, per record, a shift, a mask and the table lookup.  An ACCEPTED flip of site s xors
// (old ^ new) into the bits that hold s in the words of the sites that gather it (reverse map, L2); the second flip of
// a swap sees the first through a per-pair mask (L2, issued before the first flip is evaluated).  The words of a
// walker take 8 bytes per active site -- 4 KB for 512 sites, which still leaves 28 walkers resident per SM once the
// per-walker stash shrinks to one slot (the commit evaluates and folds flip by flip, as in lmc_spec_tf.cuh).
//
// Same chains as the gather variant: the same records, table entries and order of summation.
#pragma once
#include "lmc_spec.cuh"

namespace lmc {

// NPAIR record pairs of a lane: e = the site's word, dsc = the lane's record descriptors (table base | shift << 16);
// a0 takes the even records of the lane, a1 the odd ones -- the order spec_rec2 sums them in
template <int EB, bool P2, int NPAIR>
__device__ __forceinline__ void c64_pairs(const double* Dn, unsigned long long e, const uint32_t* dsc, uint32_t NC, double& a0, double& a1) {
  constexpr uint32_t FM = (1u << (3 * EB)) - 1u;
  uint32_t dw[2 * NPAIR];
#pragma unroll
  for (int i = 0; i + 4 <= 2 * NPAIR; i += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(dsc + i);
    dw[i] = v.x; dw[i + 1] = v.y; dw[i + 2] = v.z; dw[i + 3] = v.w;
  }
  if ((2 * NPAIR) & 2) {
    const uint2 v = *reinterpret_cast<const uint2*>(dsc + (2 * NPAIR - 2));
    dw[2 * NPAIR - 2] = v.x; dw[2 * NPAIR - 1] = v.y;
  }
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) {
    a0 += Dn[(dw[2 * p] & 0xffffu) + NC * env_cidx<EB, P2>((uint32_t)(e >> (dw[2 * p] >> 16)) & FM, NC)];
    a1 += Dn[(dw[2 * p + 1] & 0xffffu) + NC * env_cidx<EB, P2>((uint32_t)(e >> (dw[2 * p + 1] >> 16)) & FM, NC)];
  }
}

// scaled energy change of one flip.  NPAIR > 0: that many record pairs per lane as straight-line code (the kernel is
// instantiated for the usual count, three); NPAIR == 0: a loop over the model's count
template <int EB, bool P2, int NPAIR>
__device__ __forceinline__ double c64_flip_energy(const DevModel& m, const unsigned char* smem, const double* dtab,
                                                  unsigned long long e, int site, int oldc, int newc, int l) {
  const uint32_t* dsc = reinterpret_cast<const uint32_t*>(smem + m.off_c64desc) + ((int)smem[m.off_c64cls + site] * 4 + l) * m.c64NRLP;
  const double* Dn = dtab + newc * m.spL + oldc;   // the old code is the fastest index of a block
  const uint32_t NC = P2 ? (1u << EB) : (uint32_t)m.spNC;
  double a0 = 0.0, a1 = 0.0;
  if (NPAIR > 0) {
    c64_pairs<EB, P2, NPAIR ? NPAIR : 1>(Dn, e, dsc, NC, a0, a1);
  } else {
    for (int i0 = 0; i0 < m.c64NRL; i0 += 2) c64_pairs<EB, P2, 1>(Dn, e, dsc + i0, NC, a0, a1);
  }
  return a0 + a1;
}

// k-th (0-based) active position of a sublattice whose code differs from `code` (Swap.propose_step, mcusher.py:190-196),
// from the bit-plane of the code and the exclusive prefix popcounts of its words (u16 per word, kept current by the
// commit): a binary search over the words and a rank select inside one, by every lane on its own -- a third of the
// instructions of the four-lane scan of spec_select_ne and no shuffles
__device__ __forceinline__ int c64_select_ne(const uint32_t* planes, const uint16_t* pfx, int base, int nw, int n_act, int k) {
  int lo = 0, hi = nw - 1;   // largest word with at most k such positions in front of it
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (32 * mid - (int)pfx[base + mid] <= k) lo = mid; else hi = mid - 1;
  }
  const uint32_t tail = (n_act & 31) ? ((1u << (n_act & 31)) - 1u) : 0xffffffffu;
  const uint32_t word = ~planes[base + lo] & (lo == nw - 1 ? tail : 0xffffffffu);
  return lo * 32 + spec_nth_bit(word, k - (32 * lo - (int)pfx[base + lo]));
}

// words of every active site from the occupancy row (whole warp, start of a launch): four lanes per site, each with
// the records of one lane chunk (their loads come from L2 and are independent), combined by shuffles
__device__ __forceinline__ void c64_build(const DevModel& m, const unsigned char* smem, unsigned long long* env, const uint8_t* occ, int g) {
  const uint32_t b = (uint32_t)m.c64B;
  const uint32_t* dbase = reinterpret_cast<const uint32_t*>(smem + m.off_c64desc);
  const int l = g & 3;
  for (int a0 = 0; a0 < m.c64NA; a0 += 8) {   // uniform
    const int ai = a0 + (g >> 2);
    unsigned long long e = 0ull;
    if (ai < m.c64NA) {
      const int site = __ldg(m.sl_sites + ai);
      const uint2* rp = reinterpret_cast<const uint2*>(m.sp_rec + (size_t)site * m.spSb);
      const uint32_t* dsc = dbase + ((int)smem[m.off_c64cls + site] * 4 + l) * m.c64NRLP;
#pragma unroll 6
      for (int i = 0; i < m.c64NRL; ++i) {
        const uint2 rc = __ldg(rp + 2 * (l + 4 * (i >> 1)) + (i & 1));
        const uint32_t field = (uint32_t)occ[rc.x & 0xffffu] | ((uint32_t)occ[rc.x >> 16] << b) | ((uint32_t)occ[rc.y & 0xffffu] << (2u * b));
        e |= (unsigned long long)field << (dsc[i] >> 16);   // (overlapping fields carry the same codes)
      }
    }
    e |= __shfl_xor_sync(0xffffffffu, e, 1);
    e |= __shfl_xor_sync(0xffffffffu, e, 2);
    if (ai < m.c64NA && l == 0) env[ai] = e;
  }
}

// an accepted flip of active site `ai` (x = old ^ new code): the sites that gather it see the new code
__device__ __forceinline__ void c64_commit(const DevModel& m, unsigned long long* env, int ai, uint32_t x, int g) {
  const uint32_t* rv = m.c64Rev + (size_t)ai * m.c64RV;
  uint32_t* e32 = reinterpret_cast<uint32_t*>(env);
  for (int k = g; k < m.c64RV; k += 32) {
    const uint32_t ent = __ldg(rv + k);
    if (ent == 0xffffffffu) continue;
    const uint32_t bit = ent >> 16;
    atomicXor(e32 + 2 * (ent & 0xffffu) + (bit >> 5), x << (bit & 31u));   // (a code never straddles a 32-bit half: bit is a multiple of c64B)
  }
}

// P2: the code radix of the difference table is 2^EB (the field of a record IS its table index); NPAIR: record pairs per
// lane known at compile time (0: read from the model)
template <bool KONE, int USHER, int EB, bool P2, int NPAIR>
__global__ void __launch_bounds__(128, 7) lmc_spec_c64_kernel(const DevModel m, const RunArgs a) {
  static_assert(USHER == LMC_USHER_FLIP || USHER == LMC_USHER_SWAP, "flip / swap only");
  static_assert(EB == 1 || EB == 2, "one or two bits per species code");
  constexpr int SPEC_SG = 4, SPEC_B = 8, G = 32;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint64_t bar;
  const int g = threadIdx.x & 31;
  const int wl_ = threadIdx.x >> 5;                 // walker slot in block
  const int w = blockIdx.x * a.wpb + wl_;
  const int nw_blk = min(a.wpb, a.W - blockIdx.x * a.wpb);
  const bool active = wl_ < a.wpb && w < a.W;
  const int sg = g / SPEC_SG, l = g % SPEC_SG;

  unsigned char* wbase = smem + ((m.blob_c64_bytes + 15) & ~15);
  uint8_t* occ_rows = wbase;
  unsigned char* rest = wbase + (size_t)a.wpb * m.Npad;
  uint8_t* occ = occ_rows + (size_t)wl_ * m.Npad;
  unsigned char* priv = rest + (size_t)wl_ * a.walker_smem;
  double* feat = reinterpret_cast<double*>(priv + a.off_feat);
  unsigned char* stash0 = priv + a.off_stash;       // ONE slot; between commits it holds the ring (a.off_ring == a.off_stash)
  int* cnt = reinterpret_cast<int*>(priv + a.off_cnt);
  uint32_t* planes = reinterpret_cast<uint32_t*>(priv + a.off_plane);
  uint4* ring = reinterpret_cast<uint4*>(priv + a.off_ring);   // [32] x (sl<<24 | pos, site, word z, float log u)
  unsigned long long* env = reinterpret_cast<unsigned long long*>(priv + a.off_env64);
  uint16_t* pfx = reinterpret_cast<uint16_t*>(priv + a.off_eidx);   // set bits in the earlier words of the same plane

  stage_tables(m, smem, &bar, occ_rows, a.occ + (size_t)blockIdx.x * a.wpb * m.Npad, (uint32_t)(nw_blk * m.Npad),
               (uint32_t)m.blob_c64_bytes);
  const SmemTables t = smem_tables(m, smem);
  const double* dtab = reinterpret_cast<const double*>(smem + m.off_dtab);
  const double* ctab = reinterpret_cast<const double*>(smem + m.off_ctab);
  if (!active) return;
  if (g == 0) occ[m.N] = 0;   // pad byte behind the row: the zero code gathered by unused record slots

  for (int f = g; f < m.F; f += G) feat[f] = a.features[(size_t)w * m.F + f];
  double enth = a.enthalpy[w];
  for (int i = g; i < LMC_MAX_SUBLATTICES * LMC_MAX_CODES; i += G) cnt[i] = 0;
  for (int i = g; i < m.plane_words; i += G) planes[i] = 0u;
  __syncwarp();
  for (int sl = 0; sl < m.nSl; ++sl) {
    const int n_act = m.sl_off[sl + 1] - m.sl_off[sl], nw = m.sl_nwords[sl];
    for (int wd = g; wd < nw; wd += G) {
      const int jn = min(32, n_act - 32 * wd);
      for (int b = 0; b < jn; ++b) {
        const int code = occ[site_of_pos(m, sl, wd * 32 + b)];
        planes[m.sl_plane_off[sl] + code * nw + wd] |= 1u << b;
        atomicAdd(&cnt[sl * LMC_MAX_CODES + code], 1);
      }
    }
  }
  c64_build(m, smem, env, occ, g);
  __syncwarp();
  if (USHER == LMC_USHER_SWAP) {
    for (int sl = 0; sl < m.nSl; ++sl) {
      const int nw = m.sl_nwords[sl];
      for (int c = g; c < m.sl_nplanes[sl]; c += G) {
        int run = 0;
        for (int wd = 0; wd < nw; ++wd) {
          pfx[m.sl_plane_off[sl] + c * nw + wd] = (uint16_t)run;
          run += __popc(planes[m.sl_plane_off[sl] + c * nw + wd]);
        }
      }
    }
    __syncwarp();
  }

  const unsigned long long seed = a.seeds[w];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t wid = (uint32_t)(a.walker_base + w);
  const double beta = a.beta[w];
  constexpr bool MU_POSSIBLE = USHER != LMC_USHER_SWAP;
  const double nat_mu = (MU_POSSIBLE && m.muW) ? t.nat[m.muF] : 0.0;

  unsigned long long step = a.step0;
  unsigned long long rbase = step;
  bool ring_valid = false;
  long long nacc_total = 0;
  for (long long s = 0; s < a.S; ++s) {
    int nacc = 0;
    bool accepted = true;
    int it = 0;
    while (it < a.thin) {
      const int nb = min(SPEC_B, a.thin - it);
      if (!ring_valid || step + (unsigned long long)nb > rbase + 32ull) {
        // state-independent part of the next 32 steps, one step per lane
        rbase = step;
        const unsigned long long st_ = step + (unsigned long long)g;
        const U4 bq = philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1);
        const int sl_ = choose_sublattice(m, bq.x);
        const int j_ = (int)mulhi32(bq.y, (uint32_t)(m.sl_off[sl_ + 1] - m.sl_off[sl_]));
        __syncwarp();
        ring[g] = make_uint4((uint32_t)((sl_ << 24) | j_), (uint32_t)site_of_pos(m, sl_, j_), bq.z,
                             __float_as_uint(log_u_float(bq.w)));
        __syncwarp();
        ring_valid = true;
      }
      const bool live = sg < nb;
      const uint4 rq = ring[min((int)(step - rbase) + sg, 31)];
      const int sl = (int)(rq.x >> 24), pos1 = (int)(rq.x & 0xffffffu), site1 = (int)rq.y;
      const float lf = __uint_as_float(rq.w);
      const int ai1 = m.sl_off[sl] + pos1;
      const unsigned long long e1 = env[ai1];

      // ------------------------------ propose (one step per subgroup) -------------------------
      int n = 0, s1, site2 = 0, s2 = 0, pos2 = 0;
      s1 = occ[site1];
      if (USHER == LMC_USHER_FLIP) {
        // Flip.propose_step, mcusher.py:154-170
        const int nc = m.sl_ncodes[sl];
        int ci = (int)mulhi32(rq.z, (uint32_t)(nc - 1));
        int p = nc;
        for (int c = 0; c < nc; ++c) if (m.sl_codes[sl][c] == s1) { p = c; break; }
        if (ci >= p) ++ci;
        s2 = m.sl_codes[sl][ci];
        n = 1;
      } else {
        // Swap.propose_step, mcusher.py:176-200
        const int n_act = m.sl_off[sl + 1] - m.sl_off[sl];
        const int ndiff = n_act - cnt[sl * LMC_MAX_CODES + s1];
        const int k = ndiff > 0 ? (int)mulhi32(rq.z, (uint32_t)ndiff) : 0;
        if (ndiff > 0) pos2 = c64_select_ne(planes, pfx, m.sl_plane_off[sl] + s1 * m.sl_nwords[sl], m.sl_nwords[sl], n_act, k);
        if (ndiff > 0) {
          site2 = site_of_pos(m, sl, pos2);
          s2 = occ[site2];
          n = 2;
        }
      }
      // site 2 sees site 1 already holding s2 (expansion.py:217-229): the bits of its word that hold site 1 flip
      // s1 -> s2.  The mask (L2) is in flight while flip 1 is evaluated.
      unsigned long long e2 = 0ull, pm2 = 0ull;
      if (USHER == LMC_USHER_SWAP) {
        const int ai2 = m.sl_off[sl] + pos2;
        e2 = env[ai2];
        pm2 = __ldcg(m.c64Pair + (size_t)ai2 * m.c64NA + ai1);
      }

      // ------------------------------ evaluate ------------------------------------------------
      double acc = 0.0, dmu = 0.0;
      if (live && n > 0) {
        acc = c64_flip_energy<EB, P2, NPAIR>(m, smem, dtab, e1, site1, s1, s2, l);
        if (USHER == LMC_USHER_SWAP)
          acc += c64_flip_energy<EB, P2, NPAIR>(m, smem, dtab, e2 ^ (pm2 * (unsigned long long)(s1 ^ s2)), site2, s2, s1, l);
      }
      acc += __shfl_xor_sync(FULL, acc, 1);
      acc += __shfl_xor_sync(FULL, acc, 2);
      double dH = acc;
      if (MU_POSSIBLE && m.muW) {
        dmu = mu_of_s(m, ctab, site1, s2, sl) - mu_of_s(m, ctab, site1, s1, sl);
        dH += nat_mu * dmu;
      }

      // ------------------------------ accept (metropolis.py:31-49) ----------------------------
      const double exponent = __dmul_rn(-beta, dH);
      const int af = accept_fast(exponent, lf);
      bool acc_ = af != 0;
      if (af < 0) {
        const unsigned long long st_ = step + (unsigned long long)sg;
        acc_ = exponent > log(u01(philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1).w));
      }
      const uint32_t bal = __ballot_sync(FULL, acc_ && live);
      if (bal == 0u) {
        step += (unsigned long long)nb;
        it += nb;
        accepted = false;
        continue;
      }

      // ------------------------------ commit the first accepted step --------------------------
      const int src = __ffs(bal) - 1;
      const int j = src / SPEC_SG;
      const int c_n = __shfl_sync(FULL, n, src);
      const int c_sl = __shfl_sync(FULL, sl, src);
      const int c_site1 = __shfl_sync(FULL, site1, src), c_s1 = __shfl_sync(FULL, s1, src), c_pos1 = __shfl_sync(FULL, pos1, src);
      const int c_site2 = __shfl_sync(FULL, site2, src), c_s2 = __shfl_sync(FULL, s2, src), c_pos2 = __shfl_sync(FULL, pos2, src);
      const double c_dH = __shfl_sync(FULL, dH, src);
      const double c_dmu = __shfl_sync(FULL, dmu, src);
      // classic record path flip by flip (per-record differences in the reference's cluster order for the feature
      // update, evaluator.pyx:253-263); the second flip of a swap is evaluated with the first written
      for (int f = 0; f < c_n; ++f) {   // uniform
        const int fs = f == 0 ? c_site1 : c_site2, fo = f == 0 ? c_s1 : c_s2, fn = f == 0 ? c_s2 : c_s1;
        const int fp = f == 0 ? c_pos1 : c_pos2;
        const RecChunk pre = load_records<G>(m, fs, g);
        (void)flip_energy<G, KONE>(m, t, occ, fs, fo, fn, stash0, g, pre);
        __syncwarp();
        flip_features<G, KONE>(m, t, fs, stash0, feat, g, load_segment<G>(m, fs, g));
        __syncwarp();
        if (g == 0) {
          occ[fs] = (uint8_t)fn;
          cnt[c_sl * LMC_MAX_CODES + fo]--;
          cnt[c_sl * LMC_MAX_CODES + fn]++;
          const int nw = m.sl_nwords[c_sl];
          uint32_t* pl = planes + m.sl_plane_off[c_sl] + (fp >> 5);
          const uint32_t bit = 1u << (fp & 31);
          pl[fo * nw] ^= bit;
          pl[fn * nw] ^= bit;
        }
        if (USHER == LMC_USHER_SWAP) {   // words behind the position: one set bit fewer in the old code's plane, one more in the new one's
          const int nw = m.sl_nwords[c_sl];
          uint16_t* po = pfx + m.sl_plane_off[c_sl] + fo * nw;
          uint16_t* pn = pfx + m.sl_plane_off[c_sl] + fn * nw;
          for (int wd = (fp >> 5) + 1 + g; wd < nw; wd += G) { po[wd] -= 1; pn[wd] += 1; }
        }
        c64_commit(m, env, m.sl_off[c_sl] + fp, (uint32_t)(fo ^ fn), g);
        __syncwarp();
      }
      if (g == 0 && MU_POSSIBLE && m.muW && c_n > 0) feat[m.muF] += c_dmu;
      __syncwarp();
      if (c_n > 0) ring_valid = false;   // the stash slot of the commit is where the ring lives
      enth += c_dH;
      ++nacc;
      accepted = true;
      step += (unsigned long long)(j + 1);
      it += j + 1;
    }  // thin

    // ------------------------------ sample trace ------------------------------------------
    nacc_total += nacc;
    const size_t sw = (size_t)s * a.W + w;
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
      if (a.tr_enth) a.tr_enth[sw] = enth;
      if (a.tr_acc) a.tr_acc[sw] = accepted ? 1 : 0;
      if (a.tr_nacc) a.tr_nacc[sw] = nacc;
    }
    __syncwarp();
  }

  // ------------------------------ final state ---------------------------------------------
  for (int i = g; i < m.N; i += G) a.occ[(size_t)w * m.Npad + i] = (int8_t)occ[i];
  for (int f = g; f < m.F; f += G) a.features[(size_t)w * m.F + f] = feat[f];
  if (g == 0) {
    a.enthalpy[w] = enth;
    if (a.stats) {
      atomicAdd(a.stats, (unsigned long long)nacc_total);
      atomicAdd(a.stats + 1, (unsigned long long)(a.S * (long long)a.thin));
    }
  }
}

}  // namespace lmc
