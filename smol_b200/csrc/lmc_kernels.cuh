// Kernels of the lattice-MC engine (sm_100a).  See DESIGN.md for the data layout.
#pragma once
#include <math.h>
#include "lmc_device.cuh"
#include "lmc_model.cuh"

namespace lmc {

// shared-memory view of the staged tables
struct SmemTables {
  const uint4* cls;     // (strides u8x4 [other0, other1, other2, self], atab_off, coef lo, coef hi)
  const double* tabA;   // phase-A table
  const double* nat;    // natural parameters
  const OrbDev* orb;
  const double* qtab;   // distinct charge values of the factorised Ewald matrix (+ 0.0 for vacancies)
};

__device__ __forceinline__ SmemTables smem_tables(const DevModel& m, const unsigned char* base) {
  SmemTables t;
  t.cls = reinterpret_cast<const uint4*>(base);
  t.tabA = reinterpret_cast<const double*>(base + m.off_tabA);
  t.nat = reinterpret_cast<const double*>(base + m.off_nat);
  t.orb = reinterpret_cast<const OrbDev*>(base + m.off_orb);
  t.qtab = reinterpret_cast<const double*>(base + m.off_qtab);
  return t;
}

// Stage the table blob into shared memory with one TMA bulk copy (thread 0 issues, all wait).
// `extra_*` lets the caller piggy-back a second bulk copy (the block's occupancy rows).  `blob_bytes`: the
// classic kernels stage the blob up to the difference table of the speculative kernel (m.off_dtab), which
// they never read -- for multi-species models that table is tens of KB of shared memory per block.
__device__ __forceinline__ void stage_tables(const DevModel& m, unsigned char* smem, uint64_t* bar,
                                             void* extra_dst, const void* extra_src, uint32_t extra_bytes,
                                             uint32_t blob_bytes) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, blob_bytes + extra_bytes);
    tma_load_1d(smem, m.blob, blob_bytes, bar);
    if (extra_bytes) tma_load_1d(extra_dst, extra_src, extra_bytes, bar);
  }
  mbar_wait(bar, 0);
}

// ------------------------------------------------------------------------------------------
// phase A: energy change of ONE flip (site: olda -> newb) against the current occupancy.
// Lanes of the group stride over the site's local cluster records.  Restates
// delta_interactions_from_occupancies / delta_correlations_from_occupancies
// (smol/utils/cluster/evaluator.pyx:211-317) contracted with the natural parameters.
// The per-record differences are stashed for phase B.
// ------------------------------------------------------------------------------------------
struct RecChunk { uint2 r[4]; };  // the first 4*G records of a site, one lane's share

template <int G>
__device__ __forceinline__ RecChunk load_records_pred(const DevModel& m, int site, int g) {
  const uint2* rp = m.site_rec + (size_t)site * m.Rstride;
  RecChunk c;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = g + u * G;
    c.r[u] = r < m.Rstride ? __ldg(rp + r) : make_uint2(0u, (uint32_t)m.nCls << 16);
  }
  return c;
}

// (uniform-branch loads instead of predicated ones were measured slower: they raise the register pressure of
// the 72-register variants, profiles/r01_variants.md)
template <int G>
__device__ __forceinline__ RecChunk load_records(const DevModel& m, int site, int g) {
  return load_records_pred<G>(m, site, g);
}

template <bool KONE>
__device__ __forceinline__ void eval_record(const SmemTables& t, const uint8_t* occ, const uint2 rec, uint32_t oldsh,
                                            uint32_t xmask, void* stash, int r, double& acc) {
  const uint4 ci = t.cls[rec.y >> 16];
  // occupancy bytes of the three other sites + the old code in the self slot; one dp4a each gives
  // the flat tensor index before / after the flip
  const uint32_t o = (uint32_t)occ[rec.x & 0xffffu] | ((uint32_t)occ[rec.x >> 16] << 8) |
                     ((uint32_t)occ[rec.y & 0xffffu] << 16) | oldsh;
  const uint32_t ii = __dp4a(o, ci.x, ci.y);
  const uint32_t ff = __dp4a(o ^ xmask, ci.x, ci.y);
  const double d = t.tabA[ff] - t.tabA[ii];
  if (KONE) {
    acc += __hiloint2double((int)ci.w, (int)ci.z) * d;
    reinterpret_cast<double*>(stash)[r] = d;
  } else {
    acc += d;
    reinterpret_cast<uint32_t*>(stash)[r] = (ff - ci.y) | ((ii - ci.y) << 16);
  }
}

// NU records of ONE flip per lane, fully unrolled (no per-record branches: the record table is
// padded to a multiple of the group size with zero-class records)
template <int G, bool KONE, int NU>
__device__ __forceinline__ void eval_chunk1(const SmemTables& t, const uint8_t* occ, const uint2* rec, uint32_t oldsh,
                                            uint32_t xmask, void* stash, int r0, double& acc) {
#pragma unroll
  for (int u = 0; u < NU; ++u) eval_record<KONE>(t, occ, rec[u], oldsh, xmask, stash, r0 + u * G, acc);
}
// the same for TWO flips interleaved (independent dependency chains -> twice the ILP)
template <int G, bool KONE, int NU>
__device__ __forceinline__ void eval_chunk2(const SmemTables& t, const uint8_t* occ, const uint2* ra, const uint2* rb,
                                            uint32_t sha, uint32_t xa, uint32_t shb, uint32_t xb, void* stasha,
                                            void* stashb, int r0, double& acca, double& accb) {
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    eval_record<KONE>(t, occ, ra[u], sha, xa, stasha, r0 + u * G, acca);
    eval_record<KONE>(t, occ, rb[u], shb, xb, stashb, r0 + u * G, accb);
  }
}

template <int G>
__device__ __forceinline__ void load_chunk(const DevModel& m, const uint2* rp, int r0, uint2* rec) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + u * G;
    rec[u] = r < m.Rstride ? __ldg(rp + r) : make_uint2(0u, (uint32_t)m.nCls << 16);
  }
}

template <int G, bool KONE>
__device__ __forceinline__ void eval_n1(int nu, const SmemTables& t, const uint8_t* occ, const uint2* rec, uint32_t oldsh,
                                        uint32_t xmask, void* stash, int r0, double& acc) {
  switch (nu) {   // warp-uniform
    case 1: eval_chunk1<G, KONE, 1>(t, occ, rec, oldsh, xmask, stash, r0, acc); break;
    case 2: eval_chunk1<G, KONE, 2>(t, occ, rec, oldsh, xmask, stash, r0, acc); break;
    case 3: eval_chunk1<G, KONE, 3>(t, occ, rec, oldsh, xmask, stash, r0, acc); break;
    default: eval_chunk1<G, KONE, 4>(t, occ, rec, oldsh, xmask, stash, r0, acc); break;
  }
}
template <int G, bool KONE>
__device__ __forceinline__ void eval_n2(int nu, const SmemTables& t, const uint8_t* occ, const uint2* ra, const uint2* rb,
                                        uint32_t sha, uint32_t xa, uint32_t shb, uint32_t xb, void* stasha, void* stashb,
                                        int r0, double& acca, double& accb) {
  switch (nu) {
    case 1: eval_chunk2<G, KONE, 1>(t, occ, ra, rb, sha, xa, shb, xb, stasha, stashb, r0, acca, accb); break;
    case 2: eval_chunk2<G, KONE, 2>(t, occ, ra, rb, sha, xa, shb, xb, stasha, stashb, r0, acca, accb); break;
    case 3: eval_chunk2<G, KONE, 3>(t, occ, ra, rb, sha, xa, shb, xb, stasha, stashb, r0, acca, accb); break;
    default: eval_chunk2<G, KONE, 4>(t, occ, ra, rb, sha, xa, shb, xb, stasha, stashb, r0, acca, accb); break;
  }
}

// energy change of one flip; `pre` holds the lane's first four records
template <int G, bool KONE>
__device__ __forceinline__ double flip_energy(const DevModel& m, const SmemTables& t, const uint8_t* occ, int site,
                                              int olda, int newb, void* stash, int g, const RecChunk& pre) {
  const uint32_t oldsh = (uint32_t)olda << 24;
  const uint32_t xmask = (uint32_t)(olda ^ newb) << 24;
  const int nper = m.Rstride / G;   // records per lane (uniform: Rstride is a multiple of 32)
  double acc = 0.0;
  eval_n1<G, KONE>(min(nper, 4), t, occ, pre.r, oldsh, xmask, stash, g, acc);
  const uint2* rp = m.site_rec + (size_t)site * m.Rstride;
  for (int base = 4; base < nper; base += 4) {
    uint2 rec[4];
    load_chunk<G>(m, rp, g + base * G, rec);
    eval_n1<G, KONE>(min(nper - base, 4), t, occ, rec, oldsh, xmask, stash, g + base * G, acc);
  }
  return acc;
}

// the same against the occupancy with site `ps` reading as code `pc` (an earlier flip applied in this warp's view
// only: lmc_wl.cuh evaluates step t+1 "given t accepted" without writing the shared row)
template <bool KONE>
__device__ __forceinline__ void eval_record_patched(const SmemTables& t, const uint8_t* occ, const uint2 rec, uint32_t oldsh,
                                                    uint32_t xmask, void* stash, int r, double& acc, uint32_t ps, uint32_t pc) {
  const uint4 ci = t.cls[rec.y >> 16];
  const uint32_t s0 = rec.x & 0xffffu, s1 = rec.x >> 16, s2 = rec.y & 0xffffu;
  const uint32_t o0 = occ[s0], o1 = occ[s1], o2 = occ[s2];
  const uint32_t o = (s0 == ps ? pc : o0) | ((s1 == ps ? pc : o1) << 8) | ((s2 == ps ? pc : o2) << 16) | oldsh;
  const uint32_t ii = __dp4a(o, ci.x, ci.y);
  const uint32_t ff = __dp4a(o ^ xmask, ci.x, ci.y);
  const double d = t.tabA[ff] - t.tabA[ii];
  if (KONE) {
    acc += __hiloint2double((int)ci.w, (int)ci.z) * d;
    reinterpret_cast<double*>(stash)[r] = d;
  } else {
    acc += d;
    reinterpret_cast<uint32_t*>(stash)[r] = (ff - ci.y) | ((ii - ci.y) << 16);
  }
}
template <int G, bool KONE>
__device__ __forceinline__ double flip_energy_patched(const DevModel& m, const SmemTables& t, const uint8_t* occ, int site,
                                                      int olda, int newb, void* stash, int g, const RecChunk& pre,
                                                      uint32_t ps, uint32_t pc) {
  const uint32_t oldsh = (uint32_t)olda << 24;
  const uint32_t xmask = (uint32_t)(olda ^ newb) << 24;
  const int nper = m.Rstride / G;
  double acc = 0.0;
  const uint2* rp = m.site_rec + (size_t)site * m.Rstride;
  for (int base = 0; base < nper; base += 4) {
    uint2 rec[4];
    if (base == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) rec[u] = pre.r[u];
    } else {
      load_chunk<G>(m, rp, g + base * G, rec);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (base + u < nper) eval_record_patched<KONE>(t, occ, rec[u], oldsh, xmask, stash, g + (base + u) * G, acc, ps, pc);
  }
  return acc;
}

// energy change of a two-flip step (swap).  The caller has already written flip a's new code to
// the occupancy: flip a's records never contain its own site, flip b's records must see flip a
// applied (sequential semantics of expansion.py:217-229), so both can be evaluated interleaved.
template <int G, bool KONE>
__device__ __forceinline__ double flip_energy_pair(const DevModel& m, const SmemTables& t, const uint8_t* occ,
                                                   int sitea, int olda, int newa, int siteb, int oldb, int newb,
                                                   void* stasha, void* stashb, int g, const RecChunk& prea,
                                                   const RecChunk& preb) {
  const uint32_t sha = (uint32_t)olda << 24, xa = (uint32_t)(olda ^ newa) << 24;
  const uint32_t shb = (uint32_t)oldb << 24, xb = (uint32_t)(oldb ^ newb) << 24;
  const int nper = m.Rstride / G;
  double acca = 0.0, accb = 0.0;
  eval_n2<G, KONE>(min(nper, 4), t, occ, prea.r, preb.r, sha, xa, shb, xb, stasha, stashb, g, acca, accb);
  const uint2* rpa = m.site_rec + (size_t)sitea * m.Rstride;
  const uint2* rpb = m.site_rec + (size_t)siteb * m.Rstride;
  for (int base = 4; base < nper; base += 4) {
    uint2 ra[4], rb[4];
    load_chunk<G>(m, rpa, g + base * G, ra);
    load_chunk<G>(m, rpb, g + base * G, rb);
    eval_n2<G, KONE>(min(nper - base, 4), t, occ, ra, rb, sha, xa, shb, xb, stasha, stashb, g + base * G, acca, accb);
  }
  return acca + accb;
}

// phase B: fold the stashed per-record differences of an ACCEPTED flip into the running feature
// vector.  One lane sums one piece of an orbit segment in cluster order (the reference's order,
// evaluator.pyx:253-263).  lmc_model_create cuts long segments into up to four pieces on ADJACENT lanes
// (never across a multiple of four, so they sit in one group for every group size): the head lane adds
// the pieces of the next sg.w lanes (shuffles, still in cluster order) and owns the feature,
// feature += p * (size / J_total).  sg = (first, count, orbit, pieces that follow | -1: not a head).
template <int G>
__device__ __forceinline__ double merge_pieces(double p, int follow, uint32_t gmask) {
  const double q1 = __shfl_down_sync(gmask, p, 1, G), q2 = __shfl_down_sync(gmask, p, 2, G),
               q3 = __shfl_down_sync(gmask, p, 3, G);
  if (follow >= 1) p += q1;
  if (follow >= 2) p += q2;
  if (follow >= 3) p += q3;
  return p;
}
// (called by every lane of the group: the merge shuffles are group wide)
template <int G, bool KONE>
__device__ __forceinline__ void fold_segment(const DevModel& m, const SmemTables& t, const int4 sg, const void* stash,
                                             double* feat, uint32_t gmask) {
  const bool live = sg.y > 0;   // count 0 = padding entry
  const OrbDev& o = t.orb[live ? sg.z : 0];
  if (KONE) {
    const double* d = reinterpret_cast<const double*>(stash) + sg.x;
    double p = 0.0;
    // pieces hold at most LMC_SEG_PIECE clusters unless a segment is longer than four pieces: straight-line
    // chunks with the tail masked (adding +0.0 leaves the sum unchanged)
    for (int j0 = 0; j0 < sg.y; j0 += LMC_SEG_PIECE) {
#pragma unroll
      for (int j = 0; j < LMC_SEG_PIECE; ++j) p += (j0 + j < sg.y) ? d[j0 + j] : 0.0;
    }
    p = merge_pieces<G>(p, sg.w, gmask);
    if (live && sg.w >= 0) feat[o.fidx] += p * o.w;
  } else {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(stash) + sg.x;
    for (int k = 0; k < m.Kmax; ++k) {   // uniform trip count (largest K of the model): shuffles inside
      double p = 0.0;
      const bool on = live && k < o.K;
      if (on) {
        const double* tk = m.ftab + o.ftab_off + k * o.T;
        for (int j = 0; j < sg.y; ++j) p += __ldg(tk + (u[j] & 0xffffu)) - __ldg(tk + (u[j] >> 16));
      }
      p = merge_pieces<G>(p, sg.w, gmask);
      if (on && sg.w >= 0) feat[o.fidx + k] += p * o.w;
    }
  }
}
// the lane's first segment entry of a site (fixed-stride table, count 0 = padding); state independent,
// so it can be fetched together with the records
template <int G>
__device__ __forceinline__ int4 load_segment(const DevModel& m, int site, int g) {
  return g < m.Sstride ? __ldg(m.site_seg + (size_t)site * m.Sstride + g) : make_int4(0, 0, 0, -1);
}
template <int G, bool KONE>
__device__ __forceinline__ void flip_features(const DevModel& m, const SmemTables& t, int site, const void* stash,
                                              double* feat, int g, const int4 seg0) {
  const uint32_t gmask = group_mask<G>();
  fold_segment<G, KONE>(m, t, seg0, stash, feat, gmask);
  for (int base = G; base < m.Sstride; base += G) {   // uniform over the group
    const int sidx = base + g;
    const int4 sg = sidx < m.Sstride ? __ldg(m.site_seg + (size_t)site * m.Sstride + sidx) : make_int4(0, 0, 0, -1);
    fold_segment<G, KONE>(m, t, sg, stash, feat, gmask);
  }
}

// Ewald energy change of one flip (delta_ewald_single_flip, smol/utils/cluster/ewald.pyx:9-59); lanes
// stride over sites.  Two forms:
//  * factorised matrix M[i,j] = q_i q_j K[site_i, site_j] (what an Ewald summation produces; detected
//    and verified entry by entry at model creation):
//        dE = 2 (q_add - q_sub) sum_k q_k K[site, k] + M[add,add] - M[sub,sub]        (K[s,s] = 0)
//    `ecache` holds one byte per site: the index of the site's current charge in a small table
//    (shared memory); the inner loop is LDS.U8 + LDS.64 + one fully coalesced LDG + DFMA;
//  * generic symmetric matrix: `ecache` holds the u16 matrix row of each site's current species
//    (0xFFFF = vacancy) and the loop gathers the two (transposed) rows `add` and `sub`.
template <int G>
__device__ __forceinline__ double flip_ewald(const DevModel& m, const SmemTables& t, const void* ecache, int site,
                                             int olda, int newb, int g) {
  const int add = __ldg(m.ewInds + site * m.ewW + newb);
  const int sub = __ldg(m.ewInds + site * m.ewW + olda);
  if (m.ewK) {
    const uint8_t* qi = reinterpret_cast<const uint8_t*>(ecache);
    const double qa = add >= 0 ? __ldg(m.ewQ + add) : 0.0, qs = sub >= 0 ? __ldg(m.ewQ + sub) : 0.0;
    const double* krow = m.ewK + (size_t)site * m.N;
    // two sites per lane and iteration: one 16-byte load of the row, one 2-byte load of the charge indices
    // (rows are 16-byte aligned for even N; an odd N leaves one trailing site to lane 0)
    double acc0 = 0.0, acc1 = 0.0;
    const int npair = (m.N & 1) ? 0 : (m.N >> 1);
    const double2* krow2 = reinterpret_cast<const double2*>(krow);
    const uint16_t* qi2 = reinterpret_cast<const uint16_t*>(qi);
#pragma unroll 4
    for (int k = g; k < npair; k += G) {
      const uint32_t q2 = qi2[k];
      const double2 kv = __ldg(krow2 + k);
      acc0 += t.qtab[q2 & 0xffu] * kv.x;
      acc1 += t.qtab[q2 >> 8] * kv.y;
    }
#pragma unroll 1
    for (int k = 2 * npair + g; k < m.N; k += G) acc0 += t.qtab[qi[k]] * __ldg(krow + k);   // odd N only
    double accf = acc0 + acc1;
    accf *= 2.0 * (qa - qs);
    if (g == 0) accf += (add >= 0 ? __ldg(m.ewD + add) : 0.0) - (sub >= 0 ? __ldg(m.ewD + sub) : 0.0);
    return accf;
  }
  const uint16_t* eidx = reinterpret_cast<const uint16_t*>(ecache);
  const double* rowA = m.ewMt + (size_t)(add < 0 ? 0 : add) * m.E;
  const double* rowS = m.ewMt + (size_t)(sub < 0 ? 0 : sub) * m.E;
  const double ca = add >= 0 ? 2.0 : 0.0, cs = sub >= 0 ? 2.0 : 0.0;
  double acc = 0.0;
#pragma unroll 8
  for (int k = g; k < m.N; k += G) {
    const uint32_t e = eidx[k];
    const bool skip = (e == 0xffffu) || (k == site);
    const uint32_t ev = skip ? 0u : e;
    const double tk = ca * __ldg(rowA + ev) - cs * __ldg(rowS + ev);   // (2 M[i,add]) - (2 M[j,sub])
    acc += skip ? 0.0 : tk;
  }
  if (g == 0) {
    double tk = 0.0;
    if (add >= 0) tk += __ldg(rowA + add);
    if (sub >= 0) tk -= __ldg(rowS + sub);
    acc += tk;
  }
  return acc;
}
// refresh the per-walker Ewald cache entry of `site` after its code changed
__device__ __forceinline__ void ewald_cache_set(const DevModel& m, void* ecache, int site, int code) {
  const int e = __ldg(m.ewInds + site * m.ewW + code);
  if (m.ewK) reinterpret_cast<uint8_t*>(ecache)[site] = e < 0 ? (uint8_t)m.ewNQ : __ldg(m.ewQidx + e);
  else reinterpret_cast<uint16_t*>(ecache)[site] = e < 0 ? (uint16_t)0xffffu : (uint16_t)e;
}

// Ewald through the potential cache fld[k] = sum_j q_j K[k][j] of the walker's current occupancy:
//   dE(flip k: a -> b) = 2 (q_b - q_a) fld[k] + M[bb] - M[aa]          (K[k][k] = 0)
// and an accepted flip adds (q_b - q_a) K[k][:] to the cache.  Flips of one step are sequential
// (ewald.py:168-181): flip f sees the cache shifted by the earlier flips, dq_h K[site_h][site_f].
__device__ __forceinline__ double2 ewald_qd(const DevModel& m, int site, int code) {
  return __ldg(m.ewQD + site * m.ewW + code);
}
// the same for a site of ACTIVE sublattice `sl` (per-sublattice table in the parameter bank when the rows agree)
__device__ __forceinline__ double2 ewald_qd(const DevModel& m, int site, int code, int sl) {
  return m.qdC ? make_double2(m.qc_c[sl][code], m.qg_c[sl][code]) : __ldg(m.ewQD + site * m.ewW + code);
}
__device__ __forceinline__ double mu_of(const DevModel& m, int site, int code, int sl) {
  return m.muC ? m.mu_c[sl][code] : __ldg(m.mu + site * m.muW + code);
}
// the same from the staged copy of the tables (ct = smem + m.off_ctab): the lanes of a speculative warp look up
// different (sublattice, code) entries in one instruction, which the constant bank would serialise
__device__ __forceinline__ double mu_of_s(const DevModel& m, const double* ct, int site, int code, int sl) {
  return m.muC ? ct[sl * LMC_MAX_CODES + code] : __ldg(m.mu + site * m.muW + code);
}
__device__ __forceinline__ double2 ewald_qd_s(const DevModel& m, const double* ct, int site, int code, int sl) {
  constexpr int T = LMC_MAX_SUBLATTICES * LMC_MAX_CODES;
  return m.qdC ? make_double2(ct[T + sl * LMC_MAX_CODES + code], ct[2 * T + sl * LMC_MAX_CODES + code])
               : __ldg(m.ewQD + site * m.ewW + code);
}

template <int G>
__device__ __forceinline__ int select_pos_scan(const DevModel& m, const uint32_t* planes, int sl, int code, int k, bool ne,
                                          int g, uint32_t mask) {
  const int n_act = m.sl_off[sl + 1] - m.sl_off[sl];
  const int nw = m.sl_nwords[sl];
  const uint32_t* pl = planes + m.sl_plane_off[sl] + code * nw;
  if (G > 1 && nw <= G) {
    // one word per lane: prefix over the popcounts, then every lane tests bits of the owning word
    const uint32_t tl = (n_act & 31) ? ((1u << (n_act & 31)) - 1u) : 0xffffffffu;
    uint32_t b = 0u;
    if (g < nw) {
      b = pl[g];
      if (ne) b = ~b & (g == nw - 1 ? tl : 0xffffffffu);
    }
    const int c = __popc(b);
    int incl = c;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      const int tt = __shfl_up_sync(mask, incl, o, G);
      if (g >= o) incl += tt;
    }
    const int excl = incl - c;
    const uint32_t own = __ballot_sync(mask, (k >= excl) && (k < incl));
    const int src = own ? (__ffs(own) - 1) : (int)(threadIdx.x & 31);
    const uint32_t word = __shfl_sync(mask, b, src);
    const int rem = k - __shfl_sync(mask, excl, src);
    const int wd = src & (G - 1);
    int hit = -1;
#pragma unroll
    for (int bit = 0; bit < 32; bit += G) {
      const int q = bit + g;
      if (((word >> q) & 1u) && __popc(word & ((1u << q) - 1u)) == rem) hit = q;
    }
    const uint32_t hb = __ballot_sync(mask, hit >= 0);
    const int hl = hb ? (__ffs(hb) - 1) : (int)(threadIdx.x & 31);
    hit = __shfl_sync(mask, hit, hl);
    return wd * 32 + hit;
  }
  const int cw = (nw + G - 1) / G;
  const int lo = g * cw, hi = min(lo + cw, nw);
  int cnt = 0;
  const uint32_t tail = (n_act & 31) ? ((1u << (n_act & 31)) - 1u) : 0xffffffffu;
  for (int wd = lo; wd < hi; ++wd) {
    uint32_t b = pl[wd];
    if (ne) b = ~b & (wd == nw - 1 ? tail : 0xffffffffu);
    cnt += __popc(b);
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const int tt = __shfl_up_sync(mask, incl, o, G);
    if (g >= o) incl += tt;
  }
  const int excl = incl - cnt;
  const bool found = (k >= excl) && (k < incl);
  int res = -1;
  if (found) {
    int rem = k - excl;
    for (int wd = lo; wd < hi; ++wd) {
      uint32_t b = pl[wd];
      if (ne) b = ~b & (wd == nw - 1 ? tail : 0xffffffffu);
      const int c = __popc(b);
      if (rem < c) {
        int pos = 0;
        int c16 = __popc(b & 0xffffu);
        if (rem >= c16) { rem -= c16; pos += 16; b >>= 16; }
        int c8 = __popc(b & 0xffu);
        if (rem >= c8) { rem -= c8; pos += 8; b >>= 8; }
        int c4 = __popc(b & 0xfu);
        if (rem >= c4) { rem -= c4; pos += 4; b >>= 4; }
        int c2 = __popc(b & 0x3u);
        if (rem >= c2) { rem -= c2; pos += 2; b >>= 2; }
        if (rem >= (int)(b & 1u)) pos += 1;
        res = wd * 32 + pos;
        break;
      }
      rem -= c;
    }
  }
  if (G > 1) {
    const uint32_t b = __ballot_sync(mask, found);
    const int src = b ? (__ffs(b) - 1) : (int)(threadIdx.x & 31);
    res = __shfl_sync(mask, res, src);
  }
  return res;
}

// (cached prefix popcounts per plane word were measured neutral against the shuffle scan, profiles/r01_variants.md)
template <int G>
__device__ __forceinline__ int select_pos(const DevModel& m, const uint32_t* planes, int sl, int code, int k, bool ne,
                                          int g, uint32_t mask) {
  return select_pos_scan<G>(m, planes, sl, code, k, ne, g, mask);
}

__device__ __forceinline__ int site_of_pos(const DevModel& m, int sl, int pos) {
  return m.sl_first[sl] >= 0 ? m.sl_first[sl] + pos : __ldg(m.sl_sites + m.sl_off[sl] + pos);
}

// Metropolis test `exponent >= 0 or exponent > log(u)` (kernel/metropolis.py:46-48).  `lf` is a
// float logarithm of u (computed ahead of time); with a guard band it decides almost every case,
// inside the band the exact double log is evaluated, so the decision is always the one of the
// double-precision test.
__device__ __forceinline__ float log_u_float(uint32_t r) { return __logf((float)u01(r)); }

// returns 1 accept, 0 reject, -1 undecided (inside the guard band: evaluate the exact double test)
__device__ __forceinline__ int accept_fast(double exponent, float lf) {
  if (exponent >= 0.0) return 1;
  const float ex = (float)exponent;
  const float eps = 1e-5f * (1.0f + fabsf(lf)) + 1e-6f * fabsf(ex);
  if (ex < lf - eps) return 0;
  if (ex > lf + eps) return 1;
  return -1;
}

__device__ __forceinline__ int choose_sublattice(const DevModel& m, uint32_t r0) {
  if (m.nSl == 1) return 0;
  const double u = u01(r0);
  int s = 0;
  while (s < m.nSl - 1 && !(m.sl_cum[s] > u)) ++s;
  return s;
}

// feasibility mask * weights of the 2*nf flip directions (utils/math.py:832-867).  Called by every lane of the
// group with the same counts: lane i checks direction i, the weights are exchanged with shuffles and summed in
// direction order (the redundant all-lanes version was 13 % of the instructions of a table-flip step).
template <int G>
__device__ __forceinline__ double tf_masked_weights(const DevModel& m, const int* n, double* wout, int g, uint32_t gmask) {
  const int ndir = 2 * m.tfNF;
  double sum = 0.0;
  for (int i0 = 0; i0 < ndir; i0 += G) {   // uniform
    const int i = min(i0 + g, ndir - 1);
    const int sgn = (i & 1) ? -1 : 1;
    bool ok = true;
#pragma unroll 1
    for (int d = 0; d < m.tfD; ++d) {
      const int v = n[d] + sgn * m.tf_table[i >> 1][d];
      ok = ok && v >= 0 && v <= m.tf_max_n[d];
    }
    const double w = ok ? m.tf_w[i] : 0.0;
    const int cnt = min(G, ndir - i0);
#pragma unroll 1
    for (int j = 0; j < cnt; ++j) {
      const double wj = __shfl_sync(gmask, w, j, G);
      wout[i0 + j] = wj;
      sum += wj;
    }
  }
  return sum;
}

template <int MF>
struct Step {
  int n;                       // number of flips (0 = empty step)
  int site[MF], oldc[MF], newc[MF], sl[MF], pos[MF];
  double log_priori;
};
// append without dynamic indexing (keeps the arrays in registers)
template <int MF>
__device__ __forceinline__ void push_flip(Step<MF>& st, int site, int oldc, int newc, int sl, int pos) {
#pragma unroll
  for (int i = 0; i < MF; ++i)
    if (i == st.n) { st.site[i] = site; st.oldc[i] = oldc; st.newc[i] = newc; st.sl[i] = sl; st.pos[i] = pos; }
  if (st.n < MF) ++st.n;
}

// floor(a / b) of the EXACT quotient for b > 0: candidate from the rounded division, corrected with the
// exact FMA remainder.  CPython's float `//` (fmod based, used by WangLandau._get_bin_id) returns the
// same value; this form needs no iterative fmod.
__device__ __forceinline__ double exact_floordiv(double a, double b) {
  double q = floor(a / b);
  const double r = fma(-q, b, a);
  if (r < 0.0) q -= 1.0;
  else if (r >= b) q += 1.0;
  return q;
}
// the same with the candidate from a multiplication by 1/b (within one of the exact floor; the remainder
// test is exact, so the result is the same value): no double division on the step's critical path
__device__ __forceinline__ double exact_floordiv_inv(double a, double b, double inv_b) {
  double q = floor(a * inv_b);
  const double r = fma(-q, b, a);
  if (r < 0.0) q -= 1.0;
  else if (r >= b) q += 1.0;
  return q;
}
// st.<field>[f] for a RUNTIME f without dynamic indexing (keeps the arrays in registers)
template <int MF>
__device__ __forceinline__ int pick(const int (&arr)[MF], int f) {
  int v = arr[0];
#pragma unroll
  for (int i = 1; i < MF; ++i) v = (f == i) ? arr[i] : v;
  return v;
}

// next modification factor after a flat histogram (wanglandau.py:253-264): m / mod_update, or the successor of m in the
// host-tabulated sequence of a callable mod_update (rare path: a linear search over a short table)
__device__ __forceinline__ double wl_next_mod_factor(const LmcWangLandau& wl, double m) {
  if (wl.mod_table_dev == nullptr || wl.mod_table_len <= 0) return m / wl.mod_update;
  for (int i = 0; i + 1 < wl.mod_table_len; ++i)
    if (__ldg(wl.mod_table_dev + i) == m) return __ldg(wl.mod_table_dev + i + 1);
  return __ldg(wl.mod_table_dev + wl.mod_table_len - 1);
}

// Wang-Landau per-walker arrays: plain loads from the shared-memory copy, L2 loads (the global arrays are
// updated with reductions that bypass L1) otherwise
template <typename T>
__device__ __forceinline__ T wl_load(const T* smem_copy, const T* global, int i, bool in_smem) {
  return in_smem ? smem_copy[i] : __ldcg(global + i);
}

// Python float floor division `a // b` (CPython float_floor_div), used by WangLandau._get_bin_id
__device__ __forceinline__ double py_floordiv(double a, double b) {
  double mod = fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
  if (div != 0.0) {
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
  }
  return copysign(0.0, a / b);
}

// ------------------------------------------------------------------------------------------
// the fused MC kernel: propose -> delta features/energy (+Ewald, +mu) -> accept -> update,
// num_samples * thin_by attempted steps per walker in ONE launch.
// ------------------------------------------------------------------------------------------
// EWMODE: 0 no Ewald term, 1 matrix rows gathered at every flip (flip_ewald), 2 potential cache (ewald_qd).
// A template parameter, not a run-time switch: the table-flip variants are instruction-fetch bound and
// carry one Ewald path each.
// DIST: distance processor (processor/distance.py): the running features are the distance vector
// [L, |f_i - target_i| ...]; every proposal folds its per-record differences into the change of the
// correlation / interaction vector f (phase B for each proposal, not only on accept) to get the new distances.
template <int G, bool KONE, int EWMODE, int USHER, bool WLMODE, bool DIST = false>
#ifndef LMC_TF_MINB
#define LMC_TF_MINB 2   // resident 256-thread blocks per SM the table-flip + Ewald variants are compiled for (register cap)
#endif
__global__ void __launch_bounds__((EWMODE || WLMODE || USHER >= LMC_USHER_TABLEFLIP) ? 256 : 128,
                                  (USHER == LMC_USHER_TABLEFLIP && EWMODE != 0 && !WLMODE) ? LMC_TF_MINB :
                                  ((EWMODE || WLMODE || USHER >= LMC_USHER_TABLEFLIP) ? 2 : (G < 32 ? 5 : 7)))
lmc_run_kernel(const DevModel m, const RunArgs a) {
  constexpr bool EWALD = EWMODE != 0, EWGATHER = EWMODE == 1, EWFIELD = EWMODE == 2;
  constexpr bool COMP = USHER == LMC_USHER_COMPOSITE;   // flip / swap sub-ushers picked per step
  constexpr bool MULTI = USHER == LMC_USHER_MULTISTEP;  // chained proposals of one flip / swap sub-usher
  constexpr int MF = USHER == LMC_USHER_FLIP ? 1 : ((USHER == LMC_USHER_SWAP || COMP) ? 2 : LMC_MAX_FLIPS);
  constexpr int I1 = MF > 1 ? 1 : 0;   // index of the second flip (dead code when MF == 1)
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint64_t bar;
  const int g = threadIdx.x % G;
  const int wl_ = threadIdx.x / G;                  // walker slot in block
  const int w = blockIdx.x * a.wpb + wl_;            // walker on this device
  const int nw_blk = min(a.wpb, a.W - blockIdx.x * a.wpb);
  const bool active = wl_ < a.wpb && w < a.W && (a.mask == nullptr || a.mask[min(w, a.W - 1)] != 0);
  const uint32_t gmask = group_mask<G>();

  unsigned char* wbase = smem + ((m.off_dtab + 15) & ~15);
  // occupancy rows of the block are contiguous in global memory: slab layout keeps the rows of
  // all walkers first ([wpb][Npad]) so that ONE bulk copy loads them.
  uint8_t* occ_rows = wbase;
  unsigned char* rest = wbase + (size_t)a.wpb * m.Npad;
  uint8_t* occ = occ_rows + (size_t)wl_ * m.Npad;
  unsigned char* priv = rest + (size_t)wl_ * a.walker_smem;
  double* feat = reinterpret_cast<double*>(priv + a.off_feat);
  unsigned char* stash0 = priv + a.off_stash;
  int* cnt = reinterpret_cast<int*>(priv + a.off_cnt);
  uint32_t* planes = reinterpret_cast<uint32_t*>(priv + a.off_plane);
  void* eidx = priv + a.off_eidx;   // per-walker Ewald cache (see flip_ewald)
  double* fld = EWFIELD ? a.ew_field + (size_t)w * m.N : nullptr;   // potential cache (global / L2)

  stage_tables(m, smem, &bar, occ_rows, a.occ + (size_t)blockIdx.x * a.wpb * m.Npad, (uint32_t)(nw_blk * m.Npad),
               (uint32_t)m.off_dtab);
  const SmemTables t = smem_tables(m, smem);
  if (!active) return;

  const int stash_stride = m.Rstride * (KONE ? 8 : 4);
  // running state
  for (int f = g; f < m.F; f += G) feat[f] = a.features[(size_t)w * m.F + f];
  double enth = a.enthalpy[w];
  if (EWGATHER) {
    for (int i = g; i < m.N; i += G) ewald_cache_set(m, eidx, i, occ[i]);
  }
  // species counts per (active sublattice, code) and one bit-plane per code
  for (int i = g; i < LMC_MAX_SUBLATTICES * LMC_MAX_CODES; i += G) cnt[i] = 0;
  for (int i = g; i < m.plane_words; i += G) planes[i] = 0u;
  group_sync<G>(gmask);
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
  group_sync<G>(gmask);
  const unsigned long long seed = a.seeds[w];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t wid = (uint32_t)(a.walker_base + w);
  constexpr bool wl_mode = WLMODE;
  const double beta = wl_mode ? 0.0 : a.beta[w];
  const double acc_off = (!WLMODE && a.acc_off) ? a.acc_off[w] : 0.0;
  const double nat_ew = EWALD ? t.nat[m.ewF] : 0.0;
  const double nat_mu = m.muW ? t.nat[m.muF] : 0.0;

  // Wang-Landau per-walker state
  double* wlS = nullptr; long long* wlH = nullptr; long long* wlO = nullptr; double* wlM = nullptr;
  double wl_m = 0.0, wl_m_traced = 0.0; long long wl_cnt = 0;
  double cur_fb = -1.0, s_cur = 0.0;   // current bin (floor value) and its entropy, kept in registers
  const bool wl_sum = a.wl.reserved != 0;  // mean_features buffer holds per-bin SUMS (update_period == 1)
  const int nb = a.wl.num_bins;
  const double wl_inv_bin = WLMODE ? 1.0 / a.wl.bin_size : 0.0;
  // entropy and histogram of the walker live in its shared-memory slab while they fit (a.off_wl >= 0,
  // lmc_run): the entropy of the proposed bin is on the critical path of every step
  const bool wl_sm = WLMODE && a.off_wl >= 0;
  // (shared-memory copies addressed from the slab pointer, so that they compile to LDS / STS rather than generic
  // accesses that wait on the long scoreboard)
  double* wlSs = reinterpret_cast<double*>(priv + (WLMODE ? max(a.off_wl, 0) : 0));
  long long* wlHs = reinterpret_cast<long long*>(wlSs) + a.wl.num_bins;
  int upd_rem = 0, chk_rem = 0;   // wl_cnt modulo update_period / check_period, carried along (no 64-bit division per step)
  if (wl_mode) {
    wlS = a.wl.entropy_dev + (size_t)w * nb;
    wlH = reinterpret_cast<long long*>(a.wl.histogram_dev) + (size_t)w * nb;
    wlO = reinterpret_cast<long long*>(a.wl.occurrences_dev) + (size_t)w * nb;
    wlM = a.wl.mean_features_dev + (size_t)w * nb * m.F;
    wl_m = a.wl.mod_factor_dev[w];
    wl_cnt = a.wl.steps_counter_dev[w];
    upd_rem = (int)(wl_cnt % a.wl.update_period);
    chk_rem = (int)(wl_cnt % a.wl.check_period);
    if (wl_sm) {
      for (int b = g; b < nb; b += G) { wlSs[b] = __ldcg(wlS + b); wlHs[b] = __ldcg(wlH + b); }
      group_sync<G>(gmask);
    }
    cur_fb = exact_floordiv(enth - a.wl.min_enthalpy, a.wl.bin_size);
    s_cur = (cur_fb >= 0.0 && cur_fb < (double)nb) ? wl_load(wlSs, wlS, (int)cur_fb, wl_sm) : 0.0;
  }

  double* dvec = reinterpret_cast<double*>(priv + a.off_dist);   // DIST: [vector F][delta F][new distances F]
  if (DIST) {
    for (int f = g; f < m.F; f += G) dvec[f] = a.dist_vec[(size_t)w * m.F + f];
    group_sync<G>(gmask);
  }
  // bias term of the exponent (bias.py; Metropolis only): running value and table sum, all lanes alike
  // (kept in the walker's shared-memory slab, not in registers: the swap / flip variants are register capped)
  double* bstate = reinterpret_cast<double*>(priv + a.off_bias);
  if (!WLMODE && a.bias_mode) {
    if (g == 0) {
      bstate[0] = a.bias[w];
      for (int r = 0; r < a.bias_rows; ++r) bstate[1 + r] = a.bias_sum[(size_t)w * a.bias_rows + r];
    }
    group_sync<G>(gmask);
  }

  unsigned long long step = a.step0;
  // State-independent part of the next G steps, one step per lane (counter-based RNG): random
  // words, sublattice, first site and the float log of the acceptance uniform.  Every step then
  // costs a few shuffles instead of a redundant Philox evaluation in all lanes.
  uint4* ring = reinterpret_cast<uint4*>(priv + a.off_ring);   // [G] x (sl<<24 | pos, site, word z, float log u)
  int bphase = 0;
  double* tfc = reinterpret_cast<double*>(priv + a.off_tfc);   // table-flip proposal tables of the walker's current counts
  bool tfc_valid = false;
  // latency-bound variants (few resident warps: Wang-Landau flips): records of step t + 1 are fetched during step t
  constexpr bool PREF = WLMODE && USHER == LMC_USHER_FLIP && !DIST;
  // (asynchronous copies into the walker's slab, each lane its own records and segment entry: a register
  // prefetch was measured useless -- its loads share a scoreboard slot with the next waits of the step)
  uint2* nxt_rec = reinterpret_cast<uint2*>(priv + a.off_pref);                   // [2][Rstride]
  int4* nxt_seg = reinterpret_cast<int4*>(priv + a.off_pref + 2 * m.Rstride * 8);   // [2][Sstride]
  bool have_nxt = false;
  int nxt_buf = 0;
  long long nacc_total = 0;
  for (long long s = 0; s < a.S; ++s) {
    int nacc = 0;
    bool accepted = true;
    for (int it = 0; it < a.thin; ++it, ++step) {
      U4 r;
      float lf;
      int pre_sl = 0, pre_j = 0, pre_site = 0;
      if (USHER == LMC_USHER_TABLEFLIP || COMP || MULTI) {
        r = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 0u, wid, k0, k1);
        lf = log_u_float(r.w);
      } else {
        if (bphase == 0) {
          const unsigned long long st_ = step + (unsigned long long)g;
          const U4 bq = philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1);
          const int sl_ = choose_sublattice(m, bq.x);
          const int j_ = (int)mulhi32(bq.y, (uint32_t)(m.sl_off[sl_ + 1] - m.sl_off[sl_]));
          group_sync<G>(gmask);   // every lane has consumed the previous batch
          ring[g] = make_uint4((uint32_t)((sl_ << 24) | j_), (uint32_t)site_of_pos(m, sl_, j_), bq.z,
                               __float_as_uint(log_u_float(bq.w)));
          group_sync<G>(gmask);
        }
        const uint4 rq = ring[bphase];   // same address in all lanes: one broadcast load
        pre_sl = (int)(rq.x >> 24); pre_j = (int)(rq.x & 0xffffffu);
        pre_site = (int)rq.y;
        r.x = 0u; r.y = 0u; r.w = 0u;
        r.z = rq.z;
        lf = __uint_as_float(rq.w);
        bphase = (bphase + 1) & (G - 1);
      }
      Step<MF> st;
      st.n = 0;
      st.log_priori = 0.0;
#pragma unroll
      for (int i = 0; i < MF; ++i) { st.site[i] = 0; st.oldc[i] = 0; st.newc[i] = 0; st.sl[i] = 0; st.pos[i] = 0; }

      // ------------------------------ propose ------------------------------------------
      int usher = USHER;
      uint32_t q0 = r.x, q1 = r.y, q2 = r.z;   // words of the simple ushers
      U4 r1blk{0, 0, 0, 0};
      int ms_toggled = 0;   // flips of chained swaps whose plane bits are temporarily toggled
      int tf_idx = -1;
      if (USHER == LMC_USHER_TABLEFLIP) {
        // TableFlip.propose_step, mcusher.py:553-639
        r1blk = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 1u, wid, k0, k1);
        if (!tfc_valid) {
          int nd[LMC_MAX_DIMS];
          for (int d = 0; d < m.tfD; ++d)
            nd[d] = m.tf_dim_sl[d] >= 0 ? cnt[m.tf_dim_sl[d] * LMC_MAX_CODES + m.tf_dim_code[d]] : 0;
          // The direction weights (utils/math.py:832-867), their cumulative probabilities and the a-priori factor of
          // every direction (mcusher.py:656-711) depend on the species COUNTS only, which change when a table flip is
          // accepted: evaluated here once per such change (the same expressions, in the same order) and kept in the
          // walker's slab -- [weights 16][cumulative 16][log a-priori factor 16][sum].
          double tw[2 * LMC_MAX_TABLE_FLIPS];
          const double sum = tf_masked_weights<G>(m, nd, tw, g, gmask);
          group_sync<G>(gmask);
          if (g == 0) {
            double cum = 0.0;
            for (int i = 0; i < 2 * m.tfNF; ++i) {
              cum += tw[i] / sum;
              tfc[i] = tw[i];
              tfc[2 * LMC_MAX_TABLE_FLIPS + i] = cum;
            }
            tfc[6 * LMC_MAX_TABLE_FLIPS] = sum;
          }
          for (int i = 0; i < 2 * m.tfNF; ++i) {   // uniform over the group
            double lfi = 0.0;
            if (sum > 0.0 && tw[i] > 0.0) {
              const int sgn_i = (i & 1) ? -1 : 1;
              const int* urow_i = m.tf_table[i >> 1];
              int nn[LMC_MAX_DIMS];
              for (int d = 0; d < m.tfD; ++d) nn[d] = nd[d] + sgn_i * urow_i[d];
              double tw2[2 * LMC_MAX_TABLE_FLIPS];
              const double sum2 = tf_masked_weights<G>(m, nn, tw2, g, gmask);
              const double p_now = (1.0 - m.tf_sw) * tw[i] / sum;
              const double p_next = (1.0 - m.tf_sw) * tw2[i ^ 1] / sum2;
              lfi = log(p_next / p_now);
              for (int d = 0; d < m.tfD; ++d)
                if (urow_i[d] != 0) lfi += __ldg(m.lgam + nd[d]) - __ldg(m.lgam + nn[d]);   // ln n! table (gammaln(n+1))
            }
            if (g == 0) tfc[4 * LMC_MAX_TABLE_FLIPS + i] = lfi;
          }
          group_sync<G>(gmask);
          tfc_valid = true;
        }
        const double tfsum = tfc[6 * LMC_MAX_TABLE_FLIPS];
        const bool do_swap = u01(r.x) < m.tf_sw || !(tfsum > 0.0);
        if (do_swap) {
          usher = LMC_USHER_SWAP;
          q0 = r1blk.x; q1 = r1blk.y; q2 = r1blk.z;
        } else {
          // choose_section_from_partition, utils/math.py:870-893
          const double u = u01(r.y);
          tf_idx = 2 * m.tfNF - 1;
          for (int i = 0; i < 2 * m.tfNF; ++i)
            if (tfc[2 * LMC_MAX_TABLE_FLIPS + i] > u && tfc[i] > 0.0) { tf_idx = i; break; }
        }
      }
      if (COMP) {
        // Composite.propose_step, mcusher.py:392-394: word 4 of the step picks the sub-usher, which then draws
        // the sublattice by its own probabilities and the site (words 0, 1)
        r1blk = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 1u, wid, k0, k1);
        const double uc = u01(r1blk.x);
        int ci = 0;
        while (ci < a.comp_num - 1 && !(a.comp_cum[ci] > uc)) ++ci;
        usher = a.comp_usher[ci];
        const double us = u01(r.x);
        int s = 0;
        while (s < m.nSl - 1 && !(a.comp_sl_cum[ci][s] > us)) ++s;
        pre_sl = s;
        pre_j = (int)mulhi32(r.y, (uint32_t)(m.sl_off[s + 1] - m.sl_off[s]));
        pre_site = site_of_pos(m, s, pre_j);
      }
      if (MULTI) {
        // MultiStep.propose_step, mcusher.py:284-304.  Proposal j sees proposals < j applied: they only ever
        // touch other sites (a colliding proposal is dropped), so the occupancy bytes can be read as they are;
        // the species bit-planes that serve a swap's partner choice are toggled for the chained swaps and
        // restored before the evaluation.
        r1blk = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 1u, wid, k0, k1);
        const double ul = u01(r1blk.x);
        int li = 0;
        while (li < a.ms_num - 1 && !(a.ms_cum[li] > ul)) ++li;
        const int len = a.ms_len[li];
        for (int j = 0; j < len; ++j) {
          const U4 rj = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)(2 + j), wid, k0, k1);
          const int sl = choose_sublattice(m, rj.x);
          const int n_act = m.sl_off[sl + 1] - m.sl_off[sl];
          const int pos1 = (int)mulhi32(rj.y, (uint32_t)n_act);
          const int site1 = site_of_pos(m, sl, pos1);
          const int s1 = occ[site1];
          bool clash = false;
#pragma unroll
          for (int f = 0; f < MF; ++f) clash = clash || (f < st.n && st.site[f] == site1);
          if (a.ms_usher == LMC_USHER_FLIP) {
            const int nc = m.sl_ncodes[sl];
            int ci = (int)mulhi32(rj.z, (uint32_t)(nc - 1));
            int cp = nc;
            for (int c = 0; c < nc; ++c) if (m.sl_codes[sl][c] == s1) { cp = c; break; }
            if (ci >= cp) ++ci;
            if (!clash) push_flip(st, site1, s1, m.sl_codes[sl][ci], sl, pos1);
          } else {
            const int ndiff = n_act - cnt[sl * LMC_MAX_CODES + s1];
            if (!clash && ndiff > 0) {   // uniform over the group (a clash makes the proposal void whatever it picks)
              const int k = (int)mulhi32(rj.z, (uint32_t)ndiff);
              const int pos2 = select_pos<G>(m, planes, sl, s1, k, true, g, gmask);
              const int site2 = site_of_pos(m, sl, pos2);
              const int s2 = occ[site2];
#pragma unroll
              for (int f = 0; f < MF; ++f) clash = clash || (f < st.n && st.site[f] == site2);
              if (!clash) {
                push_flip(st, site1, s1, s2, sl, pos1);
                push_flip(st, site2, s2, s1, sl, pos2);
                if (j + 1 < len) {   // later proposals pick partners in the swapped configuration
                  group_sync<G>(gmask);
                  if (g == 0) {
                    const int nw = m.sl_nwords[sl];
                    uint32_t* pl = planes + m.sl_plane_off[sl];
                    pl[s1 * nw + (pos1 >> 5)] ^= 1u << (pos1 & 31); pl[s2 * nw + (pos1 >> 5)] ^= 1u << (pos1 & 31);
                    pl[s2 * nw + (pos2 >> 5)] ^= 1u << (pos2 & 31); pl[s1 * nw + (pos2 >> 5)] ^= 1u << (pos2 & 31);
                  }
                  group_sync<G>(gmask);
                  ms_toggled += 2;
                }
              }
            }
          }
        }
        if (ms_toggled) {   // restore the planes of the current occupancy
          group_sync<G>(gmask);
          if (g == 0) {
#pragma unroll
            for (int f = 0; f < MF; ++f)
              if (f < ms_toggled) {
                const int nw = m.sl_nwords[st.sl[f]];
                uint32_t* pl = planes + m.sl_plane_off[st.sl[f]] + (st.pos[f] >> 5);
                pl[st.oldc[f] * nw] ^= 1u << (st.pos[f] & 31);
                pl[st.newc[f] * nw] ^= 1u << (st.pos[f] & 31);
              }
          }
          group_sync<G>(gmask);
        }
      }
      if (USHER == LMC_USHER_FLIP || (COMP && usher == LMC_USHER_FLIP)) {
        // Flip.propose_step, mcusher.py:154-170
        const int sl = pre_sl, j = pre_j, site = pre_site;
        const int cur = occ[site];
        const int nc = m.sl_ncodes[sl];
        int ci = (int)mulhi32(q2, (uint32_t)(nc - 1));
        const int pos = m.sl_code_pos[sl][cur];   // index of the current code in the encoding (nc if it is not in it)
        if (ci >= pos) ++ci;
        st.n = 1; st.site[0] = site; st.oldc[0] = cur; st.newc[0] = m.sl_codes[sl][ci]; st.sl[0] = sl; st.pos[0] = j;
      } else if (USHER == LMC_USHER_SWAP || usher == LMC_USHER_SWAP) {
        // Swap.propose_step, mcusher.py:176-200
        int sl = pre_sl, j = pre_j, site1 = pre_site;
        if (USHER == LMC_USHER_TABLEFLIP) {  // fallback swap of the table-flip usher: words 4,5,6
          sl = choose_sublattice(m, q0);
          j = (int)mulhi32(q1, (uint32_t)(m.sl_off[sl + 1] - m.sl_off[sl]));
          site1 = site_of_pos(m, sl, j);
        }
        const int n_act = m.sl_off[sl + 1] - m.sl_off[sl];
        const int s1 = occ[site1];
        const int ndiff = n_act - cnt[sl * LMC_MAX_CODES + s1];
        if (ndiff > 0) {
          const int k = (int)mulhi32(q2, (uint32_t)ndiff);
          const int p2 = select_pos<G>(m, planes, sl, s1, k, true, g, gmask);
          const int site2 = site_of_pos(m, sl, p2);
          const int s2 = occ[site2];
          st.n = 2;
          st.site[0] = site1; st.oldc[0] = s1; st.newc[0] = s2; st.sl[0] = sl; st.pos[0] = j;
          st.site[I1] = site2; st.oldc[I1] = s2; st.newc[I1] = s1; st.sl[I1] = sl; st.pos[I1] = p2;
        }
      } else if (USHER == LMC_USHER_TABLEFLIP) {
        // table flip: sequential picks, one random word each (words 4.. of the step)
        const int sgn = (tf_idx & 1) ? -1 : 1;
        const int* urow = m.tf_table[tf_idx >> 1];
        int wi = 0;  // pick counter
        U4 rb = r1blk;
        int cur_blk = 1;
        auto next_word = [&]() -> uint32_t {
          const int b = 1 + (wi >> 2);
          if (b != cur_blk) { rb = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)b, wid, k0, k1); cur_blk = b; }
          const int l = wi & 3;
          ++wi;
          return l == 0 ? rb.x : l == 1 ? rb.y : l == 2 ? rb.z : rb.w;
        };
        int d0 = 0;
        // dims of one sublattice are consecutive (occu_utils.py:20-25); walk sublattice by sublattice
        while (d0 < m.tfD) {
          const int sl = m.tf_dim_sl[d0];
          int d1 = d0 + 1;
          while (d1 < m.tfD && m.tf_dim_sl[d1] == sl) ++d1;
          if (sl >= 0) {
            // picked sites / positions / ranks: four 16-bit fields of one register pair each, indexed by shifts (arrays
            // indexed at run time live in local memory: every pick then waited on L1 / L2)
            unsigned long long pool = 0ull, ppos = 0ull;
            int npool = 0;
            for (int d = d0; d < d1; ++d) {
              const int ud = sgn * urow[d];
              if (ud >= 0) continue;
              unsigned long long ranks = 0ull;   // ascending
              int nr = 0;
              const int ndd = cnt[sl * LMC_MAX_CODES + m.tf_dim_code[d]];
              for (int p = 0; p < -ud; ++p) {
                int idx = (int)mulhi32(next_word(), (uint32_t)(ndd - p));
                // index among the remaining sites -> rank in the original (ascending-site) list
                int at = 0;
                for (int q = 0; q < nr; ++q)
                  if (idx >= (int)((ranks >> (16 * q)) & 0xffffull)) { ++idx; at = q + 1; }
                const unsigned long long low = (1ull << (16 * at)) - 1ull;
                ranks = (ranks & low) | ((unsigned long long)idx << (16 * at)) | ((ranks & ~low) << 16);
                ++nr;
                const int pp = select_pos<G>(m, planes, sl, m.tf_dim_code[d], idx, false, g, gmask);
                if (npool < LMC_MAX_FLIPS) {
                  pool |= (unsigned long long)site_of_pos(m, sl, pp) << (16 * npool);
                  ppos |= (unsigned long long)pp << (16 * npool);
                  ++npool;
                }
              }
            }
            for (int d = d0; d < d1; ++d) {
              const int ud = sgn * urow[d];
              if (ud <= 0) continue;
              for (int p = 0; p < ud; ++p) {
                const int idx = (int)mulhi32(next_word(), (uint32_t)npool);
                const int site = (int)((pool >> (16 * idx)) & 0xffffull), pp = (int)((ppos >> (16 * idx)) & 0xffffull);
                const unsigned long long low = (1ull << (16 * idx)) - 1ull;
                pool = (pool & low) | ((pool >> 16) & ~low);
                ppos = (ppos & low) | ((ppos >> 16) & ~low);
                --npool;
                push_flip(st, site, occ[site], m.tf_dim_code[d], sl, pp);
              }
            }
          }
          d0 = d1;
        }
        // compute_log_priori_factor, mcusher.py:656-711 (tabulated per direction above)
        st.log_priori = tfc[4 * LMC_MAX_TABLE_FLIPS + tf_idx];
      }

      // ------------------------------ evaluate ------------------------------------------
      // lanes of a group are not assumed to run in lockstep: all reads of the proposal phase are
      // complete before lane 0 starts writing flips into the occupancy
      group_sync<G>(gmask);
      double acc = 0.0, acc_ew = 0.0, dmu = 0.0;
      constexpr bool MU_POSSIBLE = USHER != LMC_USHER_SWAP;  // a swap leaves the chemical work unchanged
      // the cluster records depend on the sites only: fetch the first two flips' records up front so
      // that their L2 latency overlaps with the rest of the proposal
      RecChunk pre0, pre1;
      bool deferred1 = false;
      // orbit segments for phase B: prefetched only by the variants that are not register-capped
      // (high-acceptance workloads: Wang-Landau, Ewald, table flips); the 72-register swap/flip
      // kernel fetches them on accept
      constexpr bool SEGPRE = EWALD || WLMODE || USHER == LMC_USHER_TABLEFLIP || G < 32;
      int4 seg0 = make_int4(0, 0, 0, -1), seg1 = make_int4(0, 0, 0, -1);
      if (PREF && have_nxt) {   // fetched during the previous step
        cp_async_wait_all();
        const uint2* br = nxt_rec + nxt_buf * m.Rstride;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = g + u * G;
          pre0.r[u] = r < m.Rstride ? br[r] : make_uint2(0u, (uint32_t)m.nCls << 16);
        }
        seg0 = g < m.Sstride ? nxt_seg[nxt_buf * m.Sstride + g] : make_int4(0, 0, 0, -1);
      } else if (st.n > 0) { pre0 = load_records<G>(m, st.site[0], g); if (SEGPRE) seg0 = load_segment<G>(m, st.site[0], g); }
      if (st.n > 1) { pre1 = load_records<G>(m, st.site[I1], g); if (SEGPRE) seg1 = load_segment<G>(m, st.site[I1], g); }
      if (PREF) {
        // the site of the NEXT step is already in the ring (state independent): fetch its records and segment
        // entries into the other buffer now; a whole step hides their L2 latency
        have_nxt = bphase != 0;
        if (have_nxt) {
          nxt_buf ^= 1;
          const int ns = (int)ring[bphase].y;
          const uint2* rp = m.site_rec + (size_t)ns * m.Rstride;
          uint2* br = nxt_rec + nxt_buf * m.Rstride;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = g + u * G;
            if (r < m.Rstride) cp_async_8(br + r, rp + r);
          }
          if (g < m.Sstride) cp_async_16(nxt_seg + nxt_buf * m.Sstride + g, m.site_seg + (size_t)ns * m.Sstride + g);
          cp_async_commit();
        }
      }
      // chemical work (ensemble.py:369-373): table lookups of the proposal only, issued ahead of the evaluation
      if (MU_POSSIBLE && m.muW) {
#pragma unroll
        for (int f = 0; f < MF; ++f)
          if (f < st.n)
            dmu += mu_of(m, st.site[f], st.newc[f], st.sl[f]) - mu_of(m, st.site[f], st.oldc[f], st.sl[f]);
      }
      // Ewald part first: it only touches the per-walker Ewald cache, never the occupancy.  Flip f is
      // evaluated with flips < f applied to the cache (sequential semantics, ewald.py:168-181); the
      // last flip is applied on accept only.  One loop, not unrolled: one copy of the row-gather code.
      double dq[MF];
#pragma unroll
      for (int f = 0; f < MF; ++f) dq[f] = 0.0;
      if (EWFIELD) {
        double e = 0.0;
#pragma unroll
        for (int f = 0; f < MF; ++f)
          if (f < st.n) {
            const double2 qn = ewald_qd(m, st.site[f], st.newc[f], st.sl[f]), qo = ewald_qd(m, st.site[f], st.oldc[f], st.sl[f]);
            dq[f] = qn.x - qo.x;
            double phi = fld[st.site[f]];
#pragma unroll
            for (int h = 0; h < f; ++h) phi += dq[h] * __ldg(m.ewK + (size_t)st.site[h] * m.N + st.site[f]);
            e += 2.0 * dq[f] * phi + (qn.y - qo.y);
          }
        acc_ew = g == 0 ? e : 0.0;   // every lane holds the same value; the group sum below counts it once
      } else if (EWGATHER) {
#pragma unroll 1
        for (int f = 0; f < st.n; ++f) {
          const int sf = pick<MF>(st.site, f), of = pick<MF>(st.oldc, f), nf = pick<MF>(st.newc, f);
          acc_ew += flip_ewald<G>(m, t, eidx, sf, of, nf, g);
          if (f + 1 < st.n) {
            group_sync<G>(gmask);
            if (g == 0) ewald_cache_set(m, eidx, sf, nf);
            group_sync<G>(gmask);
          }
        }
      }
      constexpr bool SEQ_ONLY = USHER == LMC_USHER_TABLEFLIP;   // compact code for the big variant
      if (!SEQ_ONLY && st.n >= 2 && !a.seq_flips) {
        if (g == 0) occ[st.site[0]] = (uint8_t)st.newc[0];
        group_sync<G>(gmask);
        acc = flip_energy_pair<G, KONE>(m, t, occ, st.site[0], st.oldc[0], st.newc[0], st.site[I1], st.oldc[I1],
                                        st.newc[I1], stash0, stash0 + stash_stride, g, pre0, pre1);
        // flip 0's records read site 1 (old value): lanes are not guaranteed to run in lockstep, so
        // site 1 may only be written once every lane is done -- after a sync (more flips follow) or
        // after the group reduction below (two-flip step: written on accept only)
        if (st.n > 2) {
          group_sync<G>(gmask);
          if (g == 0) occ[st.site[I1]] = (uint8_t)st.newc[I1];
          group_sync<G>(gmask);
        } else {
          deferred1 = true;
        }
#pragma unroll
        for (int f = 2; f < MF; ++f) {
          if (f < st.n) {
            pre0 = load_records<G>(m, st.site[f], g);
            acc += flip_energy<G, KONE>(m, t, occ, st.site[f], st.oldc[f], st.newc[f], stash0 + f * stash_stride, g, pre0);
            if (g == 0) occ[st.site[f]] = (uint8_t)st.newc[f];
            if (f + 1 < st.n) group_sync<G>(gmask);
          }
        }
      } else {
        // strictly sequential evaluation (single flips, the table-flip variant, debug switch)
#pragma unroll 1
        for (int f = 0; f < st.n; ++f) {
          const int sf = pick<MF>(st.site, f), of = pick<MF>(st.oldc, f), nf = pick<MF>(st.newc, f);
          if (f > 0) pre0 = (f == 1 && !SEQ_ONLY) ? pre1 : load_records<G>(m, sf, g);
          acc += flip_energy<G, KONE>(m, t, occ, sf, of, nf, stash0 + f * stash_stride, g, pre0);
          if (g == 0) occ[sf] = (uint8_t)nf;
          if (f + 1 < st.n) group_sync<G>(gmask);
        }
      }
      double dH = group_sum<G>(acc, gmask);
      double dEw = 0.0;
      if (EWALD) { dEw = group_sum<G>(acc_ew, gmask); dH += nat_ew * dEw; }
      if (MU_POSSIBLE && m.muW) dH += nat_mu * dmu;
      double dist_L = 0.0;
      if (DIST) {
        // DistanceProcessor.compute_feature_vector_change (distance.py:156-180): distances of the vector with
        // the flips applied minus the current ones; the vector is per supercell (not times the size)
        double* dd = dvec + m.F;
        double* dn = dvec + 2 * m.F;
        for (int f = g; f < m.F; f += G) dd[f] = 0.0;
        group_sync<G>(gmask);   // stash of this proposal complete, delta zeroed
#pragma unroll
        for (int f = 0; f < MF; ++f)
          if (f < st.n) {
            flip_features<G, KONE>(m, t, st.site[f], stash0 + f * stash_stride, dd, g, load_segment<G>(m, st.site[f], g));
            group_sync<G>(gmask);
          }
        double part = 0.0;
        const double inv_size = 1.0 / (double)m.size;
        for (int f = g; f < m.F; f += G)
          if (f > 0) {
            const double v = fabs(dvec[f] + dd[f] * inv_size - __ldg(a.dist_target + f));
            dn[f] = v;
            part += t.nat[f] * (v - feat[f]);
          }
        group_sync<G>(gmask);
        if (t.nat[0] != 0.0) {   // exact_match_max_diameter, distance.py:309-331, 452-472
          for (int q = 0; q < a.dist_ngrp; ++q) {
            bool ok = true;
            for (int i = __ldg(a.dist_grp_off + q); i < __ldg(a.dist_grp_off + q + 1); ++i)
              ok = ok && dn[__ldg(a.dist_grp_idx + i)] <= a.dist_tol;
            if (!ok) break;
            dist_L = __ldg(a.dist_grp_diam + q);
          }
        }
        dH = group_sum<G>(part, gmask) + t.nat[0] * (dist_L - feat[0]);
      }

      // ------------------------------ accept --------------------------------------------
      double new_fb = cur_fb, s_new = s_cur;
      double dbias = 0.0, dbc = 0.0;   // dbc: table difference of the last row (the only one for a table-sum bias)
      if (!wl_mode) {
        // MetropolisAcceptMixin._accept_step, kernel/metropolis.py:31-49
        // (multicell hop: judged by the enthalpy of the target shape's new state minus the current shape's)
        double exponent = __dadd_rn(__dmul_rn(-beta, a.acc_off ? dH + acc_off : dH), st.log_priori);
        if (a.bias_mode) {
          // MCBias.compute_bias_change (bias.py:79-95, 193-214): table differences of the changed sites.
          // Square biases: bias(after) - bias(before) with bias = -penalty * sum_r c_r^2 (bias.py:276-287, 342-353);
          // the c_r are integer valued (charges, hyperplane residuals A n - b), so the sums are exact
          double q0 = 0.0, q1 = 0.0;
          for (int r = 0; r < a.bias_rows; ++r) {
            double d = 0.0;
#pragma unroll
            for (int f = 0; f < MF; ++f)
              if (f < st.n)
                d += __ldg(a.bias_tab + (st.site[f] * a.bias_w + st.newc[f]) * a.bias_rows + r) -
                     __ldg(a.bias_tab + (st.site[f] * a.bias_w + st.oldc[f]) * a.bias_rows + r);
            const double c0 = bstate[1 + r], c1 = c0 + d;
            q0 += __dmul_rn(c0, c0); q1 += __dmul_rn(c1, c1);
            dbc = d;
          }
          if (a.bias_mode == LMC_BIAS_TABLE_SUM) dbias = dbc;
          else dbias = __dsub_rn(-__dmul_rn(a.bias_pen, q1), -__dmul_rn(a.bias_pen, q0));
          exponent = __dadd_rn(exponent, dbias);   // metropolis.py:43-44
        }
        {
          const int af = accept_fast(exponent, lf);
          accepted = af >= 0 ? (af != 0)
                             : exponent > log(u01(philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 0u, wid, k0, k1).w));
        }
      } else {
        // WangLandau._accept_step, kernel/wanglandau.py:186-202.  The current bin and its entropy live
        // in registers (the post-step bin of step t is the pre-step bin of step t+1), so only the
        // entropy of the NEW bin is loaded.
        const double e_new = enth + dH;
        if (e_new < a.wl.min_enthalpy || e_new >= a.wl.max_enthalpy) {
          accepted = false;
        } else {
          new_fb = exact_floordiv_inv(e_new - a.wl.min_enthalpy, a.wl.bin_size, wl_inv_bin);
          s_new = new_fb == cur_fb ? s_cur
                                   : ((new_fb >= 0.0 && new_fb < (double)nb) ? wl_load(wlSs, wlS, (int)new_fb, wl_sm) : 0.0);
          const double exponent = (s_cur - s_new) + st.log_priori;
          const int af = accept_fast(exponent, lf);
          accepted = af >= 0 ? (af != 0)
                             : exponent > log(u01(philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), 0u, wid, k0, k1).w));
        }
      }

      // ------------------------------ update --------------------------------------------
      group_sync<G>(gmask);   // every lane has finished reading the occupancy / caches of this step
      if (accepted) {
        // MCKernel._do_accept_step (kernel/base.py:327-343) + trace accumulation (sampler.py:204-207)
        if (DIST) {
          const double inv_size = 1.0 / (double)m.size;
          for (int f = g; f < m.F; f += G) {
            dvec[f] += dvec[m.F + f] * inv_size;
            feat[f] = f > 0 ? dvec[2 * m.F + f] : dist_L;
          }
        } else if (USHER == LMC_USHER_TABLEFLIP) {
          // one copy of the fold (this variant is bound by instruction fetch, see above)
#pragma unroll 1
          for (int f = 0; f < st.n; ++f) {
            if (f > 0) group_sync<G>(gmask);
            const int sf = pick<MF>(st.site, f);
            flip_features<G, KONE>(m, t, sf, stash0 + f * stash_stride, feat, g,
                                   f == 0 ? seg0 : (f == 1 ? seg1 : load_segment<G>(m, sf, g)));
          }
        } else {
#pragma unroll
          for (int f = 0; f < MF; ++f)
            if (f < st.n) {
              // (lanes of different flips may own the same feature: one flip after the other)
              if (f > 0) group_sync<G>(gmask);
              flip_features<G, KONE>(m, t, st.site[f], stash0 + f * stash_stride, feat, g,
                                     (SEGPRE && f == 0) ? seg0 : ((SEGPRE && f == 1) ? seg1 : load_segment<G>(m, st.site[f], g)));
            }
        }
        if (g == 0) {
          if (deferred1) occ[st.site[I1]] = (uint8_t)st.newc[I1];
          if (EWGATHER && st.n > 0)   // the last flip enters the Ewald cache on accept only
            ewald_cache_set(m, eidx, pick<MF>(st.site, st.n - 1), pick<MF>(st.newc, st.n - 1));
          if (EWALD) feat[m.ewF] += dEw;
          if (MU_POSSIBLE && m.muW) feat[m.muF] += dmu;
#pragma unroll
          for (int f = 0; f < MF; ++f)
            if (USHER != LMC_USHER_FLIP && f < st.n) {   // (a pure flip usher never reads the counts / bit-planes)
              cnt[st.sl[f] * LMC_MAX_CODES + st.oldc[f]]--;
              cnt[st.sl[f] * LMC_MAX_CODES + st.newc[f]]++;
              const int nw = m.sl_nwords[st.sl[f]];
              uint32_t* pl = planes + m.sl_plane_off[st.sl[f]] + (st.pos[f] >> 5);
              const uint32_t bit = 1u << (st.pos[f] & 31);
              pl[st.oldc[f] * nw] ^= bit;
              pl[st.newc[f] * nw] ^= bit;
            }
        }
        if (EWFIELD) {   // accepted: shift the potential cache by the changed charges (rows of K: four loads per row in flight)
#pragma unroll 4
          for (int k = g; k < m.N; k += G) {
            double v = fld[k];
#pragma unroll
            for (int f = 0; f < MF; ++f)
              if (f < st.n) v += dq[f] * __ldg(m.ewK + (size_t)st.site[f] * m.N + k);
            fld[k] = v;
          }
        }
        enth += dH;
        if (USHER == LMC_USHER_TABLEFLIP && tf_idx >= 0) tfc_valid = false;   // the species counts changed
        if (!WLMODE && a.bias_mode && g == 0) {   // read again after the step's final sync
          bstate[0] += dbias;
          for (int r = 0; r < a.bias_rows; ++r) {
            double d = 0.0;
#pragma unroll
            for (int f = 0; f < MF; ++f)
              if (f < st.n)
                d += __ldg(a.bias_tab + (st.site[f] * a.bias_w + st.newc[f]) * a.bias_rows + r) -
                     __ldg(a.bias_tab + (st.site[f] * a.bias_w + st.oldc[f]) * a.bias_rows + r);
            bstate[1 + r] += d;
          }
        }
        if (wl_mode) { cur_fb = new_fb; s_cur = s_new; }
        ++nacc;
      } else if (st.n > 0) {
        if (g == 0) {
#pragma unroll
          for (int f = MF - 1; f >= 0; --f)
            if (f < st.n) {
              if (!(f == 1 && deferred1)) occ[st.site[f]] = (uint8_t)st.oldc[f];
              if (EWGATHER && f + 1 < st.n) ewald_cache_set(m, eidx, st.site[f], st.oldc[f]);
            }
        }
      }
      group_sync<G>(gmask);

      if (wl_mode) {
        // WangLandau._do_post_step, kernel/wanglandau.py:222-266
        if (cur_fb >= 0.0 && cur_fb < (double)nb) {
          const int bin = (int)cur_fb;
          ++wl_cnt;
          if (++upd_rem == a.wl.update_period) upd_rem = 0;
          if (++chk_rem == a.wl.check_period) chk_rem = 0;
          const bool upd = upd_rem == 0;
          if (wl_sum) {
            // update_period == 1: the running mean (x_n + (n-1) M)/n is sum/n -- accumulate the sum with
            // fire-and-forget reductions, the host divides by `occurrences`
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
            if (upd && g == 0) __stcg(wlO + bin, total + 1);
          }
          if (upd) {
            s_cur += wl_m;
            if (g == 0) {
              if (wl_sm) {
                wlSs[bin] = s_cur;
                wlHs[bin] += 1;
              } else {
                __stcg(wlS + bin, s_cur);
                red_add_u64(wlH + bin, 1ull);
              }
            }
          }
        }
        wl_m_traced = wl_m;   // trace.mod_factor is copied before the flatness check (wanglandau.py:251)
        if (chk_rem == 0) {
          __threadfence_block();
          group_sync<G>(gmask);   // lane 0's entropy/histogram updates are visible to the group
          int nvis = 0;
          double hsum = 0.0, hmin = 1e300;
          for (int b = g; b < nb; b += G)
            if (wl_load(wlSs, wlS, b, wl_sm) > 0.0) {
              const double h = (double)wl_load(wlHs, wlH, b, wl_sm);
              ++nvis; hsum += h; hmin = fmin(hmin, h);
            }
          nvis = group_sum_i<G>(nvis, gmask);
          hsum = group_sum<G>(hsum, gmask);
          hmin = group_min<G>(hmin, gmask);
          if (nvis >= 2 && hmin > a.wl.flatness * (hsum / (double)nvis)) {
            for (int b = g; b < nb; b += G) { if (wl_sm) wlHs[b] = 0ll; else __stcg(wlH + b, 0ll); }
            wl_m = wl_next_mod_factor(a.wl, wl_m);
            group_sync<G>(gmask);
          }
        }
      }
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
      if (!WLMODE && a.bias_mode && a.tr_bias) a.tr_bias[sw] = bstate[0];
      if (WLMODE && a.wl.trace_mod_factor_dev) a.wl.trace_mod_factor_dev[sw] = wl_m_traced;
    }
    if (WLMODE && (a.wl.trace_entropy_dev || a.wl.trace_histogram_dev || a.wl.trace_occurrences_dev ||
                   a.wl.trace_mean_features_dev)) {
      // the walker's Wang-Landau arrays as they stand after the sampled step (wanglandau.py:247-250: the trace
      // holds the arrays themselves, so a histogram reset of that very step shows)
      __threadfence_block();
      group_sync<G>(gmask);
      for (int b = g; b < nb; b += G) {
        const long long oc = __ldcg(wlO + b);
        if (a.wl.trace_entropy_dev) __stcs(a.wl.trace_entropy_dev + sw * nb + b, wl_load(wlSs, wlS, b, wl_sm));
        if (a.wl.trace_histogram_dev) __stcs(reinterpret_cast<long long*>(a.wl.trace_histogram_dev) + sw * nb + b, wl_load(wlHs, wlH, b, wl_sm));
        if (a.wl.trace_occurrences_dev) __stcs(reinterpret_cast<long long*>(a.wl.trace_occurrences_dev) + sw * nb + b, oc);
      }
      if (a.wl.trace_mean_features_dev) {
        for (int i = g; i < nb * m.F; i += G) {
          double v = __ldcg(wlM + i);
          if (wl_sum) {
            const long long oc = __ldcg(wlO + i / m.F);
            if (oc != 1) v = oc > 0 ? v / (double)oc : 0.0;   // (most bins of a short run were never or once visited)
          }
          __stcs(a.wl.trace_mean_features_dev + (sw * nb) * m.F + i, v);
        }
      }
    }
    group_sync<G>(gmask);
  }

  // ------------------------------ final state ---------------------------------------------
  for (int i = g; i < m.N; i += G) a.occ[(size_t)w * m.Npad + i] = (int8_t)occ[i];
  for (int f = g; f < m.F; f += G) a.features[(size_t)w * m.F + f] = feat[f];
  if (DIST)
    for (int f = g; f < m.F; f += G) a.dist_vec[(size_t)w * m.F + f] = dvec[f];
  if (wl_sm) {
    group_sync<G>(gmask);
    for (int b = g; b < nb; b += G) { wlS[b] = wlSs[b]; wlH[b] = wlHs[b]; }
  }
  if (g == 0) {
    a.enthalpy[w] = enth;
    if (!WLMODE && a.bias_mode) {
      a.bias[w] = bstate[0];
      for (int r = 0; r < a.bias_rows; ++r) a.bias_sum[(size_t)w * a.bias_rows + r] = bstate[1 + r];
    }
    if (wl_mode) {
      a.wl.mod_factor_dev[w] = wl_m;
      a.wl.steps_counter_dev[w] = wl_cnt;
    }
    if (a.stats && !wl_mode) {   // acceptance feedback for the kernel selection of the next launches
      atomicAdd(a.stats, (unsigned long long)nacc_total);
      atomicAdd(a.stats + 1, (unsigned long long)(a.S * (long long)a.thin));
    }
  }
}

// ------------------------------------------------------------------------------------------
// batched feature change for given flips (parity / Processor API): one group per walker
// ------------------------------------------------------------------------------------------
template <int G, bool KONE>
__global__ void lmc_delta_kernel(const DevModel m, const int8_t* __restrict__ occ_g, int W, const int* __restrict__ sites,
                                 const int* __restrict__ codes, int nflips, double* __restrict__ out, int wpb,
                                 int walker_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint64_t bar;
  const int g = threadIdx.x % G, wl_ = threadIdx.x / G;
  const int w = blockIdx.x * wpb + wl_;
  const int nw_blk = min(wpb, W - blockIdx.x * wpb);
  const uint32_t gmask = group_mask<G>();
  unsigned char* wbase = smem + ((m.off_dtab + 15) & ~15);
  uint8_t* occ = wbase + (size_t)wl_ * m.Npad;
  unsigned char* priv = wbase + (size_t)wpb * m.Npad + (size_t)wl_ * walker_smem;
  double* feat = reinterpret_cast<double*>(priv);
  unsigned char* stash = priv + ((m.F * 8 + 15) & ~15);
  stage_tables(m, smem, &bar, wbase, occ_g + (size_t)blockIdx.x * wpb * m.Npad, (uint32_t)(nw_blk * m.Npad),
               (uint32_t)m.off_dtab);
  const SmemTables t = smem_tables(m, smem);
  if (wl_ >= wpb || w >= W) return;
  for (int f = g; f < m.F; f += G) feat[f] = 0.0;
  void* eidx = stash + ((m.Rstride * 8 + 15) & ~15);
  if (m.E)
    for (int i = g; i < m.N; i += G) ewald_cache_set(m, eidx, i, occ[i]);
  group_sync<G>(gmask);
  double dmu = 0.0, dew = 0.0;
  for (int f = 0; f < nflips; ++f) {
    const int site = sites[(size_t)w * nflips + f], newc = codes[(size_t)w * nflips + f];
    const int oldc = occ[site];
    // chemical work against the PRE-step occupancy (ensemble.py:369-373)
    if (m.muW) dmu += m.mu[site * m.muW + newc] - m.mu[site * m.muW + (int)occ_g[(size_t)w * m.Npad + site]];
    (void)flip_energy<G, KONE>(m, t, occ, site, oldc, newc, stash, g, load_records<G>(m, site, g));
    if (m.E) dew += flip_ewald<G>(m, t, eidx, site, oldc, newc, g);
    group_sync<G>(gmask);
    flip_features<G, KONE>(m, t, site, stash, feat, g, load_segment<G>(m, site, g));
    if (g == 0) { occ[site] = (uint8_t)newc; if (m.E) ewald_cache_set(m, eidx, site, newc); }
    group_sync<G>(gmask);
  }
  if (m.E) dew = group_sum<G>(dew, gmask);
  if (g == 0) {
    if (m.E) feat[m.ewF] = dew;
    if (m.muW) feat[m.muF] = dmu;
  }
  group_sync<G>(gmask);
  for (int f = g; f < m.F; f += G) out[(size_t)w * m.F + f] = feat[f];
}

#ifdef LMC_API_TU  // non-template kernels live in the API translation unit only
// ------------------------------------------------------------------------------------------
// full feature vector: one block per walker (correlations_from_occupancy /
// interactions_from_occupancy, evaluator.pyx:121-209; ewald.py:128-145; ensemble.py:344-349)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wp] = v;
  __syncthreads();
  double tot = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  return tot;
}

__global__ void lmc_full_kernel(const DevModel m, const int8_t* __restrict__ occ_g, double* __restrict__ features,
                                double* __restrict__ enthalpy, const OrbDev* __restrict__ orbs,
                                const double* __restrict__ nat, const double* __restrict__ field) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red[32];
  uint8_t* occ = smem;
  int* eidx = reinterpret_cast<int*>(smem + m.Npad);
  const int w = blockIdx.x;
  for (int i = threadIdx.x; i < m.N; i += blockDim.x) occ[i] = (uint8_t)occ_g[(size_t)w * m.Npad + i];
  __syncthreads();
  double* out = features + (size_t)w * m.F;
  double enth = 0.0;
  if (threadIdx.x == 0) out[0] = m.feature0;
  enth += nat[0] * m.feature0;
  for (int n = 0; n < m.nOrb; ++n) {
    const OrbDev o = orbs[n];
    for (int k = 0; k < o.K; ++k) {
      const double* tk = m.ftab + o.ftab_off + k * o.T;
      double p = 0.0;
      for (int r = threadIdx.x; r < o.row_cnt; r += blockDim.x) {
        const uint2 row = __ldg(m.full_rows + o.row_off + r);
        const int idx = o.stride[0] * occ[row.x & 0xffffu] + o.stride[1] * occ[row.x >> 16] +
                        o.stride[2] * occ[row.y & 0xffffu] + o.stride[3] * occ[row.y >> 16];
        p += __ldg(tk + idx);
      }
      p = block_sum(p, red);
      const double v = p / (double)o.row_cnt * (double)m.size;
      if (threadIdx.x == 0) out[o.fidx + k] = v;
      enth += nat[o.fidx + k] * v;
    }
  }
  if (m.E) {
    for (int i = threadIdx.x; i < m.N; i += blockDim.x) eidx[i] = m.ewInds[i * m.ewW + occ[i]];
    __syncthreads();
    double p = 0.0;
    const long long np = field ? 0ll : (long long)m.N * m.N;
    if (field) {
      // factorised matrix with the walker's potential fld[k] = sum_j q_j K[k][j] at hand (lmc_ewald_field_kernel):
      // sum over occupied pairs = sum_k q_k fld[k] + sum_k M[e_k, e_k]   -- O(N) instead of O(N^2)
      for (int k = threadIdx.x; k < m.N; k += blockDim.x) {
        const double2 qd = __ldg(m.ewQD + k * m.ewW + occ[k]);
        p += qd.x * field[(size_t)w * m.N + k] + qd.y;
      }
    }
    for (long long q = threadIdx.x; q < np; q += blockDim.x) {
      const int i = (int)(q / m.N), j = (int)(q % m.N);
      const int a = eidx[i], b = eidx[j];
      if (a >= 0 && b >= 0) {
        if (m.ewK) p += i == j ? __ldg(m.ewD + a) : __ldg(m.ewQ + a) * __ldg(m.ewQ + b) * __ldg(m.ewK + q);
        else p += __ldg(m.ewMt + (size_t)a * m.E + b);
      }
    }
    p = block_sum(p, red);
    if (threadIdx.x == 0) out[m.ewF] = p;
    enth += nat[m.ewF] * p;
  }
  if (m.muW) {
    double p = 0.0;
    for (int i = threadIdx.x; i < m.N; i += blockDim.x) p += m.mu[i * m.muW + occ[i]];
    p = block_sum(p, red);
    if (threadIdx.x == 0) out[m.muF] = p;
    enth += nat[m.muF] * p;
  }
  if (enthalpy && threadIdx.x == 0) enthalpy[w] = enth;
}

// Ewald potential cache of every walker: field[w][s] = sum_k q_k(w) K[s][k].  Four walkers per block share
// each row of K (read once from L2, coalesced); one warp per row.
constexpr int FIELD_WPB = 4;
__global__ void lmc_ewald_field_kernel(const DevModel m, const int8_t* __restrict__ occ_g, int W, double* __restrict__ field) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* qs = reinterpret_cast<double*>(smem);   // [FIELD_WPB][N] charges of the walkers' current species
  const int w0 = blockIdx.x * FIELD_WPB;
  for (int i = threadIdx.x; i < FIELD_WPB * m.N; i += blockDim.x) {
    const int ww = i / m.N, k = i - ww * m.N;
    double q = 0.0;
    if (w0 + ww < W) {
      const int e = m.ewInds[k * m.ewW + occ_g[(size_t)(w0 + ww) * m.Npad + k]];
      if (e >= 0) q = m.ewQ[e];
    }
    qs[i] = q;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nwp = blockDim.x >> 5;
  for (int s = wp; s < m.N; s += nwp) {
    const double* krow = m.ewK + (size_t)s * m.N;
    double acc[FIELD_WPB];
#pragma unroll
    for (int ww = 0; ww < FIELD_WPB; ++ww) acc[ww] = 0.0;
    for (int k = lane; k < m.N; k += 32) {
      const double kv = __ldg(krow + k);
#pragma unroll
      for (int ww = 0; ww < FIELD_WPB; ++ww) acc[ww] += qs[ww * m.N + k] * kv;
    }
#pragma unroll
    for (int ww = 0; ww < FIELD_WPB; ++ww) {
      double v = acc[ww];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && w0 + ww < W) field[(size_t)(w0 + ww) * m.N + s] = v;
    }
  }
}

// The same as a tiled matrix product field[W][N] = Q[W][N] x K[N][N] (Q = charges of the walkers' species; K is
// symmetric): a block owns 128 walkers x 64 sites, a thread 8 x 4 of them; k runs in chunks of 16 staged through
// shared memory (charges looked up on the way in, rows of K coalesced), the next chunk's global loads in flight
// during the current chunk's FMAs.  Every element of K is read W / 128 times from L2 instead of W / 4 times.
constexpr int FT_W = 128, FT_S = 64, FT_K = 16;
__global__ void __launch_bounds__(256, 2) lmc_ewald_field_tiled_kernel(const DevModel m, const int8_t* __restrict__ occ_g, int W,
                                                                    double* __restrict__ field) {
  __shared__ __align__(16) double Qs[FT_K][FT_W];
  __shared__ __align__(16) double Ks[FT_K][FT_S];
  const int tid = threadIdx.x;
  const int s0 = blockIdx.x * FT_S, w0 = blockIdx.y * FT_W;
  const int ty = tid >> 4, tx = tid & 15;          // walkers ty*8.., sites tx*2, tx*2+1, 32+tx*2, 33+tx*2 (consecutive lanes read
                                                   // consecutive 16-byte pieces of a Ks row: no bank conflicts)
  // loaders: Q chunk = 16 k x 128 walkers = 2048 charges, 8 per thread (walker = tid >> 1, k = (tid & 1) * 8 ..);
  // K chunk = 16 k x 64 sites = 1024 doubles, 4 per thread (k = tid >> 4, sites (tid & 15) * 4 ..)
  const int qw = tid >> 1, qk = (tid & 1) * 8;
  const int kk = tid >> 4, ks = (tid & 15) * 4;
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  double qreg[8], kreg[4];
  auto load_chunk_regs = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + qk + u, ww = w0 + qw;
      double q = 0.0;
      if (k < m.N && ww < W) q = __ldg(m.ewQD + k * m.ewW + occ_g[(size_t)ww * m.Npad + k]).x;
      qreg[u] = q;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + kk, sidx = s0 + ks + u;
      kreg[u] = (k < m.N && sidx < m.N) ? __ldg(m.ewK + (size_t)k * m.N + sidx) : 0.0;
    }
  };
  load_chunk_regs(0);
  for (int k0 = 0; k0 < m.N; k0 += FT_K) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u) Qs[qk + u][qw] = qreg[u];
#pragma unroll
    for (int u = 0; u < 4; ++u) Ks[kk][ks + u] = kreg[u];
    __syncthreads();
    if (k0 + FT_K < m.N) load_chunk_regs(k0 + FT_K);
#pragma unroll
    for (int k = 0; k < FT_K; ++k) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&Qs[k][ty * 8 + i]);
        a[i] = v.x; a[i + 1] = v.y;
      }
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&Ks[k][(j >> 1) * 32 + tx * 2]);
        b[j] = v.x; b[j + 1] = v.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ww = w0 + ty * 8 + i;
    if (ww >= W) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sidx = s0 + (j >> 1) * 32 + tx * 2 + (j & 1);
      if (sidx < m.N) field[(size_t)ww * m.N + sidx] = acc[i][j];
    }
  }
}

// MCBias.compute_bias for every walker (bias.py:180-191, 276-287, 342-353): one warp per walker.
// sum[w][r] = sum_k tab[k][occ_k][r] - icpt[r]; bias = sum[w][0] (table sum) or -penalty * sum_r sum[w][r]^2
struct BiasIcpt { double v[LMC_MAX_BIAS_ROWS]; };
__global__ void lmc_bias_init_kernel(const int8_t* __restrict__ occ_g, int W, int N, int Npad, int mode, int bw, int rows,
                                     double pen, BiasIcpt icpt, const double* __restrict__ tab, double* __restrict__ bias,
                                     double* __restrict__ sum) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= W) return;
  double q = 0.0, first = 0.0;
  for (int r = 0; r < rows; ++r) {
    double c = 0.0;
    for (int k = lane; k < N; k += 32) c += tab[((size_t)k * bw + occ_g[(size_t)w * Npad + k]) * rows + r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    c -= icpt.v[r];
    if (lane == 0) sum[(size_t)w * rows + r] = c;
    q += c * c;
    if (r == 0) first = c;
  }
  if (lane == 0) bias[w] = mode == LMC_BIAS_SQUARE_SUM ? -(pen * q) : first;
}

// DistanceProcessor.compute_feature_vector (distance.py:133-154) from the extensive features of
// lmc_full_kernel: one thread per walker (F is a few tens)
__global__ void lmc_distance_init_kernel(int W, int F, int size, const double* __restrict__ nat, double* __restrict__ feat,
                                         double* __restrict__ vec, double* __restrict__ enth, const double* __restrict__ target,
                                         double tol, int ngrp, const int* __restrict__ goff, const int* __restrict__ gidx,
                                         const double* __restrict__ gdiam) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double* fw = feat + (size_t)w * F;
  double* vw = vec + (size_t)w * F;
  double e = 0.0;
  for (int f = 0; f < F; ++f) {
    const double v = fw[f] / (double)size;
    vw[f] = v;
    fw[f] = fabs(v - target[f]);
    if (f > 0) e += nat[f] * fw[f];
  }
  double L = 0.0;
  if (nat[0] != 0.0)
    for (int q = 0; q < ngrp; ++q) {
      bool ok = true;
      for (int i = goff[q]; i < goff[q + 1]; ++i) ok = ok && fw[gidx[i]] <= tol;
      if (!ok) break;
      L = gdiam[q];
    }
  fw[0] = L;
  if (enth) enth[w] = e + nat[0] * L;
}

// Ewald pair kernel between origin site o (blockIdx.y) and site k (blockIdx.x): one block per pair, threads
// stride over the reciprocal vectors and the real-space translations (cofe/extern/ewald.py:102-177 evaluates
// the same sums through pymatgen's EwaldSummation).
__global__ void lmc_ewald_site_kernel_k(const double* __restrict__ cart, const int* __restrict__ origins,
                                        const double* __restrict__ gv, const double* __restrict__ gc, int ng,
                                        const double* __restrict__ tv, int nt, double eta, double rcut, double vol,
                                        int nsites, double* __restrict__ out) {
  __shared__ double red[32];
  const int k = blockIdx.x, o = blockIdx.y, r0 = origins[o];
  const double dx = cart[3 * k] - cart[3 * r0], dy = cart[3 * k + 1] - cart[3 * r0 + 1], dz = cart[3 * k + 2] - cart[3 * r0 + 2];
  double rec = 0.0, real = 0.0;
  for (int i = threadIdx.x; i < ng; i += blockDim.x)
    rec += gc[i] * cos(gv[3 * i] * dx + gv[3 * i + 1] * dy + gv[3 * i + 2] * dz);
  const double rse = sqrt(eta);
  for (int i = threadIdx.x; i < nt; i += blockDim.x) {
    const double x = dx + tv[3 * i], y = dy + tv[3 * i + 1], z = dz + tv[3 * i + 2];
    const double r = sqrt(x * x + y * y + z * z);
    if (r > 1e-8 && r <= rcut) real += erfc(rse * r) / r;
  }
  const double v = block_sum((2.0 * 3.14159265358979323846 / vol) * rec + 0.5 * real, red);
  if (threadIdx.x == 0) out[(size_t)o * nsites + k] = v;
}

__global__ void lmc_cast_i32_i8_kernel(const int* __restrict__ src, int8_t* __restrict__ dst, int W, int N, int Npad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)W * Npad) return;
  const int w = (int)(i / Npad), k = (int)(i % Npad);
  dst[i] = k < N ? (int8_t)src[(size_t)w * N + k] : 0;
}
__global__ void lmc_cast_i8_i32_kernel(const int8_t* __restrict__ src, int* __restrict__ dst, long long rows, int N,
                                       int stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * N) return;
  const long long w = i / N;
  const int k = (int)(i % N);
  dst[i] = (int)src[w * stride + k];
}

#endif  // LMC_API_TU

}  // namespace lmc
