// Speculative-batch Metropolis kernel (sm_100a).
//
// The classic kernel (lmc_kernels.cuh) spends all 32 lanes of a warp on ONE attempted step: the
// per-step scalar work (proposal, rank select, reduction, accept test) is replicated 32 wide and
// dominates the instruction count.  Here a warp still owns one walker, but each group of
// SPEC_SG = 4 lanes evaluates a DIFFERENT upcoming step (SPEC_B = 32 / SPEC_SG = 8 consecutive steps per batch)
// against the current occupancy.  The RNG is counter based (Philox keyed by step index), so step
// t + j is fully defined before step t has been decided.  Rejected steps leave the state
// untouched, hence every step of the batch up to and including the FIRST accepted one has been
// evaluated against exactly the state the sequential chain would have seen: that prefix is
// committed, the rest of the batch is discarded and proposed again from the new state.  The chain
// is the sequential Metropolis chain of smol (kernel/base.py:145-166, metropolis.py:31-49), step
// for step; only the amount of wasted work depends on the acceptance ratio.
//
// Evaluation uses the merged records and the pre-differenced table built at model creation
// (lmc_api.cu): three occupancy byte gathers and ONE table lookup per record, where a record
// carries a 4-site cluster, a 3-site cluster plus a pair term, or three pair/point terms.
// The accepted step is re-evaluated by the whole warp with the classic record path, which leaves
// the per-record differences in the stash for the feature update (cluster order of
// evaluator.pyx:253-263).
#pragma once
#include "lmc_kernels.cuh"

namespace lmc {

// SG lanes evaluate one upcoming step, 32 / SG steps per batch.  SG = 4 is what lmc_run launches; SG = 2
// and SG = 1 (LMC_SPEC_SG) replicate less scalar work per step but measured slower on B200: the record
// loads and occupancy gathers of 16 / 32 unrelated flip sites per instruction cost more L1 / shared-
// memory wavefronts than the saved instructions (profiles/r01_variants.md).
//
// LISTS (swap usher): instead of the rank select over species bit-planes, every (sublattice, code c)
// keeps the ASCENDING list of active positions whose code differs from c, so the k-th swap partner of
// Swap.propose_step (mcusher.py:190-196) is one shared-memory load.  An accepted swap a <-> b replaces
// one entry in the lists of a and b (warp-wide shift), which pays off while accepts are rare.

// sorted list of n u16 positions: remove X, insert Y (X in the list, Y not).  Whole warp.
__device__ __forceinline__ void spec_list_replace(uint16_t* list, int n, int X, int Y, int g) {
  if (Y > X) {   // entries in (X, Y) move one slot down
    for (int base = 0; base < n; base += 32) {
      const int i = base + g;
      const int v = i < n ? (int)list[i] : 0xffff;
      const int nx = i + 1 < n ? (int)list[i + 1] : 0xffff;
      __syncwarp();
      if (v >= X && v < Y) list[i] = (uint16_t)(nx < Y ? nx : Y);
    }
  } else {       // entries in (Y, X) move one slot up
    for (int base = ((n - 1) >> 5) << 5; base >= 0; base -= 32) {
      const int i = base + g;
      const int v = i < n ? (int)list[i] : 0xffff;
      const int pv = (i > 0 && i - 1 < n) ? (int)list[i - 1] : -1;
      __syncwarp();
      if (v > Y && v <= X) list[i] = (uint16_t)(pv > Y ? pv : Y);
    }
  }
  __syncwarp();
}

// position of the (rem+1)-th set bit of b (rem < popc(b))
__device__ __forceinline__ int spec_nth_bit(uint32_t b, int rem) {
  int pos = 0;
  const int c16 = __popc(b & 0xffffu);
  if (rem >= c16) { rem -= c16; pos += 16; b >>= 16; }
  const int c8 = __popc(b & 0xffu);
  if (rem >= c8) { rem -= c8; pos += 8; b >>= 8; }
  const int c4 = __popc(b & 0xfu);
  if (rem >= c4) { rem -= c4; pos += 4; b >>= 4; }
  const int c2 = __popc(b & 0x3u);
  if (rem >= c2) { rem -= c2; pos += 2; b >>= 2; }
  if (rem >= (int)(b & 1u)) pos += 1;
  return pos;
}

// k-th (0-based) active position of sublattice `sl` whose code differs from `code`; the SG lanes
// of a subgroup split the plane words.  Called by all 32 lanes (full-mask shuffles of width SG).
template <int SPEC_SG>
__device__ __forceinline__ int spec_select_ne(const DevModel& m, const uint32_t* planes, int sl, int code, int k, int l) {
  const int n_act = m.sl_off[sl + 1] - m.sl_off[sl];
  const int nw = m.sl_nwords[sl];
  const uint32_t* pl = planes + m.sl_plane_off[sl] + code * nw;
  const uint32_t tail = (n_act & 31) ? ((1u << (n_act & 31)) - 1u) : 0xffffffffu;
  const int cw = (nw + SPEC_SG - 1) / SPEC_SG;
  const int lo = l * cw;
  int res = 0;
  bool found;
  if (cw <= 4) {
    // up to four words per lane (<= 512 active sites): straight-line, no divergence
    uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
    if (lo < nw) w0 = ~pl[lo] & (lo == nw - 1 ? tail : 0xffffffffu);
    if (cw > 1 && lo + 1 < nw) w1 = ~pl[lo + 1] & (lo + 1 == nw - 1 ? tail : 0xffffffffu);
    if (cw > 2 && lo + 2 < nw) w2 = ~pl[lo + 2] & (lo + 2 == nw - 1 ? tail : 0xffffffffu);
    if (cw > 3 && lo + 3 < nw) w3 = ~pl[lo + 3] & (lo + 3 == nw - 1 ? tail : 0xffffffffu);
    const int p1 = __popc(w0), p2 = p1 + __popc(w1), p3 = p2 + __popc(w2), cnt = p3 + __popc(w3);
    int incl = cnt;
    int tt = __shfl_up_sync(0xffffffffu, incl, 1, SPEC_SG);
    if (l >= 1) incl += tt;
    if (SPEC_SG > 2) {
      tt = __shfl_up_sync(0xffffffffu, incl, 2, SPEC_SG);
      if (l >= 2) incl += tt;
    }
    int rem = k - (incl - cnt);
    found = rem >= 0 && rem < cnt;
    const int wi = (rem >= p1) + (rem >= p2) + (rem >= p3);
    const uint32_t word = wi == 0 ? w0 : wi == 1 ? w1 : wi == 2 ? w2 : w3;
    rem -= wi == 0 ? 0 : wi == 1 ? p1 : wi == 2 ? p2 : p3;
    res = (lo + wi) * 32 + spec_nth_bit(word, found ? rem : 0);
  } else {
    int cnt = 0;
    for (int i = 0; i < cw; ++i) {
      const int wd = lo + i;
      uint32_t b = 0u;
      if (wd < nw) b = ~pl[wd] & (wd == nw - 1 ? tail : 0xffffffffu);
      cnt += __popc(b);
    }
    int incl = cnt;
    int tt = __shfl_up_sync(0xffffffffu, incl, 1, SPEC_SG);
    if (l >= 1) incl += tt;
    if (SPEC_SG > 2) {
      tt = __shfl_up_sync(0xffffffffu, incl, 2, SPEC_SG);
      if (l >= 2) incl += tt;
    }
    int rem = k - (incl - cnt);
    found = rem >= 0 && rem < cnt;
    if (found) {
      for (int i = 0; i < cw; ++i) {
        const int wd = lo + i;
        uint32_t b = 0u;
        if (wd < nw) b = ~pl[wd] & (wd == nw - 1 ? tail : 0xffffffffu);
        const int c = __popc(b);
        if (rem < c) { res = wd * 32 + spec_nth_bit(b, rem); break; }
        rem -= c;
      }
    }
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, found);
  const uint32_t mine = (bal >> (threadIdx.x & 31u & ~(uint32_t)(SPEC_SG - 1))) & ((1u << SPEC_SG) - 1u);
  const int own = mine ? (__ffs(mine) - 1) : 0;
  return __shfl_sync(0xffffffffu, res, own, SPEC_SG);
}

// occupancy code of site s; PATCH: site `ps` reads as `pc` (the first flip of a swap applied)
template <bool PATCH>
__device__ __forceinline__ uint32_t spec_code(const uint8_t* occ, uint32_t s, uint32_t ps, uint32_t pc) {
  uint32_t c = occ[s];
  if (PATCH) c = (s == ps) ? pc : c;
  return c;
}

// two merged records (one 16-byte chunk): three gathers + one table lookup each
template <bool PATCH>
__device__ __forceinline__ void spec_rec2(const uint8_t* occ, const double* Dn, const uint4 v, uint32_t NC, uint32_t old3,
                                          uint32_t ps, uint32_t pc, double& a0, double& a1) {
  a0 += Dn[(v.y >> 16) + old3 +
           NC * (spec_code<PATCH>(occ, v.x & 0xffffu, ps, pc) +
                 NC * (spec_code<PATCH>(occ, v.x >> 16, ps, pc) + NC * spec_code<PATCH>(occ, v.y & 0xffffu, ps, pc)))];
  a1 += Dn[(v.w >> 16) + old3 +
           NC * (spec_code<PATCH>(occ, v.z & 0xffffu, ps, pc) +
                 NC * (spec_code<PATCH>(occ, v.z >> 16, ps, pc) + NC * spec_code<PATCH>(occ, v.w & 0xffffu, ps, pc)))];
}

// scaled energy change of one flip (lane l of SG takes every SG-th 16-byte chunk of the record list)
template <int SPEC_SG>
__device__ __forceinline__ double spec_flip_energy(const DevModel& m, const uint8_t* occ, const double* dtab, int site,
                                                   int oldc, int newc, int l) {
  const uint4* rp = reinterpret_cast<const uint4*>(m.sp_rec + (size_t)site * m.spSb) + l;
  const double* Dn = dtab + newc * m.spL;
  const uint32_t NC = (uint32_t)m.spNC;
  const uint32_t old3 = (uint32_t)oldc;   // the old code is the fastest index of a block
  double a0 = 0.0, a1 = 0.0;
  const int nchunk = m.spNQ / (2 * SPEC_SG);   // spNQ is a multiple of 8
#pragma unroll 3
  for (int q = 0; q < nchunk; ++q) spec_rec2<false>(occ, Dn, __ldg(rp + q * SPEC_SG), NC, old3, 0u, 0u, a0, a1);
  return a0 + a1;
}

// both flips of a swap in the same loop (independent chains: twice the ILP); flip b sees flip a
// applied through the PATCH of its gathers (sequential semantics of expansion.py:217-229)
template <int SPEC_SG>
__device__ __forceinline__ double spec_swap_energy(const DevModel& m, const uint8_t* occ, const double* dtab, int sitea,
                                                   int olda, int newa, int siteb, int oldb, int newb, int l) {
  const uint4* ra = reinterpret_cast<const uint4*>(m.sp_rec + (size_t)sitea * m.spSb) + l;
  const uint4* rb = reinterpret_cast<const uint4*>(m.sp_rec + (size_t)siteb * m.spSb) + l;
  const double* Da = dtab + newa * m.spL;
  const double* Db = dtab + newb * m.spL;
  const uint32_t NC = (uint32_t)m.spNC;
  const uint32_t oa3 = (uint32_t)olda, ob3 = (uint32_t)oldb;   // the old code is the fastest index of a block
  const uint32_t ps = (uint32_t)sitea, pc = (uint32_t)newa;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  const int nchunk = m.spNQ / (2 * SPEC_SG);
#pragma unroll 3
  for (int q = 0; q < nchunk; ++q) {
    const uint4 va = __ldg(ra + q * SPEC_SG), vb = __ldg(rb + q * SPEC_SG);
    spec_rec2<false>(occ, Da, va, NC, oa3, 0u, 0u, a0, a1);
    spec_rec2<true>(occ, Db, vb, NC, ob3, ps, pc, b0, b1);
  }
  return (a0 + a1) + (b0 + b1);
}

// ---------------------------------------------------------------------------------------------------------------
// ENV variants: environment words instead of occupancy gathers (tables: build_env_tables, lmc_api.cu).
// The walker's words live in global memory (a.env, L2 resident: 16 or 32 bytes per active site), are read with
// ld.global.cg and updated by accepted steps with red.global.xor; one warp owns them.

// lane chunk of active site `ai`
__device__ __forceinline__ unsigned long long env_load(const DevModel& m, const uint32_t* envw, int ai, int l) {
  if (m.envWide) return __ldcg(reinterpret_cast<const unsigned long long*>(envw) + (size_t)ai * 4 + l);
  return (unsigned long long)__ldcg(envw + (size_t)ai * 4 + l);
}

// slot mask of active site `aj` inside lane chunk l of active site `ai` (0 where ai does not gather aj)
__device__ __forceinline__ unsigned long long env_pair(const DevModel& m, int ai, int aj, int l) {
  const size_t at = ((size_t)ai * m.envNA + aj) * 4 + l;
  // (ld.global.cg: the table is megabytes of mostly zero words read at random; it must not evict the L1-resident tables)
  if (m.envWide) return __ldcg(reinterpret_cast<const unsigned long long*>(m.envPair) + at);
  return (unsigned long long)__ldcg(reinterpret_cast<const uint32_t*>(m.envPair) + at);
}

// field of three EB-bit codes -> c0 + NC (c1 + NC c2); P2: NC == 2^EB, the field IS that number
template <int EB, bool P2>
__device__ __forceinline__ uint32_t env_cidx(uint32_t field, uint32_t NC) {
  if (P2) return field;
  constexpr uint32_t cm = (1u << EB) - 1u;
  return (field & cm) + NC * (((field >> EB) & cm) + NC * (field >> (2 * EB)));
}

// NP record pairs of a lane (one 16-byte group of table bases holds four); c = the chunk shifted to the group's first
// field.  a0: even records of the lane, a1: odd ones -- the order spec_rec2 sums them in.
template <int EB, bool P2, int NP>
__device__ __forceinline__ void env_pairs(const double* Dn, unsigned long long c, const uint4 tv, uint32_t NC, double& a0, double& a1) {
  constexpr uint32_t FB = 3 * EB, FM = (1u << FB) - 1u;
  const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    a0 += Dn[(tw[p] & 0xffffu) + NC * env_cidx<EB, P2>((uint32_t)(c >> (2 * p * FB)) & FM, NC)];
    a1 += Dn[(tw[p] >> 16) + NC * env_cidx<EB, P2>((uint32_t)(c >> ((2 * p + 1) * FB)) & FM, NC)];
  }
}

// (tbp: the lane's table-base list, in shared memory when the model's site classes are staged, else global)
template <int EB, bool P2>
__device__ __forceinline__ double env_flip_energy_p(const DevModel& m, const double* Dn, unsigned long long chunk, const uint4* tbp,
                                                    uint32_t NC) {
  constexpr int FB = 3 * EB;
  double a0 = 0.0, a1 = 0.0;
  // record pairs per lane: straight-line code for up to four (one group of table bases; the FCC and rocksalt cluster
  // sets have three), whole groups beyond that (the pad records of the last group add exact zeros)
  switch (m.envNRL) {   // kernel parameter: a uniform branch
    case 2: env_pairs<EB, P2, 1>(Dn, chunk, tbp[0], NC, a0, a1); break;
    case 4: env_pairs<EB, P2, 2>(Dn, chunk, tbp[0], NC, a0, a1); break;
    case 6: env_pairs<EB, P2, 3>(Dn, chunk, tbp[0], NC, a0, a1); break;
    case 8: env_pairs<EB, P2, 4>(Dn, chunk, tbp[0], NC, a0, a1); break;
    default:
      for (int i0 = 0; i0 < m.envNRL; i0 += 8) env_pairs<EB, P2, 4>(Dn, chunk >> (i0 * FB), tbp[i0 >> 3], NC, a0, a1);
  }
  return a0 + a1;
}

// scaled energy change of one flip from the lane's chunk: the records and table entries of spec_flip_energy
template <int EB>
__device__ __forceinline__ double env_flip_energy(const DevModel& m, const unsigned char* smem, const double* dtab,
                                                  unsigned long long chunk, int site, int oldc, int newc, int l) {
  const uint4* tbp = m.envNCls
      ? reinterpret_cast<const uint4*>(smem + m.off_envtb) + (((int)smem[m.off_envcls + site] * 4 + l) * m.envNRLP >> 3)
      : reinterpret_cast<const uint4*>(m.envTb + ((size_t)site * 4 + l) * m.envNRLP);
  const double* Dn = dtab + newc * m.spL + oldc;   // the old code is the fastest index of a block
  const uint32_t NC = (uint32_t)m.spNC;
  if (NC == (1u << EB)) return env_flip_energy_p<EB, true>(m, Dn, chunk, tbp, NC);
  return env_flip_energy_p<EB, false>(m, Dn, chunk, tbp, NC);
}

// words of every active site from the occupancy row (whole warp, start of a launch)
__device__ __forceinline__ void env_build(const DevModel& m, uint32_t* envw, const uint8_t* occ, int g) {
  const uint32_t b = (uint32_t)m.envB, fb = 3u * b;
  const int nrl = m.envNRL;
  for (int ai = g; ai < m.envNA; ai += 32) {
    const int site = __ldg(m.sl_sites + ai);
    const uint2* rp = reinterpret_cast<const uint2*>(m.sp_rec + (size_t)site * m.spSb);
    for (int l = 0; l < 4; ++l) {
      unsigned long long chunk = 0ull;
      for (int i = 0; i < nrl; ++i) {
        const uint2 rc = __ldg(rp + 2 * (l + 4 * (i >> 1)) + (i & 1));
        const uint32_t field = (uint32_t)occ[rc.x & 0xffffu] | ((uint32_t)occ[rc.x >> 16] << b) | ((uint32_t)occ[rc.y & 0xffffu] << (2u * b));
        chunk |= (unsigned long long)field << (i * fb);
      }
      if (m.envWide) reinterpret_cast<unsigned long long*>(envw)[(size_t)ai * 4 + l] = chunk;
      else envw[(size_t)ai * 4 + l] = (uint32_t)chunk;
    }
  }
}

// an accepted flip of active site `ai` (code old -> new, x = old ^ new): every site that gathers it sees the new code
__device__ __forceinline__ void env_commit(const DevModel& m, uint32_t* envw, int ai, uint32_t x, int g) {
  const uint32_t* rv = m.envRev + (size_t)ai * m.envRV;
  for (int e = g; e < m.envRV; e += 32) {
    const uint32_t ent = __ldg(rv + e);
    if (ent == 0xffffffffu) continue;
    const uint32_t k = ent & 0xffffu, p = ent >> 16;
    if (m.envWide) atomicXor(reinterpret_cast<unsigned long long*>(envw) + (size_t)k * 4 + (p >> 6), (unsigned long long)x << (p & 63u));
    else atomicXor(envw + (size_t)k * 4 + (p >> 5), x << (p & 31u));
  }
}

// EWF: Ewald term through the potential cache (a.ew_field, see lmc_kernels.cuh): two cached doubles and
// a charge table lookup per flip, one row of the site kernel per ACCEPTED flip.
// MAXT / MINB: launch bounds.  (128, 7) while seven blocks of four walkers fit an SM's shared memory;
// (448, 2) -- fourteen walkers share one copy of a larger table blob -- otherwise: both keep 28 walkers
// resident per SM (4096 walkers = one wave on 148 SMs) at 72 registers.
// EB > 0: environment words (a.env) of EB bits per species code instead of occupancy gathers, four lanes per step.
template <bool KONE, int USHER, int SPEC_SG, bool LISTS, bool EWF, int MAXT, int MINB, int EB = 0>
__global__ void __launch_bounds__(MAXT, MINB) lmc_spec_kernel(const DevModel m, const RunArgs a) {
  constexpr bool ENV = EB != 0;
  static_assert(!ENV || (SPEC_SG == 4 && !LISTS), "environment words: four lanes per step, rank select");
  static_assert(EB >= 0 && EB <= 2, "one or two bits per species code");
  static_assert(USHER == LMC_USHER_FLIP || USHER == LMC_USHER_SWAP, "flip / swap only");
  static_assert(SPEC_SG == 1 || SPEC_SG == 2 || SPEC_SG == 4, "1, 2 or 4 lanes per speculated step");
  static_assert(!LISTS || USHER == LMC_USHER_SWAP, "position lists serve the swap usher");
  constexpr int SPEC_B = 32 / SPEC_SG;
  constexpr int G = 32;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint64_t bar;
  const int g = threadIdx.x & 31;
  const int wl_ = threadIdx.x >> 5;                 // walker slot in block
  const int w = blockIdx.x * a.wpb + wl_;
  const int nw_blk = min(a.wpb, a.W - blockIdx.x * a.wpb);
  const bool active = wl_ < a.wpb && w < a.W;
  const int sg = g / SPEC_SG, l = g % SPEC_SG;

  const int blob_bytes = EB != 0 ? m.blob_env_bytes : m.blob_bytes;
  unsigned char* wbase = smem + ((blob_bytes + 15) & ~15);
  uint8_t* occ_rows = wbase;
  unsigned char* rest = wbase + (size_t)a.wpb * m.Npad;
  uint8_t* occ = occ_rows + (size_t)wl_ * m.Npad;
  unsigned char* priv = rest + (size_t)wl_ * a.walker_smem;
  double* feat = reinterpret_cast<double*>(priv + a.off_feat);
  unsigned char* stash0 = priv + a.off_stash;
  int* cnt = reinterpret_cast<int*>(priv + a.off_cnt);
  uint32_t* planes = reinterpret_cast<uint32_t*>(priv + a.off_plane);
  uint4* ring = reinterpret_cast<uint4*>(priv + a.off_ring);   // [32] x (sl<<24 | pos, site, word z, float log u)
  uint16_t* lists = reinterpret_cast<uint16_t*>(priv + a.off_lists);   // LISTS: [sublattice][code][n_active]

  stage_tables(m, smem, &bar, occ_rows, a.occ + (size_t)blockIdx.x * a.wpb * m.Npad, (uint32_t)(nw_blk * m.Npad),
               (uint32_t)blob_bytes);
  const SmemTables t = smem_tables(m, smem);
  const double* dtab = reinterpret_cast<const double*>(smem + m.off_dtab);
  const double* ctab = reinterpret_cast<const double*>(smem + m.off_ctab);
  if (!active) return;
  if (g == 0) occ[m.N] = 0;   // pad byte behind the row: the zero code gathered by unused record slots

  const int stash_stride = m.Rstride * (KONE ? 8 : 4);
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
  __syncwarp();
  if (LISTS) {
    // lists[sl][c] = ascending active positions with code != c, expanded from the complement of plane c
    for (int sl = 0; sl < m.nSl; ++sl) {
      const int n_act = m.sl_off[sl + 1] - m.sl_off[sl], nw = m.sl_nwords[sl];
      const uint32_t tail = (n_act & 31) ? ((1u << (n_act & 31)) - 1u) : 0xffffffffu;
      for (int c = 0; c < m.sl_nplanes[sl]; ++c) {
        uint16_t* lst = lists + m.sl_list_off[sl] + c * n_act;
        int carry = 0;
        for (int w0 = 0; w0 < nw; w0 += 32) {
          const int wd = w0 + g;
          uint32_t b = wd < nw ? (~planes[m.sl_plane_off[sl] + c * nw + wd] & (wd == nw - 1 ? tail : 0xffffffffu)) : 0u;
          const int pc = __popc(b);
          int incl = pc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int tt = __shfl_up_sync(FULL, incl, o);
            if (g >= o) incl += tt;
          }
          int at = carry + incl - pc;
          while (b) {
            lst[at++] = (uint16_t)(wd * 32 + __ffs(b) - 1);
            b &= b - 1u;
          }
          carry += __shfl_sync(FULL, incl, 31);
        }
      }
    }
    __syncwarp();
  }

  uint32_t* envw = ENV ? a.env + (size_t)w * m.envNA * (m.envWide ? 8 : 4) : nullptr;
  if (ENV) {
    env_build(m, envw, occ, g);
    __threadfence();
    __syncwarp();
  }

  const unsigned long long seed = a.seeds[w];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t wid = (uint32_t)(a.walker_base + w);
  const double beta = a.beta[w];
  constexpr bool MU_POSSIBLE = USHER != LMC_USHER_SWAP;
  const double nat_mu = (MU_POSSIBLE && m.muW) ? t.nat[m.muF] : 0.0;
  const double nat_ew = EWF ? t.nat[m.ewF] : 0.0;
  double* fld = EWF ? a.ew_field + (size_t)w * m.N : nullptr;

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
      uint4 rq;
      if (SPEC_SG == 1) {
        // one step per lane: the lane's own Philox block, no staging
        const unsigned long long st_ = step + (unsigned long long)g;
        const U4 bq = philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1);
        const int sl_ = choose_sublattice(m, bq.x);
        const int j_ = (int)mulhi32(bq.y, (uint32_t)(m.sl_off[sl_ + 1] - m.sl_off[sl_]));
        rq = make_uint4((uint32_t)((sl_ << 24) | j_), (uint32_t)site_of_pos(m, sl_, j_), bq.z,
                        __float_as_uint(log_u_float(bq.w)));
      } else if (!ring_valid || step + (unsigned long long)nb > rbase + 32ull) {
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
      if (SPEC_SG > 1) rq = ring[min((int)(step - rbase) + sg, 31)];
      const int sl = (int)(rq.x >> 24), pos1 = (int)(rq.x & 0xffffffu), site1 = (int)rq.y;
      const float lf = __uint_as_float(rq.w);
      // environment words of the first site: state-independent address, the load overlaps the proposal
      const int ai1 = ENV ? m.sl_off[sl] + pos1 : 0;
      unsigned long long ch1 = 0ull;
      if (ENV) ch1 = env_load(m, envw, ai1, l);

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
        if (LISTS) pos2 = ndiff > 0 ? (int)lists[m.sl_list_off[sl] + s1 * n_act + k] : 0;
        else pos2 = spec_select_ne<SPEC_SG == 1 ? 2 : SPEC_SG>(m, planes, sl, s1, k, l);
        if (ndiff > 0) {
          site2 = site_of_pos(m, sl, pos2);
          s2 = occ[site2];
          n = 2;
        }
      }

      // site 2 sees site 1 already holding s2 (expansion.py:217-229): its slots that gather site 1 flip s1 -> s2.
      // Both loads are in flight while flip 1 is evaluated.
      unsigned long long ch2 = 0ull, pm2 = 0ull;
      if (ENV && USHER == LMC_USHER_SWAP) {
        const int ai2 = m.sl_off[sl] + pos2;
        ch2 = env_load(m, envw, ai2, l);
        pm2 = env_pair(m, ai2, ai1, l);
      }

      // ------------------------------ evaluate ------------------------------------------------
      // Ewald term first: its (L2) loads are in flight while the cluster records are evaluated
      double dEw = 0.0;
      if (EWF && n > 0) {
        const double2 qn = ewald_qd_s(m, ctab, site1, s2, sl), qo = ewald_qd_s(m, ctab, site1, s1, sl);
        const double dq1 = qn.x - qo.x;
        dEw = 2.0 * dq1 * fld[site1] + (qn.y - qo.y);
        if (USHER == LMC_USHER_SWAP) {   // site 2 takes s1; it sees the cache shifted by flip 1
          const double2 qn2 = ewald_qd_s(m, ctab, site2, s1, sl), qo2 = ewald_qd_s(m, ctab, site2, s2, sl);
          dEw += 2.0 * (qn2.x - qo2.x) * (fld[site2] + dq1 * __ldg(m.ewK + (size_t)site1 * m.N + site2)) + (qn2.y - qo2.y);
        }
      }
      double acc = 0.0, dmu = 0.0;
      if (live && n > 0) {
        if (ENV) {
          constexpr int EBK = ENV ? EB : 1;
          const double ea = env_flip_energy<EBK>(m, smem, dtab, ch1, site1, s1, s2, l);
          acc = USHER == LMC_USHER_FLIP ? ea
              : ea + env_flip_energy<EBK>(m, smem, dtab, ch2 ^ (pm2 * (unsigned long long)(s1 ^ s2)), site2, s2, s1, l);
        } else if (USHER == LMC_USHER_FLIP) acc = spec_flip_energy<SPEC_SG>(m, occ, dtab, site1, s1, s2, l);
        else acc = spec_swap_energy<SPEC_SG>(m, occ, dtab, site1, s1, s2, site2, s2, s1, l);
      }
      if (SPEC_SG > 1) acc += __shfl_xor_sync(FULL, acc, 1);
      if (SPEC_SG > 2) acc += __shfl_xor_sync(FULL, acc, 2);
      double dH = acc;
      if (EWF) dH += nat_ew * dEw;
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
      if (EWF) {
        // accepted: the changed charges shift the potential cache (read again by the next batch: the
        // __syncwarp at the end of the commit orders these stores before those loads)
        const double c_dEw = __shfl_sync(FULL, dEw, src);
        const double dq1 = ewald_qd(m, c_site1, c_s2).x - ewald_qd(m, c_site1, c_s1).x;
        const double dq2 = c_n == 2 ? ewald_qd(m, c_site2, c_s1).x - ewald_qd(m, c_site2, c_s2).x : 0.0;
        const double* k1 = m.ewK + (size_t)c_site1 * m.N;
        const double* k2 = m.ewK + (size_t)(c_n == 2 ? c_site2 : c_site1) * m.N;
        for (int k = g; k < m.N; k += G) fld[k] += dq1 * __ldg(k1 + k) + dq2 * __ldg(k2 + k);
        if (g == 0) feat[m.ewF] += c_dEw;
      }
      if (c_n == 2) {
        if (g == 0) occ[c_site1] = (uint8_t)c_s2;
        __syncwarp();
        const RecChunk pre0 = load_records<G>(m, c_site1, g), pre1 = load_records<G>(m, c_site2, g);
        (void)flip_energy_pair<G, KONE>(m, t, occ, c_site1, c_s1, c_s2, c_site2, c_s2, c_s1, stash0, stash0 + stash_stride,
                                        g, pre0, pre1);
        __syncwarp();
        flip_features<G, KONE>(m, t, c_site1, stash0, feat, g, load_segment<G>(m, c_site1, g));
        __syncwarp();
        flip_features<G, KONE>(m, t, c_site2, stash0 + stash_stride, feat, g, load_segment<G>(m, c_site2, g));
        if (g == 0) {
          occ[c_site2] = (uint8_t)c_s1;
          const int nw = m.sl_nwords[c_sl];
          uint32_t* pl = planes + m.sl_plane_off[c_sl];
          const uint32_t bit1 = 1u << (c_pos1 & 31), bit2 = 1u << (c_pos2 & 31);
          pl[c_s1 * nw + (c_pos1 >> 5)] ^= bit1;
          pl[c_s2 * nw + (c_pos1 >> 5)] ^= bit1;
          pl[c_s2 * nw + (c_pos2 >> 5)] ^= bit2;
          pl[c_s1 * nw + (c_pos2 >> 5)] ^= bit2;
        }
        if (LISTS) {
          // position 1 now differs from code s1 and position 2 no longer does; the reverse for code s2
          const int n_act = m.sl_off[c_sl + 1] - m.sl_off[c_sl];
          uint16_t* lb = lists + m.sl_list_off[c_sl];
          spec_list_replace(lb + c_s1 * n_act, n_act - cnt[c_sl * LMC_MAX_CODES + c_s1], c_pos2, c_pos1, g);
          spec_list_replace(lb + c_s2 * n_act, n_act - cnt[c_sl * LMC_MAX_CODES + c_s2], c_pos1, c_pos2, g);
        }
      } else if (c_n == 1) {
        const RecChunk pre0 = load_records<G>(m, c_site1, g);
        (void)flip_energy<G, KONE>(m, t, occ, c_site1, c_s1, c_s2, stash0, g, pre0);
        __syncwarp();
        flip_features<G, KONE>(m, t, c_site1, stash0, feat, g, load_segment<G>(m, c_site1, g));
        if (g == 0) {
          occ[c_site1] = (uint8_t)c_s2;
          if (MU_POSSIBLE && m.muW) feat[m.muF] += c_dmu;
          cnt[c_sl * LMC_MAX_CODES + c_s1]--;
          cnt[c_sl * LMC_MAX_CODES + c_s2]++;
          const int nw = m.sl_nwords[c_sl];
          uint32_t* pl = planes + m.sl_plane_off[c_sl] + (c_pos1 >> 5);
          const uint32_t bit = 1u << (c_pos1 & 31);
          pl[c_s1 * nw] ^= bit;
          pl[c_s2 * nw] ^= bit;
        }
      }
      if (ENV && c_n > 0) {
        env_commit(m, envw, m.sl_off[c_sl] + c_pos1, (uint32_t)(c_s1 ^ c_s2), g);
        if (c_n == 2) env_commit(m, envw, m.sl_off[c_sl] + c_pos2, (uint32_t)(c_s1 ^ c_s2), g);
        __threadfence();   // the next batch reads the words (ld.global.cg) after the __syncwarp below
      }
      __syncwarp();
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
