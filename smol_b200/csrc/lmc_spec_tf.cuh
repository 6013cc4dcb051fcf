// Speculative-batch Metropolis kernel for the TABLE-FLIP usher (sm_100a).
//
// The classic kernel (lmc_kernels.cuh) spends a whole warp on one table-flip step: two Philox blocks, the
// direction choice, up to eight sequential site picks (a rank select over species bit-planes each) and the
// a-priori factor are scalar work replicated 32 wide -- ~2400 warp instructions per step on BASELINE config 5,
// of which the record evaluation is a tenth.  Here, as in lmc_spec.cuh, a warp still owns one walker but each
// group of four lanes proposes and evaluates a DIFFERENT upcoming step (eight per batch) against the current
// state; steps up to the first accepted one are exactly the sequential chain of smol (kernel/base.py:145-166,
// mcusher.py:553-711), the accepted step is committed by the whole warp and the rest is proposed again.
//
// A step changes up to four sites.  Flip f of a step must see flips < f of the SAME step applied
// (expansion.py:217-229) while the other seven steps of the batch read the unchanged occupancy, so nothing is
// written during the evaluation: the gathers of flip f compare their site against the f earlier ones
// (spec_flip_energy_n<f>).  The Ewald term comes from the walker's potential cache (lmc.h: ewald_field_dev) with
// the cross terms K[site_h][site_f] of the step's earlier flips, as in the classic kernel.
#pragma once
#include "lmc_spec.cuh"

namespace lmc {

// occupancy code of site s with the NP earlier flips of the step applied
template <int NP>
__device__ __forceinline__ uint32_t spec_code_n(const uint8_t* occ, uint32_t s, const uint32_t (&ps)[3], const uint32_t (&pc)[3]) {
  uint32_t c = occ[s];
#pragma unroll
  for (int i = 0; i < NP; ++i) c = (s == ps[i]) ? pc[i] : c;
  return c;
}

// scaled energy change of one flip over the merged records (lane l of four takes every fourth 16-byte chunk)
template <int NP>
__device__ __forceinline__ double spec_flip_energy_n(const DevModel& m, const uint8_t* occ, const double* dtab, int site, int oldc,
                                                     int newc, int l, const uint32_t (&ps)[3], const uint32_t (&pc)[3]) {
  const uint4* rp = reinterpret_cast<const uint4*>(m.sp_rec + (size_t)site * m.spSb) + l;
  const double* Dn = dtab + newc * m.spL + oldc;   // the old code is the fastest index of a block
  const uint32_t NC = (uint32_t)m.spNC;
  double a0 = 0.0, a1 = 0.0;
  const int nchunk = m.spNQ / 8;   // spNQ is a multiple of 8
#pragma unroll 4
  for (int q = 0; q < nchunk; ++q) {
    const uint4 v = __ldg(rp + q * 4);
    a0 += Dn[(v.y >> 16) + NC * (spec_code_n<NP>(occ, v.x & 0xffffu, ps, pc) +
                                 NC * (spec_code_n<NP>(occ, v.x >> 16, ps, pc) + NC * spec_code_n<NP>(occ, v.y & 0xffffu, ps, pc)))];
    a1 += Dn[(v.w >> 16) + NC * (spec_code_n<NP>(occ, v.z & 0xffffu, ps, pc) +
                                 NC * (spec_code_n<NP>(occ, v.z >> 16, ps, pc) + NC * spec_code_n<NP>(occ, v.w & 0xffffu, ps, pc)))];
  }
  return a0 + a1;
}

// The flips of one proposed step, packed into scalar registers (arrays written at a run-time index, as Step<MF> of the
// classic kernel, were placed in local memory here and every later read of a site waited on L2): 16-bit sites and
// positions, 4-bit codes and sublattices, flip f in field f.
struct PackedStep {
  unsigned long long site, pos;
  uint32_t codes;   // old codes in bits [4 f, 4 f + 4), new codes in bits [16 + 4 f, 16 + 4 f + 4)
  uint32_t sl;      // bits [4 f, 4 f + 4)
  int n;
  double log_priori;
  __device__ __forceinline__ void push(int site_, int oldc, int newc, int sl_, int pos_) {
    if (n < LMC_MAX_FLIPS) {
      site |= (unsigned long long)(uint32_t)site_ << (16 * n);
      pos |= (unsigned long long)(uint32_t)pos_ << (16 * n);
      codes |= ((uint32_t)oldc << (4 * n)) | ((uint32_t)newc << (16 + 4 * n));
      sl |= (uint32_t)sl_ << (4 * n);
      ++n;
    }
  }
  __device__ __forceinline__ int site_of(int f) const { return (int)((site >> (16 * f)) & 0xffffull); }
  __device__ __forceinline__ int pos_of(int f) const { return (int)((pos >> (16 * f)) & 0xffffull); }
  __device__ __forceinline__ int old_of(int f) const { return (int)((codes >> (4 * f)) & 0xfu); }
  __device__ __forceinline__ int new_of(int f) const { return (int)((codes >> (16 + 4 * f)) & 0xfu); }
  __device__ __forceinline__ int sl_of(int f) const { return (int)((sl >> (4 * f)) & 0xfu); }
};

// Descriptor tables of the usher and the sublattices, copied from the kernel parameters to shared memory once per
// block: the eight steps of a batch index them with DIFFERENT dimensions / sublattices in one instruction, which the
// constant bank serves one address at a time (the proposal waited on those loads more than on anything else)
struct TfShared {
  int table[LMC_MAX_TABLE_FLIPS][LMC_MAX_DIMS];
  int dim_sl[LMC_MAX_DIMS], dim_code[LMC_MAX_DIMS];
  int npick[2 * LMC_MAX_TABLE_FLIPS];
  uint32_t pick[2 * LMC_MAX_TABLE_FLIPS][2 * LMC_MAX_FLIPS];
  int sl_off[LMC_MAX_SUBLATTICES + 1], sl_nwords[LMC_MAX_SUBLATTICES], sl_plane_off[LMC_MAX_SUBLATTICES], sl_first[LMC_MAX_SUBLATTICES];
  double sl_cum[LMC_MAX_SUBLATTICES];
};

__device__ __forceinline__ int tf_site_of_pos(const DevModel& m, const TfShared& ts, int sl, int pos) {
  return ts.sl_first[sl] >= 0 ? ts.sl_first[sl] + pos : __ldg(m.sl_sites + ts.sl_off[sl] + pos);
}

// k-th (0-based) active position of sublattice `sl` that holds species `code` (ne: that does NOT hold it), from the
// bit-plane of the code and the exclusive prefix popcounts of its words: a binary search over the words and a rank
// select inside one, by every lane on its own (no shuffles; the scan of select_pos costs four times as much with
// four lanes per step and 54 words per plane)
__device__ __forceinline__ int tf_select(const TfShared& ts, const uint32_t* planes, const uint16_t* pfx, int sl, int code, int k, bool ne) {
  const int nw = ts.sl_nwords[sl];
  const int base = ts.sl_plane_off[sl] + code * nw;
  int lo = 0, hi = nw - 1;   // largest word whose prefix rank is <= k
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const int before = ne ? 32 * mid - (int)pfx[base + mid] : (int)pfx[base + mid];
    if (before <= k) lo = mid; else hi = mid - 1;
  }
  uint32_t word = planes[base + lo];
  int rem = k - (int)pfx[base + lo];
  if (ne) {
    const int n_act = ts.sl_off[sl + 1] - ts.sl_off[sl];
    const uint32_t tail = (n_act & 31) ? ((1u << (n_act & 31)) - 1u) : 0xffffffffu;
    word = ~word & (lo == nw - 1 ? tail : 0xffffffffu);
    rem = k - (32 * lo - (int)pfx[base + lo]);
  }
  return lo * 32 + spec_nth_bit(word, rem);
}

// EWF: Ewald term through the potential cache (a.ew_field).  One block per SM: up to 16 walkers share one copy of
// the staged tables (the difference table of a five-species model is ~50 KB).
template <bool KONE, bool EWF>
__global__ void __launch_bounds__(512, 1) lmc_spec_tf_kernel(const DevModel m, const RunArgs a) {
  constexpr int SG = 4, SPEC_B = 8, G = 32, MF = LMC_MAX_FLIPS, TF = LMC_MAX_TABLE_FLIPS;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ TfShared ts;
  const int g = threadIdx.x & 31;
  const int wl_ = threadIdx.x >> 5;                 // walker slot in block
  const int w = blockIdx.x * a.wpb + wl_;
  const int nw_blk = min(a.wpb, a.W - blockIdx.x * a.wpb);
  const bool active = wl_ < a.wpb && w < a.W;
  const int sg = g / SG, l = g % SG;
  for (int i = threadIdx.x; i < LMC_MAX_TABLE_FLIPS * LMC_MAX_DIMS; i += blockDim.x)
    ts.table[i / LMC_MAX_DIMS][i % LMC_MAX_DIMS] = m.tf_table[i / LMC_MAX_DIMS][i % LMC_MAX_DIMS];
  for (int i = threadIdx.x; i < LMC_MAX_DIMS; i += blockDim.x) { ts.dim_sl[i] = m.tf_dim_sl[i]; ts.dim_code[i] = m.tf_dim_code[i]; }
  for (int i = threadIdx.x; i < 2 * LMC_MAX_TABLE_FLIPS * 2 * LMC_MAX_FLIPS; i += blockDim.x)
    ts.pick[i / (2 * LMC_MAX_FLIPS)][i % (2 * LMC_MAX_FLIPS)] = m.tf_pick[i / (2 * LMC_MAX_FLIPS)][i % (2 * LMC_MAX_FLIPS)];
  for (int i = threadIdx.x; i < 2 * LMC_MAX_TABLE_FLIPS; i += blockDim.x) ts.npick[i] = m.tf_npick[i];
  for (int i = threadIdx.x; i < LMC_MAX_SUBLATTICES; i += blockDim.x) {
    ts.sl_off[i] = m.sl_off[i]; ts.sl_nwords[i] = m.sl_nwords[i]; ts.sl_plane_off[i] = m.sl_plane_off[i];
    ts.sl_first[i] = m.sl_first[i]; ts.sl_cum[i] = m.sl_cum[i];
    if (i == 0) ts.sl_off[LMC_MAX_SUBLATTICES] = m.sl_off[LMC_MAX_SUBLATTICES];
  }
  // (stage_tables below ends with a block-wide barrier: the tables are complete before anyone reads them)

  unsigned char* wbase = smem + ((m.blob_bytes + 15) & ~15);
  uint8_t* occ_rows = wbase;
  unsigned char* rest = wbase + (size_t)a.wpb * m.Npad;
  uint8_t* occ = occ_rows + (size_t)wl_ * m.Npad;
  unsigned char* priv = rest + (size_t)wl_ * a.walker_smem;
  double* feat = reinterpret_cast<double*>(priv + a.off_feat);
  unsigned char* stash0 = priv + a.off_stash;       // ONE slot: the commit evaluates and folds flip by flip.  Between
                                                    // commits the slot holds the two random-word rings (a.off_ring ==
                                                    // a.off_stash): a commit overwrites them, the next batch refills them
  int* cnt = reinterpret_cast<int*>(priv + a.off_cnt);
  uint32_t* planes = reinterpret_cast<uint32_t*>(priv + a.off_plane);
  uint16_t* pfx = reinterpret_cast<uint16_t*>(priv + a.off_eidx);   // set bits in the earlier words of the same plane
  uint4* ring = reinterpret_cast<uint4*>(priv + a.off_ring);   // [32] x (word 0, word 1, word 2, float log u) of block 0
  uint4* ring2 = ring + 32;                                    // [32] x block 1 (words 4..7 of the step)
  double* tfc = reinterpret_cast<double*>(priv + a.off_tfc);   // [weights][cumulative][log a-priori][sum], see lmc_kernels.cuh

  stage_tables(m, smem, &bar, occ_rows, a.occ + (size_t)blockIdx.x * a.wpb * m.Npad, (uint32_t)(nw_blk * m.Npad),
               (uint32_t)m.blob_bytes);
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
  __syncwarp();
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

  const unsigned long long seed = a.seeds[w];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t wid = (uint32_t)(a.walker_base + w);
  const double beta = a.beta[w];
  const double nat_mu = m.muW ? t.nat[m.muF] : 0.0;
  const double nat_ew = EWF ? t.nat[m.ewF] : 0.0;
  double* fld = EWF ? a.ew_field + (size_t)w * m.N : nullptr;

  unsigned long long step = a.step0;
  unsigned long long rbase = step;
  bool ring_valid = false, tfc_valid = false;
  long long nacc_total = 0;
  for (long long s = 0; s < a.S; ++s) {
    int nacc = 0;
    bool accepted = true;
    int it = 0;
    while (it < a.thin) {
      __syncwarp();
      const int nb = min(SPEC_B, a.thin - it);
      if (!ring_valid || step + (unsigned long long)nb > rbase + 32ull) {
        // random words of the next 32 steps, one step per lane: blocks 0 and 1
        rbase = step;
        const unsigned long long st_ = step + (unsigned long long)g;
        const U4 b0 = philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 0u, wid, k0, k1);
        const U4 b1 = philox4x32_10((uint32_t)st_, (uint32_t)(st_ >> 32), 1u, wid, k0, k1);
        __syncwarp();
        ring[g] = make_uint4(b0.x, b0.y, b0.z, __float_as_uint(log_u_float(b0.w)));
        ring2[g] = make_uint4(b1.x, b1.y, b1.z, b1.w);
        __syncwarp();
        ring_valid = true;
      }
      // species count of table dimension d (shared memory; a per-thread array would live in local memory)
      auto count_of = [&](int d) -> int {
        return ts.dim_sl[d] >= 0 ? cnt[ts.dim_sl[d] * LMC_MAX_CODES + ts.dim_code[d]] : 0;
      };
      if (!tfc_valid) {
        int nd[LMC_MAX_DIMS];
        for (int d = 0; d < m.tfD; ++d) nd[d] = count_of(d);
        // direction weights, cumulative probabilities and a-priori factors of the current species counts
        // (whole warp; the same expressions in the same order as the classic kernel)
        double tw[2 * TF];
        const double sum = tf_masked_weights<G>(m, nd, tw, g, FULL);
        __syncwarp();
        if (g == 0) {
          double cum = 0.0;
          for (int i = 0; i < 2 * m.tfNF; ++i) {
            cum += tw[i] / sum;
            tfc[i] = tw[i];
            tfc[2 * TF + i] = cum;
          }
          tfc[6 * TF] = sum;
        }
        for (int i = 0; i < 2 * m.tfNF; ++i) {
          double lfi = 0.0;
          if (sum > 0.0 && tw[i] > 0.0) {
            const int sgn_i = (i & 1) ? -1 : 1;
            const int* urow_i = m.tf_table[i >> 1];
            int nn[LMC_MAX_DIMS];
            for (int d = 0; d < m.tfD; ++d) nn[d] = nd[d] + sgn_i * urow_i[d];
            double tw2[2 * TF];
            const double sum2 = tf_masked_weights<G>(m, nn, tw2, g, FULL);
            const double p_now = (1.0 - m.tf_sw) * tw[i] / sum;
            const double p_next = (1.0 - m.tf_sw) * tw2[i ^ 1] / sum2;
            lfi = log(p_next / p_now);
            for (int d = 0; d < m.tfD; ++d)
              if (urow_i[d] != 0) lfi += __ldg(m.lgam + nd[d]) - __ldg(m.lgam + nn[d]);
          }
          if (g == 0) tfc[4 * TF + i] = lfi;
        }
        __syncwarp();
        tfc_valid = true;
      }

      // ------------------------------ propose (one step per four-lane group) ------------------
      const bool live = sg < nb;
      const int ri = min((int)(step - rbase) + sg, 31);
      const unsigned long long mystep = rbase + (unsigned long long)ri;
      const uint4 rq = ring[ri];
      const uint4 r1 = ring2[ri];
      const float lf = __uint_as_float(rq.w);
      PackedStep st;
      st.site = 0ull; st.pos = 0ull; st.codes = 0u; st.sl = 0u; st.n = 0;
      st.log_priori = 0.0;
      int tf_idx = -1;
      const double tfsum = tfc[6 * TF];
      const bool do_swap = u01(rq.x) < m.tf_sw || !(tfsum > 0.0);
      if (do_swap) {
        // fallback swap of the table-flip usher (mcusher.py:597-600): Swap.propose_step on words 4, 5, 6
        int sl = 0;   // choose_sublattice
        if (m.nSl > 1) {
          const double us = u01(r1.x);
          while (sl < m.nSl - 1 && !(ts.sl_cum[sl] > us)) ++sl;
        }
        const int n_act = ts.sl_off[sl + 1] - ts.sl_off[sl];
        const int j = (int)mulhi32(r1.y, (uint32_t)n_act);
        const int site1 = tf_site_of_pos(m, ts, sl, j);
        const int s1 = occ[site1];
        const int ndiff = n_act - cnt[sl * LMC_MAX_CODES + s1];
        if (ndiff > 0) {
          const int k = (int)mulhi32(r1.z, (uint32_t)ndiff);
          const int p2 = tf_select(ts, planes, pfx, sl, s1, k, true);
          const int site2 = tf_site_of_pos(m, ts, sl, p2);
          const int s2 = occ[site2];
          st.push(site1, s1, s2, sl, j);
          st.push(site2, s2, s1, sl, p2);
        }
      } else {
        // choose_section_from_partition, utils/math.py:870-893
        const double u = u01(rq.y);
        tf_idx = 2 * m.tfNF - 1;
        for (int i = 0; i < 2 * m.tfNF; ++i)
          if (tfc[2 * TF + i] > u && tfc[i] > 0.0) { tf_idx = i; break; }
        // table flip: sequential picks, one random word each (words 4.. of the step), mcusher.py:602-639
        // one descriptor per random word, in the reference's order (DevModel::tf_pick): the groups of the warp walk the
        // same loop whatever their direction (nested loops over dimensions and counts left them on different paths)
        int wi = 0;
        U4 rb{r1.x, r1.y, r1.z, r1.w};
        int cur_blk = 1;
        auto next_word = [&]() -> uint32_t {
          const int b = 1 + (wi >> 2);
          if (b != cur_blk) { rb = philox4x32_10((uint32_t)mystep, (uint32_t)(mystep >> 32), (uint32_t)b, wid, k0, k1); cur_blk = b; }
          const int c = wi & 3;
          ++wi;
          return c == 0 ? rb.x : c == 1 ? rb.y : c == 2 ? rb.z : rb.w;
        };
        // picked sites / positions / ranks: four 16-bit fields of one register pair each (indexed by shifts; arrays
        // indexed at run time would be local memory)
        unsigned long long pool = 0ull, ppos = 0ull, ranks = 0ull;
        int npool = 0, nr = 0;
        const int npk = ts.npick[tf_idx];
        for (int k = 0; k < npk; ++k) {
          const uint32_t pd = ts.pick[tf_idx][k];
          const int sl = (int)((pd >> 1) & 7u), code = (int)((pd >> 4) & 15u), d = (int)((pd >> 8) & 15u), p = (int)((pd >> 12) & 15u);
          const uint32_t word = next_word();
          if (pd & 0x10000u) { pool = 0ull; ppos = 0ull; npool = 0; }   // first pick of a sublattice
          if (!(pd & 1u)) {
            // a site that holds `code`: index among the remaining ones -> rank in the ascending-site list
            if (p == 0) { ranks = 0ull; nr = 0; }
            int idx = (int)mulhi32(word, (uint32_t)(count_of(d) - p));
            int at = 0;
            for (int q = 0; q < nr; ++q)
              if (idx >= (int)((ranks >> (16 * q)) & 0xffffull)) { ++idx; at = q + 1; }
            const unsigned long long low = (1ull << (16 * at)) - 1ull;
            ranks = (ranks & low) | ((unsigned long long)idx << (16 * at)) | ((ranks & ~low) << 16);
            ++nr;
            const int pp = tf_select(ts, planes, pfx, sl, code, idx, false);
            if (npool < LMC_MAX_FLIPS) {
              pool |= (unsigned long long)tf_site_of_pos(m, ts, sl, pp) << (16 * npool);
              ppos |= (unsigned long long)pp << (16 * npool);
              ++npool;
            }
          } else {
            // one of the vacated sites takes `code`
            const int idx = (int)mulhi32(word, (uint32_t)npool);
            const int site = (int)((pool >> (16 * idx)) & 0xffffull), pp = (int)((ppos >> (16 * idx)) & 0xffffull);
            const unsigned long long low = (1ull << (16 * idx)) - 1ull;
            pool = (pool & low) | ((pool >> 16) & ~low);
            ppos = (ppos & low) | ((ppos >> 16) & ~low);
            --npool;
            st.push(site, occ[site], code, sl, pp);
          }
        }
        st.log_priori = tfc[4 * TF + tf_idx];   // compute_log_priori_factor, mcusher.py:656-711 (tabulated per direction)
      }

      // ------------------------------ evaluate ------------------------------------------------
      // the groups took different paths through the proposal (swap / table flip, different numbers of picks): bring the
      // warp back together, or the record loops below run once per group
      __syncwarp();
      double dmu = 0.0, dEw = 0.0;
      if (m.muW) {
#pragma unroll
        for (int f = 0; f < MF; ++f)
          if (f < st.n) dmu += mu_of_s(m, ctab, st.site_of(f), st.new_of(f), st.sl_of(f)) - mu_of_s(m, ctab, st.site_of(f), st.old_of(f), st.sl_of(f));
      }
      // Ewald term: the cached potential of every changed site and the site-kernel elements between them are
      // loaded here (L2 / HBM) and consumed after the record loops
      double fl[MF], kx[MF * (MF - 1) / 2];
      if (EWF) {
#pragma unroll
        for (int f = 0; f < MF; ++f) {
          fl[f] = 0.0;
          if (f < st.n) fl[f] = fld[st.site_of(f)];
#pragma unroll
          for (int h = 0; h < f; ++h) {
            kx[f * (f - 1) / 2 + h] = 0.0;
            if (f < st.n) kx[f * (f - 1) / 2 + h] = __ldg(m.ewK + (size_t)st.site_of(h) * m.N + st.site_of(f));
          }
        }
      }
      // (a full-warp barrier in front of every record loop: the groups hold steps with two, three or four flips and
      // must walk each loop together)
      double acc = 0.0;
      uint32_t ps[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, pc[3] = {0u, 0u, 0u};
      __syncwarp();
      if (live && st.n > 0) acc = spec_flip_energy_n<0>(m, occ, dtab, st.site_of(0), st.old_of(0), st.new_of(0), l, ps, pc);
      ps[0] = (uint32_t)st.site_of(0); pc[0] = (uint32_t)st.new_of(0);
      __syncwarp();
      if (live && st.n > 1) acc += spec_flip_energy_n<1>(m, occ, dtab, st.site_of(1), st.old_of(1), st.new_of(1), l, ps, pc);
      ps[1] = (uint32_t)st.site_of(1); pc[1] = (uint32_t)st.new_of(1);
      __syncwarp();
      if (__any_sync(FULL, live && st.n > 2)) {
        if (live && st.n > 2) acc += spec_flip_energy_n<2>(m, occ, dtab, st.site_of(2), st.old_of(2), st.new_of(2), l, ps, pc);
        ps[2] = (uint32_t)st.site_of(2); pc[2] = (uint32_t)st.new_of(2);
        __syncwarp();
        if (live && st.n > 3) acc += spec_flip_energy_n<3>(m, occ, dtab, st.site_of(3), st.old_of(3), st.new_of(3), l, ps, pc);
        __syncwarp();
      }
      if (EWF) {
        // flip f sees the potential shifted by the step's earlier flips (ewald.py:168-181)
        double dq[MF];
#pragma unroll
        for (int f = 0; f < MF; ++f) dq[f] = 0.0;
#pragma unroll
        for (int f = 0; f < MF; ++f)
          if (f < st.n) {
            const double2 qn = ewald_qd_s(m, ctab, st.site_of(f), st.new_of(f), st.sl_of(f)), qo = ewald_qd_s(m, ctab, st.site_of(f), st.old_of(f), st.sl_of(f));
            dq[f] = qn.x - qo.x;
            double phi = fl[f];
#pragma unroll
            for (int h = 0; h < f; ++h) phi += dq[h] * kx[f * (f - 1) / 2 + h];
            dEw += 2.0 * dq[f] * phi + (qn.y - qo.y);
          }
      }
      acc += __shfl_xor_sync(FULL, acc, 1);
      acc += __shfl_xor_sync(FULL, acc, 2);
      double dH = acc;
      if (EWF) dH += nat_ew * dEw;
      if (m.muW) dH += nat_mu * dmu;

      // ------------------------------ accept (metropolis.py:31-49) ----------------------------
      const double exponent = __dadd_rn(__dmul_rn(-beta, dH), st.log_priori);
      const int af = accept_fast(exponent, lf);
      bool acc_ = af != 0;
      if (af < 0)
        acc_ = exponent > log(u01(philox4x32_10((uint32_t)mystep, (uint32_t)(mystep >> 32), 0u, wid, k0, k1).w));
      const uint32_t bal = __ballot_sync(FULL, acc_ && live);
      if (bal == 0u) {
        step += (unsigned long long)nb;
        it += nb;
        accepted = false;
        continue;
      }

      // ------------------------------ commit the first accepted step --------------------------
      const int src = __ffs(bal) - 1;
      const int j = src / SG;
      const int c_n = __shfl_sync(FULL, st.n, src);
      const bool c_tf = __shfl_sync(FULL, tf_idx, src) >= 0;
      const double c_dH = __shfl_sync(FULL, dH, src);
      const double c_dmu = __shfl_sync(FULL, dmu, src);
      const double c_dEw = __shfl_sync(FULL, dEw, src);
      PackedStep cs;
      cs.site = __shfl_sync(FULL, st.site, src); cs.pos = __shfl_sync(FULL, st.pos, src);
      cs.codes = __shfl_sync(FULL, st.codes, src); cs.sl = __shfl_sync(FULL, st.sl, src);
      cs.n = c_n; cs.log_priori = 0.0;
      int c_site[MF], c_old[MF], c_new[MF], c_sl[MF], c_pos[MF];
#pragma unroll
      for (int f = 0; f < MF; ++f) {
        c_site[f] = cs.site_of(f); c_old[f] = cs.old_of(f); c_new[f] = cs.new_of(f); c_sl[f] = cs.sl_of(f); c_pos[f] = cs.pos_of(f);
      }
      if (EWF && c_n > 0) {
        // the changed charges shift the potential cache (rows of K), as in the classic kernel
        double dq[MF];
#pragma unroll
        for (int f = 0; f < MF; ++f)
          dq[f] = f < c_n ? ewald_qd(m, c_site[f], c_new[f], c_sl[f]).x - ewald_qd(m, c_site[f], c_old[f], c_sl[f]).x : 0.0;
        if ((m.N & 1) == 0) {
          // two sites per lane and load: half as many L2 round trips (the rows are 16-byte aligned when N is even)
          double2* fld2 = reinterpret_cast<double2*>(fld);
          const double2* kr[MF];
#pragma unroll
          for (int f = 0; f < MF; ++f) kr[f] = reinterpret_cast<const double2*>(m.ewK + (size_t)c_site[f < c_n ? f : 0] * m.N);
#pragma unroll 4
          for (int k = g; k < (m.N >> 1); k += G) {
            double2 v = fld2[k];
#pragma unroll
            for (int f = 0; f < MF; ++f)
              if (f < c_n) {
                const double2 kv = __ldg(kr[f] + k);
                v.x += dq[f] * kv.x;
                v.y += dq[f] * kv.y;
              }
            fld2[k] = v;
          }
        } else {
#pragma unroll 4
          for (int k = g; k < m.N; k += G) {
            double v = fld[k];
#pragma unroll
            for (int f = 0; f < MF; ++f)
              if (f < c_n) v += dq[f] * __ldg(m.ewK + (size_t)c_site[f] * m.N + k);
            fld[k] = v;
          }
        }
        if (g == 0) feat[m.ewF] += c_dEw;
      }
      // classic record path flip by flip: per-record differences in the reference's cluster order (evaluator.pyx:253-263)
      // for the feature update; flip f is evaluated with flips < f written to the occupancy
#pragma unroll
      for (int f = 0; f < MF; ++f)
        if (f < c_n) {   // uniform
          const RecChunk pre = load_records<G>(m, c_site[f], g);
          (void)flip_energy<G, KONE>(m, t, occ, c_site[f], c_old[f], c_new[f], stash0, g, pre);
          __syncwarp();
          flip_features<G, KONE>(m, t, c_site[f], stash0, feat, g, load_segment<G>(m, c_site[f], g));
          __syncwarp();
          if (g == 0) {
            occ[c_site[f]] = (uint8_t)c_new[f];
            cnt[c_sl[f] * LMC_MAX_CODES + c_old[f]]--;
            cnt[c_sl[f] * LMC_MAX_CODES + c_new[f]]++;
            const int nw = m.sl_nwords[c_sl[f]];
            uint32_t* pl = planes + m.sl_plane_off[c_sl[f]] + (c_pos[f] >> 5);
            const uint32_t bit = 1u << (c_pos[f] & 31);
            pl[c_old[f] * nw] ^= bit;
            pl[c_new[f] * nw] ^= bit;
          }
          {   // words behind the flipped position: one set bit fewer in the old code's plane, one more in the new one's
            const int nw = m.sl_nwords[c_sl[f]];
            uint16_t* po = pfx + m.sl_plane_off[c_sl[f]] + c_old[f] * nw;
            uint16_t* pn = pfx + m.sl_plane_off[c_sl[f]] + c_new[f] * nw;
            for (int wd = (c_pos[f] >> 5) + 1 + g; wd < nw; wd += G) { po[wd] -= 1; pn[wd] += 1; }
          }
          __syncwarp();
        }
      if (g == 0 && m.muW && c_n > 0) feat[m.muF] += c_dmu;
      __syncwarp();
      if (c_tf) tfc_valid = false;   // the species counts changed
      if (c_n > 0) ring_valid = false;   // the stash slot of the commit is where the rings live
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
