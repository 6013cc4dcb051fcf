// Device-resident model description shared by all kernels (passed by value as a kernel parameter).
#pragma once
#include <stdint.h>
#include "../../include/lmc.h"

namespace lmc {

// one orbit of clusters: where its tensors live and how to fold them into features
#define LMC_SEG_PIECE 6   // clusters per piece of an orbit segment (flip_features)

struct OrbDev {
  int ftab_off;    // offset of the K*T block in ftab (doubles)
  int atab_off;    // offset of the T block in the phase-A table (raw tensor if K==1, else contracted)
  int T, K, fidx;  // tensor length, functions, first feature index
  int csize;       // sites per cluster
  int row_off, row_cnt;  // rows in full_rows
  int stride[4];
  double w;        // size / total rows  (== 1 / multiplicity)
};

struct DevModel {
  int N, Npad, F, Fce, nOrb, nCls, nSl, size, kone;
  int Rstride;     // records per site in the padded table (multiple of 4; pads use class nCls)
  int tabA_len;    // doubles in the phase-A table
  double feature0;
  // shared-memory blob (staged once per block by TMA): [cls uint4 x (nCls + 1)][tabA double x tabA_len]
  // [nat double x F][orb OrbDev x nOrb][qtab][dtab: speculative kernel only]
  const unsigned char* blob;
  int blob_bytes, off_tabA, off_nat, off_orb, off_qtab;
  const double* ftab;      // global copy of all feature tensors (phase B when K > 1, full evaluation)
  const uint2* site_rec;   // [N][Rstride] (idx0 | idx1<<16, idx2 | cls<<16), padded per site
  const int4* site_seg;    // [N][Sstride] (first, count, orbit, pieces that follow | -1); count 0 = padding
  int Sstride;             // segment entries per site (multiple of 4; long segments are cut into pieces)
  int Kmax;                // largest number of functions of one orbit
  const uint2* full_rows;  // [rows] 4 x u16 site indices
  // Ewald
  int E, ewW, ewF;
  const double* ewMt;      // [E][E] TRANSPOSED matrix: ewMt[a*E + i] == M[i][a]
  const int* ewInds;       // [N][ewW]
  // factorised form (when M[i,j] == q_i q_j K[site_i, site_j] off the site-diagonal blocks):
  const double* ewK;       // [N][N] site kernel, nullptr when the matrix does not factorise
  const double* ewQ;       // [E] relative charge of a matrix row
  const double* ewD;       // [E] diagonal M[i,i]
  const uint8_t* ewQidx;   // [E] index of the row's charge in the shared-memory charge table
  int ewNQ;                // number of distinct charges; index ewNQ is the 0.0 used for vacancies
  const double2* ewQD;     // [N][ewW] (charge, diagonal M[e,e]) of species code c on site k; (0, 0) for vacancies
  // chemical potentials
  int muW, muF;
  const double* mu;        // [N][muW]
  // per-sublattice forms of the two per-site tables above, valid when every active site of a sublattice carries the
  // same row (what ChemicalPotentialManager / an Ewald summation produce): the proposal knows the sublattice, so the
  // lookups come from the parameter bank instead of L2
  int muC, qdC;
  int off_ctab;            // blob: [mu][charge][diagonal] x [LMC_MAX_SUBLATTICES][LMC_MAX_CODES] doubles (speculative kernels)
  double mu_c[LMC_MAX_SUBLATTICES][LMC_MAX_CODES];
  double qc_c[LMC_MAX_SUBLATTICES][LMC_MAX_CODES], qg_c[LMC_MAX_SUBLATTICES][LMC_MAX_CODES];   // charge, diagonal (two 8-byte tables: a 16-byte element was copied to a misaligned local frame)
  // active sublattices
  int sl_off[LMC_MAX_SUBLATTICES + 1];
  int sl_ncodes[LMC_MAX_SUBLATTICES];
  int sl_codes[LMC_MAX_SUBLATTICES][LMC_MAX_CODES];
  unsigned char sl_code_pos[LMC_MAX_SUBLATTICES][LMC_MAX_CODES];   // index of a code in sl_codes (ncodes if absent)
  int sl_first[LMC_MAX_SUBLATTICES];   // first site if the active sites are one contiguous range, else -1
  double sl_cum[LMC_MAX_SUBLATTICES];  // cumulative proposal probabilities
  int sl_nwords[LMC_MAX_SUBLATTICES];  // 32-bit words of one species bit-plane (ceil(n_active/32))
  int sl_plane_off[LMC_MAX_SUBLATTICES];  // word offset of the sublattice's planes [code][word]
  int plane_words;                     // total words of all planes of one walker
  int sl_nplanes[LMC_MAX_SUBLATTICES];    // planes of the sublattice (max code + 1)
  int sl_list_off[LMC_MAX_SUBLATTICES];   // u16 offset of the sublattice's position lists [code][n_active] (lmc_spec.cuh)
  int list_entries;                    // u16 entries of all position lists of one walker
  const int* sl_sites;
  // table flips
  int tfD, tfNF;
  int tf_table[LMC_MAX_TABLE_FLIPS][LMC_MAX_DIMS];
  double tf_w[2 * LMC_MAX_TABLE_FLIPS];
  int tf_max_n[LMC_MAX_DIMS];
  int tf_dim_sl[LMC_MAX_DIMS];
  int tf_dim_code[LMC_MAX_DIMS];
  double tf_sw;
  // the site picks of every flip direction in the order TableFlip.propose_step draws them (mcusher.py:602-639: per
  // sublattice the sites to vacate, dimension by dimension, then the species that take them): one descriptor per
  // random word -- kind (bit 0: 0 = pick a site holding `code`, 1 = hand a picked site to `code`) | sublattice << 1
  // | code << 4 | dimension << 8 | pick index inside its dimension << 12 | first pick of a sublattice << 16
  // (speculative table-flip kernel: every step of a batch walks ONE loop over its descriptors)
  int tf_npick[2 * LMC_MAX_TABLE_FLIPS];
  uint32_t tf_pick[2 * LMC_MAX_TABLE_FLIPS][2 * LMC_MAX_FLIPS];
  const double* lgam;      // [max_n + 2] ln(n!) table for the table-flip a-priori factor
  // speculative-batch kernel (lmc_spec.cuh): merged three-gather records + pre-differenced tables
  int spOK;                // tables built (0: model outside the limits of the speculative kernel)
  int spNC;                // code radix of the difference-table index (max species per site)
  int spL;                 // doubles per new-code plane of the difference table
  int spNQ;                // merged records per site (padded to a multiple of 8 with zero records)
  int spSb;                // bytes per site of sp_rec (8 spNQ)
  int off_dtab;            // offset of the difference table [spNC][spL] in the blob
  const unsigned char* sp_rec;  // [N][spNQ] x uint2 (s0 | s1<<16, s2 | tbase<<16)
  const double* spFtab;    // [spNC][spL][F] per-feature form of the difference table (cluster decomposition), or nullptr
  // environment words (ENV variants of the speculative kernel, see build_env_tables in lmc_api.cu)
  int envOK;               // tables built
  int envB;                // bits per species code (1 or 2); a record's field is 3 envB bits
  int envNRL, envNRLP;     // records per lane of the 4-lane step group (spNQ / 4), rounded up to a multiple of 8
  int envWide;             // lane chunk: 0 = one 32-bit word (4 words per site), 1 = one 64-bit word (8 words per site)
  int envNA;               // active sites (index = sl_off[sublattice] + position)
  int envRV;               // reverse-map entries per active site (multiple of 32)
  int envNCls;             // classes of sites with equal table-base lists, staged to shared memory (0: read envTb)
  int off_envtb, off_envcls;   // blob, behind blob_bytes: [envNCls][4][envNRLP] u16 lists, [N] u8 class of a site
  int blob_env_bytes;          // blob size including them (what the environment-word variants stage)
  const uint16_t* envTb;   // [N][4][envNRLP] table base of the lane's i-th record
  const uint32_t* envRev;  // [envNA][envRV] gathering site (active index) | bit position << 16; 0xffffffff = none
  // compact environment words (lmc_spec_c64.cuh, build_c64_tables in lmc_api.cu): one 64-bit word per active site and
  // walker in SHARED memory
  int c64OK, c64B, c64NRL, c64NRLP, c64NA, c64RV, c64NCls, c64Bits;
  int off_c64desc, off_c64cls;   // blob, behind blob_bytes: [c64NCls][4][c64NRLP] u32 (table base | shift << 16), [N] u8 class of a site
  int blob_c64_bytes;            // blob size including them (what the compact-word kernel stages)
  const uint32_t* c64Rev;        // [c64NA][c64RV] gathering site (active index) | bit << 16; 0xffffffff = none
  const unsigned long long* c64Pair;   // [c64NA][c64NA] bits of the row site's word that hold the column site
  const unsigned char* envPair;   // [envNA][envNA][4] x (u32 | u64 if envWide): lowest slot bits, per lane chunk, where the column
                                  // site sits among the codes the row site gathers; nullptr when not built (swaps need it)
};

struct RunArgs {
  int W, walker_base, usher, kernel;
  long long S;
  int thin;
  unsigned long long step0;
  const unsigned long long* seeds;
  const double* beta;
  int8_t* occ;
  double* features;
  double* enthalpy;
  int8_t* tr_occ;
  double* tr_feat;
  double* tr_enth;
  uint8_t* tr_acc;
  int* tr_nacc;
  double* ew_field;   // [W][N] Ewald potential cache (nullptr: gather the matrix rows), see lmc.h
  int bias_mode, bias_w, bias_rows;   // LMC_BIAS_*, codes and rows per code of bias_tab
  double bias_pen;
  const double* bias_tab;      // [N][bias_w][bias_rows]
  double* bias;                // [W] running bias value
  double* bias_sum;            // [W][bias_rows] running table sums
  double* tr_bias;             // [S][W]
  // distance processor (lmc.h): target vector, match tolerance, orbit groups by diameter, running vector
  int dist_ngrp;
  double dist_tol;
  const double* dist_target;
  const int* dist_grp_off;
  const int* dist_grp_idx;
  const double* dist_grp_diam;
  double* dist_vec;            // [W][F]
  int off_dist;                // shared memory: [vector F][delta F][new distances F] doubles
  int comp_num, comp_usher[LMC_MAX_COMPOSITE];                   // composite usher, see lmc.h
  double comp_cum[LMC_MAX_COMPOSITE];
  double comp_sl_cum[LMC_MAX_COMPOSITE][LMC_MAX_SUBLATTICES];
  int ms_usher, ms_num, ms_len[LMC_MAX_COMPOSITE];               // multi-step usher, see lmc.h
  double ms_cum[LMC_MAX_COMPOSITE];
  LmcWangLandau wl;
  int wpb;            // walkers per block
  int walker_smem;    // bytes of shared memory per walker
  int off_feat, off_stash, off_cnt, off_plane, off_ring, off_eidx, off_lists, off_bias;  // offsets inside a walker's shared-memory slab
  int off_tfc;        // table-flip usher: [weights][cumulative probabilities][log a-priori factors][sum] of the current counts
  int off_pref;       // Wang-Landau flips: records and segment entries of the next step (asynchronous prefetch)
  int off_wl;         // Wang-Landau: [entropy nb][histogram nb] of the walker in its slab, -1 = kept in global memory
  unsigned long long* stats;  // [2] accepted / attempted step totals (device counters; kernel selection feedback)
  const uint8_t* mask;        // [W] multicell: walkers taking part in this launch (nullptr = all)
  const double* acc_off;      // [W] multicell: enthalpy offset inside the Metropolis exponent (nullptr = 0)
  uint32_t* env;      // [W][envNA][4 or 8] environment words of the walkers (ENV variants), workspace rebuilt by every launch
  int off_env64;      // compact environment words in the walker's slab: [c64NA] x u64
  int max_flips;      // flips per step of the selected usher (stash slots)
  int seq_flips;      // debug: evaluate the flips of a step strictly one after another
};

}  // namespace lmc
