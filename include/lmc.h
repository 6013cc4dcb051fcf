/*
 * liblmc -- B200-native lattice Monte-Carlo engine: C ABI.
 *
 * This is the drop-in boundary for smol's per-flip hot path.  smol has no FFI of its own for
 * this path: the native boundary in the reference is the set of Cython cpdef methods taking
 * typed memoryviews.  Each entry point below names the reference interface it replaces
 * (paths relative to the smol repository root):
 *
 *   lmc_model_create      <- ClusterExpansionProcessor.__init__ table build
 *                            (smol/moca/processor/expansion.py:120-156, 344-392),
 *                            OrbitContainer / IntArray2DContainer (smol/utils/cluster/container.pyx:21-341),
 *                            EwaldProcessor tables (smol/moca/processor/ewald.py:76-101),
 *                            ChemicalPotentialManager._build_table (smol/moca/ensemble.py:89-99),
 *                            Sublattice (smol/moca/sublattice.py:23-110)
 *   lmc_full_features     <- ClusterSpaceEvaluator.correlations_from_occupancy / interactions_from_occupancy
 *                            (smol/utils/cluster/evaluator.pyx:121-209), EwaldProcessor.compute_feature_vector
 *                            (smol/moca/processor/ewald.py:128-145), Ensemble.compute_feature_vector
 *                            (smol/moca/ensemble.py:323-351) -- batched over walkers
 *   lmc_delta_features    <- ClusterSpaceEvaluator.delta_correlations_from_occupancies /
 *                            delta_interactions_from_occupancies (evaluator.pyx:211-317),
 *                            delta_ewald_single_flip (smol/utils/cluster/ewald.pyx:9-59),
 *                            Ensemble.compute_feature_vector_change (ensemble.py:353-376) -- batched
 *   lmc_run               <- Sampler.sample inner loops (smol/moca/sampler/sampler.py:195-210, 436-440):
 *                            MCKernel.single_step (smol/moca/kernel/base.py:145-166),
 *                            Flip/Swap/TableFlip.propose_step (smol/moca/kernel/mcusher.py:154-200, 553-711),
 *                            Composite.propose_step over Flip / Swap sub-ushers (mcusher.py:307-394),
 *                            MultiStep.propose_step over a Flip / Swap sub-usher (mcusher.py:203-304),
 *                            Metropolis / WangLandau accept (kernel/metropolis.py:31-49,
 *                            kernel/wanglandau.py:186-266)
 *   lmc_ewald_field       <- the site sums of delta_ewald_single_flip (smol/utils/cluster/ewald.pyx:43-58),
 *                            evaluated once per walker and then kept current by lmc_run (potential cache)
 *   lmc_bias_init, LmcRunConfig.bias_* <- MCBias.compute_bias / compute_bias_change of FugacityBias and
 *                            SquareChargeBias (smol/moca/kernel/bias.py:79-287) and the bias term of the
 *                            Metropolis exponent (kernel/metropolis.py:43-44)
 *   lmc_ewald_site_kernel <- the reciprocal + real space sums of pymatgen's EwaldSummation behind
 *                            EwaldTerm.get_ewald_matrix (smol/cofe/extern/ewald.py:102-177): pair kernel between
 *                            a few origin sites and every site of the supercell (the matrix follows by
 *                            translation and by the charge products)
 *   lmc_distance_init, LmcRunConfig.dist_* <- DistanceProcessor.compute_feature_vector[_change]
 *                            (smol/moca/processor/distance.py:133-180, 281-331, 424-472) over
 *                            ClusterSpaceEvaluator.corr_distances / interaction_distances_from_occupancies
 *                            (smol/utils/cluster/evaluator.pyx:319-435)
 *   LmcRunConfig.walker_mask_dev / accept_offset_dev <- MulticellKernel.single_step / _compute_step_trace
 *                            (smol/moca/kernel/base.py:612-622, 645-692; MulticellMetropolis,
 *                            kernel/metropolis.py:102-175): one lmc_run per supercell shape over the walkers
 *                            sitting in it, hops judged by H_k'(new) - H_current
 *   lmc_cast_*            <- the int32 occupancy dtype contract (sampler.py:406)
 *
 * Conventions: every function returns 0 on success, <0 on error (message via lmc_last_error);
 * nothing throws across the ABI; no torch types.  Pointers named *_dev are DEVICE pointers owned
 * by the caller (PyTorch tensors in the Python host); model tables are HOST pointers copied to
 * device memory owned by the handle.  `stream` is a cudaStream_t passed as void*; launches are
 * asynchronous on it.  A handle is bound to the device current at creation, one host thread per
 * handle.
 */
#ifndef LMC_H_
#define LMC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMC_ABI_VERSION 16
#define LMC_MAX_CLUSTER_SITES 4 /* sites per cluster (record = 3 other sites + class) */
#define LMC_MAX_SUBLATTICES 8
#define LMC_MAX_CODES 8       /* species codes per sublattice */
#define LMC_MAX_FLIPS 4       /* changed sites per attempted step */
#define LMC_MAX_DIMS 16       /* species counts tracked by the table-flip usher */
#define LMC_MAX_TABLE_FLIPS 8
#define LMC_MAX_COMPOSITE 4   /* sub-ushers of a composite usher */
#define LMC_MAX_BIAS_ROWS 4   /* hyperplanes of a SquareHyperplaneBias */

typedef struct LmcModel LmcModel; /* opaque */

/* All arrays are host pointers, C-contiguous. */
typedef struct LmcModelDesc {
  uint32_t abi_version;
  int32_t num_sites;        /* N */
  int32_t num_features;     /* F = len(natural_parameters) */
  int32_t num_ce_features;  /* features produced by the cluster part (corr functions or orbits) */
  int32_t supercell_size;   /* number of primitive cells (Processor.size) */
  double feature0;          /* features[0]: size (CE) or offset*size (decomposition) */
  const double* natural_parameters; /* [F] */

  /* orbit tables */
  int32_t num_orbits;            /* orbits with clusters (empty cluster excluded) */
  const int32_t* orb_tab_off;    /* [num_orbits] offset of the orbit's K*T block in ftab */
  const int32_t* orb_tab_len;    /* [num_orbits] T = flattened tensor length */
  const int32_t* orb_nfunc;      /* [num_orbits] K = functions (bit combos); 1 for decomposition */
  const int32_t* orb_fidx;       /* [num_orbits] first feature index (bit_id or orbit id) */
  const int32_t* orb_csize;      /* [num_orbits] sites per cluster */
  const int32_t* orb_stride;     /* [num_orbits][LMC_MAX_CLUSTER_SITES] flat tensor strides */
  const double* orb_weight;      /* [num_orbits] size / (total cluster rows of the orbit) */
  const double* ftab;            /* [ftab_len] feature tensors, orbit-major then function-major */
  int64_t ftab_len;

  /* full evaluation rows (clusterspace.get_orbit_indices layout, padded to 4 sites) */
  const int64_t* orb_row_off;    /* [num_orbits+1] */
  const uint16_t* full_rows;     /* [rows][4] site indices; unused slots repeat site 0 with stride 0 */

  /* per-site local records (rows containing the site), sorted by orbit */
  int32_t num_classes;
  const int32_t* cls_orbit;      /* [num_classes] orbit index of the class */
  const int32_t* cls_stride;     /* [num_classes][4]: strides of the 3 other sites, self stride */
  const int64_t* site_rec_off;   /* [N+1] */
  const uint16_t* site_rec;      /* [records][4]: 3 other site indices, class id */
  const int64_t* site_seg_off;   /* [N+1] */
  const int32_t* site_seg;       /* [segments][3]: first record (site relative), count, orbit */

  /* Ewald (optional: ewald_size == 0 disables) */
  int32_t ewald_size;            /* E */
  int32_t ewald_width;           /* columns of ewald_inds */
  const double* ewald_matrix;    /* [E][E] */
  const int32_t* ewald_inds;     /* [N][ewald_width], -1 = vacancy */
  int32_t ewald_feature;         /* feature index of the Ewald term */

  /* chemical potentials (optional: mu_width == 0 disables) */
  int32_t mu_width;
  const double* mu_table;        /* [N][mu_width] */
  int32_t mu_feature;            /* feature index of the chemical work (natural parameter -1) */

  /* active sublattices */
  int32_t num_sublattices;
  const int32_t* sl_site_off;    /* [num_sublattices+1] offsets into sl_sites */
  const int32_t* sl_sites;       /* active site indices, in Sublattice.active_sites order */
  const int32_t* sl_ncodes;      /* [num_sublattices] */
  const int32_t* sl_codes;       /* [num_sublattices][LMC_MAX_CODES] encoding */
  const double* sl_prob;         /* [num_sublattices] proposal probabilities */

  /* table-flip usher (optional: tf_num_flips == 0 disables) */
  int32_t tf_num_dims;           /* d: one per (sublattice, species) over ALL sublattices */
  int32_t tf_num_flips;          /* rows of the flip table */
  const int32_t* tf_table;       /* [tf_num_flips][tf_num_dims] "counts" format */
  const double* tf_weights;      /* [2*tf_num_flips] forward/backward weights */
  const int32_t* tf_max_n;       /* [tf_num_dims] active sites of the dim's sublattice */
  const int32_t* tf_dim_sl;      /* [tf_num_dims] ACTIVE sublattice index of the dim, -1 if inactive */
  const int32_t* tf_dim_code;    /* [tf_num_dims] species code of the dim */
  double tf_swap_weight;
} LmcModelDesc;

enum { LMC_USHER_FLIP = 0, LMC_USHER_SWAP = 1, LMC_USHER_TABLEFLIP = 2, LMC_USHER_COMPOSITE = 3, LMC_USHER_MULTISTEP = 4 };
enum { LMC_KERNEL_METROPOLIS = 0, LMC_KERNEL_WANGLANDAU = 1 };
/* bias terms of the Metropolis kernel: value = sum_k table[k][occ[k]] (LMC_BIAS_TABLE_SUM; FugacityBias with
 * table = log fugacity fractions) or -penalty * sum_r (sum_k table[k][occ[k]][r] - intercept[r])^2
 * (LMC_BIAS_SQUARE_SUM; SquareChargeBias: one row of oxidation states, intercept 0; SquareHyperplaneBias: row r =
 * column A[r][dim(k, code)] of the hyperplane normals, intercept b[r]) */
enum { LMC_BIAS_NONE = 0, LMC_BIAS_TABLE_SUM = 1, LMC_BIAS_SQUARE_SUM = 2 };

typedef struct LmcWangLandau {
  double min_enthalpy, max_enthalpy, bin_size, flatness, mod_update;
  int32_t num_bins, check_period, update_period;
  int32_t reserved; /* 1: mean_features_dev holds per-bin SUMS (valid when update_period == 1) */
  /* per-walker state, device pointers */
  double* entropy_dev;        /* [W][num_bins] */
  int64_t* histogram_dev;     /* [W][num_bins] */
  int64_t* occurrences_dev;   /* [W][num_bins] */
  double* mean_features_dev;  /* [W][num_bins][F] */
  double* mod_factor_dev;     /* [W] */
  int64_t* steps_counter_dev; /* [W] valid-state counter (wanglandau.py:230-232) */
  /* per-sample traces of the state above (wanglandau.py:247-251), each may be NULL */
  double* trace_entropy_dev;        /* [S][W][num_bins] */
  int64_t* trace_histogram_dev;     /* [S][W][num_bins] */
  int64_t* trace_occurrences_dev;   /* [S][W][num_bins] */
  double* trace_mean_features_dev;  /* [S][W][num_bins][F] cumulative MEAN features (sums / occurrences if reserved == 1) */
  double* trace_mod_factor_dev;     /* [S][W] modification factor before the flatness check of the sampled step */
  /* optional successor table of the modification factor (a callable `mod_update`, wanglandau.py:100-105, tabulated by
   * the host: entry k + 1 = mod_update(entry k), entry 0 = the initial factor).  A flatness event replaces a factor
   * equal to entry k by entry k + 1 (the last entry maps to itself); NULL: factor / mod_update */
  const double* mod_table_dev;      /* [mod_table_len] */
  int32_t mod_table_len;
  int32_t reserved2;
} LmcWangLandau;

typedef struct LmcRunConfig {
  int32_t num_walkers;        /* W walkers resident on this device */
  int32_t walker_id_base;     /* global id of walker 0 (RNG counter word 3): rank sharding */
  int32_t usher;              /* LMC_USHER_* */
  int32_t kernel;             /* LMC_KERNEL_* */
  int64_t num_samples;        /* S sampling intervals */
  int32_t thin_by;            /* steps per interval */
  int32_t group_size;         /* lanes cooperating on one walker: 0 = auto, else 1..32 (power of 2) */
  int32_t block_threads;      /* 0 = auto */
  int32_t spec_mode;          /* Metropolis flip/swap kernel: 0 = auto (by the acceptance the library has seen so far,
                                 refreshed asynchronously: the choice may differ between identical runs -- the two kernels
                                 sum dH in different orders), 1 = classic (one step per warp), 2 = speculative batch
                                 (8 steps per warp, first accept wins; error if unsupported), 3 = speculative where the
                                 model / run supports it, else classic (deterministic; what the Python host passes) */
  uint64_t step_begin;        /* global index of the first step (RNG counter words 0,1) */
  const uint64_t* seeds_dev;  /* [W] Philox key per walker */
  const double* beta_dev;     /* [W] 1/(kB T); ignored by Wang-Landau */
  /* state, in/out */
  int8_t* occ_dev;            /* [W][row_stride] int8 codes, row_stride = lmc_row_stride(N) */
  double* features_dev;       /* [W][F] running features */
  double* enthalpy_dev;       /* [W] running enthalpy */
  /* per-sample traces, each may be NULL */
  int8_t* trace_occ_dev;      /* [S][W][N] */
  double* trace_features_dev; /* [S][W][F] */
  double* trace_enthalpy_dev; /* [S][W] */
  uint8_t* trace_accepted_dev;/* [S][W] flag of the LAST step of the interval (sampler.py:199-201) */
  int32_t* trace_naccepted_dev;/* [S][W] accepted steps in the interval (engine extension) */
  /* optional Ewald potential cache, in/out: field[w][k] = sum_j q_j K[k][j] of walker w's CURRENT occupancy
     (lmc_ewald_field).  Non-NULL: a flip costs O(1) reads of it and every accepted step updates it;
     NULL: every flip gathers its Ewald matrix rows.  Needs a factorisable Ewald matrix (lmc_model_info). */
  double* ewald_field_dev;    /* [W][N] */
  /* optional bias term (Metropolis only), see LMC_BIAS_*; state from lmc_bias_init, kept current by lmc_run */
  int32_t bias_mode;
  int32_t bias_width;            /* species codes per site of bias_table_dev */
  int32_t bias_rows;             /* rows per (site, code): 1, or the number of hyperplanes (<= LMC_MAX_BIAS_ROWS) */
  double bias_penalty;
  const double* bias_table_dev;  /* [N][bias_width][bias_rows] */
  double* bias_dev;              /* [W] running bias value, in/out */
  double* bias_sum_dev;          /* [W][bias_rows] running table sums minus intercepts, in/out */
  double* trace_bias_dev;        /* [S][W], may be NULL */
  /* optional distance processor (processor/distance.py): the features are [L, |f_i - target_i| ...] with f the
     correlation / cluster-interaction vector per supercell and L the largest orbit diameter up to which all
     features match the target within dist_tol; natural parameters [-w, W ...].  Metropolis flip / swap steps
     without Ewald term.  State from lmc_distance_init, kept current by lmc_run. */
  int32_t dist_mode;             /* 0 off, 1 on */
  int32_t dist_num_groups;       /* orbit groups of equal diameter, ascending (clusterspace.py:367-381) */
  double dist_tol;
  const double* dist_target_dev; /* [F] */
  const int32_t* dist_group_off_dev;  /* [dist_num_groups + 1] offsets into dist_group_idx_dev */
  const int32_t* dist_group_idx_dev;  /* feature indices of the groups' orbits */
  const double* dist_group_diam_dev;  /* [dist_num_groups] */
  double* dist_vector_dev;       /* [W][F] running correlation / interaction vector per supercell, in/out */
  /* LMC_USHER_COMPOSITE (mcusher.py:307-394): every step picks one sub-usher by weight (random word 4 of the
     step), which proposes with its OWN sublattice probabilities (0 = sublattice not served by it) */
  int32_t comp_num;                                              /* 1..LMC_MAX_COMPOSITE */
  int32_t comp_usher[LMC_MAX_COMPOSITE];                         /* LMC_USHER_FLIP or LMC_USHER_SWAP */
  double comp_cum[LMC_MAX_COMPOSITE];                            /* cumulative pick probabilities */
  double comp_sl_cum[LMC_MAX_COMPOSITE][LMC_MAX_SUBLATTICES];    /* cumulative sublattice probabilities */
  /* LMC_USHER_MULTISTEP (mcusher.py:203-304): a step chains step_length proposals of one sub-usher, each against
     the occupancy with the earlier ones applied; a proposal touching an already changed site is dropped.  The
     length is picked with random word 4 of the step, proposal j draws from block 2 + j.  At most LMC_MAX_FLIPS
     changed sites: lengths <= 4 (Flip) or <= 2 (Swap) */
  int32_t ms_usher;                                              /* LMC_USHER_FLIP or LMC_USHER_SWAP */
  int32_t ms_num;                                                /* 1..LMC_MAX_COMPOSITE step lengths */
  int32_t ms_len[LMC_MAX_COMPOSITE];
  double ms_cum[LMC_MAX_COMPOSITE];                              /* cumulative probabilities of the lengths */
  LmcWangLandau wl;           /* used when kernel == LMC_KERNEL_WANGLANDAU */
  /* Multicell sampling (MulticellKernel, kernel/base.py:439-722: one chain hops between supercell shapes, each with
     its own occupancy; the host keeps one state array per shape and drives them with these two):
     walker_mask_dev: only walkers with a non-zero byte take part in this call (the others keep their state and
     their trace rows are not written); accept_offset_dev: added to the enthalpy change inside the Metropolis
     exponent only -- a hop into shape k' is one step of k' judged by H_k'(after) - H_current
     = dH_step + (H_k' - H_current) (base.py:612-622).  Classic Metropolis kernels; both may be NULL. */
  const uint8_t* walker_mask_dev;   /* [W] */
  const double* accept_offset_dev;  /* [W] */
  /* optional workspace of the speculative kernel (ABI 16): W x lmc_model_info()[4] bytes.  Non-NULL (and info[4] > 0):
     every active site keeps the species codes its merged records gather as packed "environment words" here, rebuilt
     from occ_dev at the start of every call and kept current by accepted steps, so a rejected step reads one word per
     flip and lane instead of three occupancy bytes per record.  Same chains, same arithmetic; NULL: gathers. */
  void* spec_env_dev;
} LmcRunConfig;

int lmc_version(void);
const char* lmc_last_error(void);
int lmc_row_stride(int num_sites); /* bytes per walker row of occ_dev (16-byte multiple, at least one zero pad byte) */

int lmc_model_create(const LmcModelDesc* desc, LmcModel** out);
int lmc_model_destroy(LmcModel* model);
int lmc_model_num_features(const LmcModel* model);
/* info[0] = 1 if the Ewald matrix factorises as M[i,j] = q_i q_j K[site_i, site_j] (potential cache usable),
 * info[1] = speculative-kernel tables built, info[2] = bytes of the staged table blob, info[3] = records per site
 * of the speculative kernel, info[4] = bytes per walker of LmcRunConfig.spec_env_dev (0: environment words not
 * available for this model), info[5] = bits per site of the compact environment words the speculative flip / swap kernel
 * keeps in shared memory (0: the records do not fit 64 bits, occupancy gathers); entries beyond `n` are not written */
int lmc_model_info(const LmcModel* model, int32_t* info, int n);

/* int32 [W][N] <-> int8 [W][row_stride] */
int lmc_cast_i32_to_i8(const int32_t* src_dev, int8_t* dst_dev, int num_walkers, int num_sites, void* stream);
int lmc_cast_i8_to_i32(const int8_t* src_dev, int32_t* dst_dev, int64_t num_rows, int num_sites,
                       int src_row_stride, void* stream);

/* features_dev [W][F] <- full evaluation of every walker's occupancy; enthalpy_dev [W] may be NULL */
int lmc_full_features(const LmcModel* model, const int8_t* occ_dev, int num_walkers, double* features_dev,
                      double* enthalpy_dev, void* stream);

/* The same with the Ewald term evaluated through the walkers' potential: field_dev [W][N] (scratch owned by the caller) is
 * filled as by lmc_ewald_field and the Ewald feature follows as sum_k q_k field[k] + sum_k M[e_k, e_k] -- O(W N^2) as one
 * tiled product instead of a pair sum per walker; on return field_dev IS the potential cache of the occupancies.
 * field_dev == NULL or a matrix that does not factorise (lmc_model_info): plain lmc_full_features. */
int lmc_full_features_field(const LmcModel* model, const int8_t* occ_dev, int num_walkers, double* features_dev,
                            double* enthalpy_dev, double* field_dev, void* stream);

/* field_dev [W][N] <- Ewald potential cache of every walker's occupancy (see LmcRunConfig.ewald_field_dev) */
int lmc_ewald_field(const LmcModel* model, const int8_t* occ_dev, int num_walkers, double* field_dev, void* stream);

/* bias_dev [W], sum_dev [W][bias_rows] <- bias value and table sums (minus the host array intercepts[bias_rows],
 * NULL = zeros) of every walker's occupancy (MCBias.compute_bias) */
int lmc_bias_init(const int8_t* occ_dev, int num_walkers, int num_sites, int bias_mode, int bias_width, int bias_rows,
                  double bias_penalty, const double* intercepts, const double* bias_table_dev, double* bias_dev,
                  double* sum_dev, void* stream);

/* out_dev [W][F] <- feature change of walker w for its k flips (sites/codes [W][k] int32, applied
 * sequentially, chemical work against the pre-step occupancy) */
int lmc_delta_features(const LmcModel* model, const int8_t* occ_dev, int num_walkers, const int32_t* sites_dev,
                       const int32_t* codes_dev, int num_flips, double* out_dev, void* stream);

/* advance every walker by num_samples*thin_by attempted steps */
int lmc_run(const LmcModel* model, const LmcRunConfig* cfg, void* stream);

/* host-only (no device needed): tables of the speculative-batch kernel for a model description.
 * info[8] = {built, code radix NC, entries per new-code plane, records per site, blocks, merged, table bytes,
 * record bytes}; dtab_out [NC][L] doubles and rec_out [N][records] x 8 bytes are filled when large enough */
int lmc_spec_tables_host(const LmcModelDesc* desc, int32_t* info, double* dtab_out, int64_t dtab_cap,
                         uint8_t* rec_out, int64_t rec_cap);
/* host-only: environment-word tables of the speculative kernel (LmcRunConfig.spec_env_dev).  info[8] = {built, bits
 * per code b, records per lane, padded records per lane P, wide (64-bit lane chunks), active sites A, reverse entries
 * per site R, pair table built}; tb_out [N][4][P] u16 table bases in lane order, rev_out [A][R] u32 (gathering
 * active site | bit << 16, 0xffffffff = none), pair_out [A][A][4] x (u32, or u64 when wide) slot masks of the column
 * site inside the row site's words; each filled when large enough (capacities in elements / bytes for pair_out) */
int lmc_spec_env_host(const LmcModelDesc* desc, int32_t* info, uint16_t* tb_out, int64_t tb_cap, uint32_t* rev_out,
                      int64_t rev_cap, uint8_t* pair_out, int64_t pair_cap);

/* host-only: compact environment words (one 64-bit word per active site and walker in shared memory; the kernel the
 * library takes for Metropolis flip / swap steps of models whose merged records fit 64 bits per site).  info[8] = {built,
 * bits per code, records per lane, padded records per lane P, active sites A, reverse entries per site R, site classes C,
 * bits used}; desc_out [C][4][P] u32 (table base | field shift << 16, lane order), cls_out [N] u8, rev_out [A][R] u32
 * (gathering active site | bit << 16, 0xffffffff = none), pair_out [A][A] u64 (bits of the row site's word that hold the
 * column site); each filled when large enough (capacities in elements) */
int lmc_spec_c64_host(const LmcModelDesc* desc, int32_t* info, uint32_t* desc_out, int64_t desc_cap, uint8_t* cls_out,
                      int64_t cls_cap, uint32_t* rev_out, int64_t rev_cap, uint64_t* pair_out, int64_t pair_cap);

/* Distance processor state of every walker: features_dev [W][F] holds the EXTENSIVE features on entry
 * (lmc_full_features) and the distance vector on return; vector_dev [W][F] <- features / supercell size;
 * enthalpy_dev [W] <- natural_parameters . distance vector */
int lmc_distance_init(const LmcModel* model, int num_walkers, double* features_dev, double* vector_dev,
                      double* enthalpy_dev, const double* target_dev, double tol, int num_groups,
                      const int32_t* group_off_dev, const int32_t* group_idx_dev, const double* group_diam_dev,
                      void* stream);

/* Ewald pair kernel (no charges, no self term):
 *   out_dev[o][k] = (2 pi / V) sum_G coef[G] cos(G . (r_k - r_origin[o])) + 1/2 sum_T' erfc(sqrt(eta) |d + T|) / |d + T|
 * with d = r_k - r_origin[o], the second sum over the lattice translations T with 1e-8 < |d + T| <= real_cut.
 * cart_dev [num_sites][3], origins_dev [num_origins] (site indices), gvec_dev [num_g][3], gcoef_dev [num_g]
 * (= exp(-G^2 / 4 eta) / G^2), tvec_dev [num_t][3]; all device pointers, doubles. */
int lmc_ewald_site_kernel(const double* cart_dev, int num_sites, const int32_t* origins_dev, int num_origins,
                          const double* gvec_dev, const double* gcoef_dev, int num_g, const double* tvec_dev, int num_t,
                          double eta, double real_cut, double volume, double* out_dev, void* stream);

/* number of kernel launches issued by this library since load (for bench accounting) */
int64_t lmc_launch_count(void);
/* how many of them were environment-word variants of the speculative kernel (LmcRunConfig.spec_env_dev) */
int64_t lmc_env_launch_count(void);
/* ... and how many were launches of the compact-word kernel (lmc_spec_c64_host) */
int64_t lmc_c64_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LMC_H_ */
