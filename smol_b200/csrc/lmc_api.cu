// C ABI of liblmc (see include/lmc.h): model upload, launch configuration, kernel dispatch.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>

#include <array>
#include <algorithm>
#include <atomic>
#include <map>
#include <string>
#include <vector>

#define LMC_API_TU
#include "lmc_kernels.cuh"
#include "lmc_launch.h"

using namespace lmc;

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_c64_launches{0};   // launches of the compact-word kernel
static std::atomic<long long> g_env_launches{0};   // launches of the environment-word variants (tests / diagnostics)

static int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}
#define CK(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct LmcModel {
  DevModel dm;
  std::vector<void*> allocs;
  OrbDev* orb_dev = nullptr;
  double* nat_dev = nullptr;
  int device = 0;
  int smem_optin = 0;
  int num_sms = 0;
  // acceptance feedback for the kernel selection: the kernels add their accepted / attempted step
  // totals to device counters; an async copy after each launch brings them to page-locked host
  // memory and lmc_run reads whatever has arrived (no synchronisation)
  unsigned long long* stats_host = nullptr;
  unsigned long long* stats_dev = nullptr;
  unsigned long long snap[2] = {0, 0};
  double acc_rate = -1.0;   // acceptance ratio of the most recent completed launches, < 0: unknown
};

template <typename T>
static int upload(LmcModel* mdl, const T* host, size_t count, const T** out) {
  void* p = nullptr;
  const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
  if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
  mdl->allocs.push_back(p);
  if (count) {
    e = cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(std::string("cudaMemcpy: ") + cudaGetErrorString(e));
  }
  *out = reinterpret_cast<const T*>(p);
  return 0;
}

extern "C" int lmc_version(void) { return LMC_ABI_VERSION; }
extern "C" const char* lmc_last_error(void) { return g_err.c_str(); }
extern "C" int lmc_row_stride(int n) { return (n + 16) & ~15; }   // always at least one zero pad byte behind the row
extern "C" int64_t lmc_launch_count(void) { return g_launches.load(); }
extern "C" int64_t lmc_env_launch_count(void) { return g_env_launches.load(); }
extern "C" int64_t lmc_c64_launch_count(void) { return g_c64_launches.load(); }
extern "C" int lmc_model_num_features(const LmcModel* m) { return m ? m->dm.F : -1; }

extern "C" int lmc_model_destroy(LmcModel* m) {
  if (!m) return 0;
  for (void* p : m->allocs) cudaFree(p);
  if (m->stats_host) cudaFreeHost(m->stats_host);
  delete m;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// host-only table construction (no CUDA calls: also reachable through lmc_spec_tables_host for tests)
// ------------------------------------------------------------------------------------------------
static int host_orbits(const LmcModelDesc* d, std::vector<OrbDev>& orbs, std::vector<double>& tabA, bool& kone) {
  const int nOrb = d->num_orbits;
  orbs.assign(nOrb, OrbDev());
  kone = true;
  int tabA_len = 0;
  for (int n = 0; n < nOrb; ++n) {
    OrbDev& o = orbs[n];
    o.ftab_off = d->orb_tab_off[n];
    o.T = d->orb_tab_len[n];
    o.K = d->orb_nfunc[n];
    o.fidx = d->orb_fidx[n];
    o.csize = d->orb_csize[n];
    if (o.csize > LMC_MAX_CLUSTER_SITES) return fail("clusters with more than 4 sites are not supported");
    if (o.T > 65535) return fail("flattened tensor longer than 65535");
    o.row_off = (int)d->orb_row_off[n];
    o.row_cnt = (int)(d->orb_row_off[n + 1] - d->orb_row_off[n]);
    for (int i = 0; i < 4; ++i) o.stride[i] = d->orb_stride[n * LMC_MAX_CLUSTER_SITES + i];
    o.w = d->orb_weight[n];
    o.atab_off = tabA_len;
    tabA_len += o.T;
    if (o.K != 1) kone = false;
  }
  // phase-A table: raw tensors (K == 1 everywhere) or tensors contracted with nat * w
  tabA.assign(tabA_len, 0.0);
  for (int n = 0; n < nOrb; ++n) {
    const OrbDev& o = orbs[n];
    for (int t = 0; t < o.T; ++t) {
      if (kone) {
        tabA[o.atab_off + t] = d->ftab[o.ftab_off + t];
      } else {
        double s = 0.0;
        for (int k = 0; k < o.K; ++k) s += d->natural_parameters[o.fidx + k] * o.w * d->ftab[o.ftab_off + k * o.T + t];
        tabA[o.atab_off + t] = s;
      }
    }
  }
  return 0;
}

// Tables of the speculative-batch kernel (lmc_spec.cuh).  The local cluster records of a site are
// MERGED into records with three gathered sites and ONE lookup in a pre-differenced table
//   D[new][tbase + old + NC (c0 + NC (c1 + NC c2))] = sum over the merged terms of
//                                                     coef * (T[idx after] - T[idx before])
// (a 4-site cluster fills a record; a 3-site cluster takes a pair term along; pair / point terms go
// three to a record), so a flip costs about half the table lookups of one lookup per cluster.
// Unused gather slots point at the zero pad byte behind the occupancy row (site index N), so a
// block with k used slots needs NC^(k+1) entries only.
struct SpecTables {
  int ok = 0, NC = 0, L = 0, NQ = 0, nblocks = 0, merged = 0;
  std::vector<double> dtab;          // [NC][L]
  std::vector<unsigned char> rec;    // [N][NQ] x uint2 (s0 | s1<<16, s2 | tbase<<16)
  // per-feature form of the same table for cluster-decomposition models (one feature per orbit):
  // ftab[(new * L + entry) * F + f] = sum over the record's clusters of orbit f of w_f (T[after] - T[before]);
  // an accepted flip's feature change is the sum of its records' rows (lmc_wl.cuh: bookkeeping warp)
  std::vector<double> ftab;          // [NC][L][F], empty when not built
  int F = 0;
};

static void build_spec_tables(const LmcModelDesc* d, const std::vector<OrbDev>& orbs, const std::vector<double>& tabA,
                              bool kone, SpecTables& sp) {
  const int N = d->num_sites, nCls = d->num_classes, nOrb = d->num_orbits;
  sp = SpecTables();
  int NC = 2;
  for (int n = 0; n < nOrb; ++n)
    for (int i = 0; i < orbs[n].csize; ++i) {
      const int hi = i == 0 ? orbs[n].T : orbs[n].stride[i - 1];
      if (orbs[n].stride[i] > 0) NC = std::max(NC, hi / orbs[n].stride[i]);
    }
  if (NC > LMC_MAX_CODES || nCls <= 0 || N >= 65535) return;
  if (const char* e = getenv("LMC_SPEC_TABLES")) { if (atoi(e) == 0) return; }
  std::vector<int> noth(nCls, 0);
  for (int c = 0; c < nCls; ++c) {
    const int* st = d->cls_stride + c * 4;
    int n = 0;
    while (n < 3 && st[n] > 0) ++n;
    for (int i = n; i < 3; ++i) if (st[i] != 0) return;   // other sites must be compacted to the first slots
    noth[c] = std::max(n, 1);                              // point terms take a (dummy) slot
  }
  // contribution of a class-c record for other-site codes oc[0..], old -> new
  auto term = [&](int c, const int* oc, int oldc, int newc) -> double {
    const OrbDev& o = orbs[d->cls_orbit[c]];
    const int* st = d->cls_stride + c * 4;
    const double coef = kone ? d->natural_parameters[o.fidx] * o.w : 1.0;
    long ii = (long)st[3] * oldc, ff = (long)st[3] * newc;
    for (int i = 0; i < 3 && st[i] > 0; ++i) { ii += (long)st[i] * oc[i]; ff += (long)st[i] * oc[i]; }
    if (ii >= o.T || ff >= o.T || newc == oldc) return 0.0;
    return coef * (tabA[o.atab_off + ff] - tabA[o.atab_off + ii]);
  };
  // the same contribution as a feature change: (feature index, w_orbit (T[after] - T[before])); cluster-decomposition
  // models only (kone: the phase-A table IS the orbit's interaction tensor)
  const int F = d->num_features;
  const bool want_f = kone && F <= 32 && !getenv("LMC_SPEC_FTAB_OFF");
  auto term_f = [&](int c, const int* oc, int oldc, int newc, int& fidx) -> double {
    const OrbDev& o = orbs[d->cls_orbit[c]];
    const int* st = d->cls_stride + c * 4;
    fidx = o.fidx;
    long ii = (long)st[3] * oldc, ff = (long)st[3] * newc;
    for (int i = 0; i < 3 && st[i] > 0; ++i) { ii += (long)st[i] * oc[i]; ff += (long)st[i] * oc[i]; }
    if (ii >= o.T || ff >= o.T || newc == oldc) return 0.0;
    return o.w * (tabA[o.atab_off + ff] - tabA[o.atab_off + ii]);
  };
  // a merged record: up to three gathered sites and the cluster terms folded into its table block;
  // slot[i] = which gathered site feeds the i-th other site of the term's cluster
  struct Term { int cls; int slot[3]; };
  struct MRec { uint16_t site[3]; int ns; std::vector<Term> terms; };
  const uint16_t DUMMY = (uint16_t)N;   // zero pad byte of the occupancy row
  // merge level 3: site-set cover (a record takes every cluster whose other sites are among its three
  // gathered sites: an FCC tetrahedron record also carries its three triangles and nearest-neighbour
  // pairs), 2: one cluster + pair terms per record, 1: one cluster per record
  int max_level = 3;
  if (const char* e = getenv("LMC_SPEC_MERGE")) max_level = std::min(3, std::max(1, atoi(e) + 1));
  // Bytes of shared memory the difference table may take.  Two passes: first the highest level whose table leaves room
  // for 28 walkers per SM next to seven or two copies of it (flip / swap kernels, blob <= 40 KB); failing that the
  // highest level that fits ONE copy per SM beside fourteen walkers (the table-flip kernel of lmc_spec_tf.cuh: a
  // five-species rocksalt model takes 47 KB at level 1 with 152 records per site, 112 KB at level 3 with 64)
  size_t budgets[2] = {39 * 1024, 114 * 1024};
  if (const char* e = getenv("LMC_SPEC_TABLE_KB")) budgets[0] = budgets[1] = (size_t)atoi(e) * 1024;
  for (int pass = 0; pass < 2 && !sp.ok; ++pass)
  for (int level = max_level; level >= 1 && !sp.ok; --level) {
    const size_t budget = budgets[pass];
    // deduplicated [new][entry] blocks; block 0 = zeros (padding records)
    std::vector<std::vector<double>> store;
    std::vector<std::vector<double>> store_f;   // feature rows of the blocks ([new][entry][F]), want_f only
    std::vector<long> store_base;
    std::map<std::vector<int>, int> keys;   // (ns, cls, slots ...) -> store index
    long L = (long)NC * NC * NC * NC;
    store.emplace_back((size_t)L * NC, 0.0);
    store_f.emplace_back(want_f ? (size_t)L * NC * F : 0, 0.0);
    store_base.push_back(0);
    auto block_of = [&](MRec& r) -> int {
      std::sort(r.terms.begin(), r.terms.end(), [](const Term& a, const Term& b) {
        if (a.cls != b.cls) return a.cls < b.cls;
        for (int i = 0; i < 3; ++i) if (a.slot[i] != b.slot[i]) return a.slot[i] < b.slot[i];
        return false;
      });
      std::vector<int> key{r.ns};
      for (const Term& t : r.terms) { key.push_back(t.cls); key.push_back(t.slot[0]); key.push_back(t.slot[1]); key.push_back(t.slot[2]); }
      auto it = keys.find(key);
      if (it != keys.end()) return it->second;
      long combos = 1;
      for (int i = 0; i < r.ns; ++i) combos *= NC;
      const long len = combos * NC;
      std::vector<double> blk((size_t)len * NC, 0.0);
      std::vector<double> blkf(want_f ? (size_t)len * NC * F : 0, 0.0);
      for (long q = 0; q < combos; ++q) {
        const int code[3] = {(int)(q % NC), (int)((q / NC) % NC), (int)(q / ((long)NC * NC))};
        for (int oldc = 0; oldc < NC; ++oldc)
          for (int newc = 0; newc < NC; ++newc) {
            double v = 0.0;
            for (const Term& t : r.terms) {
              const int oc[3] = {code[t.slot[0]], code[t.slot[1]], code[t.slot[2]]};
              v += term(t.cls, oc, oldc, newc);
              if (want_f) {
                int fi = 0;
                const double vf = term_f(t.cls, oc, oldc, newc, fi);
                blkf[((size_t)newc * len + oldc + NC * q) * F + fi] += vf;
              }
            }
            blk[(size_t)newc * len + oldc + NC * q] = v;
          }
      }
      int idx = -1;
      for (size_t b = 1; b < store.size() && idx < 0; ++b)
        if (store[b].size() == blk.size() && memcmp(store[b].data(), blk.data(), blk.size() * 8) == 0 &&
            (!want_f || memcmp(store_f[b].data(), blkf.data(), blkf.size() * 8) == 0)) idx = (int)b;
      if (idx < 0) {
        store.push_back(std::move(blk));
        store_f.push_back(std::move(blkf));
        store_base.push_back(L);
        L += len;
        idx = (int)store.size() - 1;
      }
      keys.emplace(std::move(key), idx);
      return idx;
    };
    std::vector<std::vector<MRec>> merged(N);
    size_t nq = 0;
    auto nreal = [&](int c) { return d->cls_stride[c * 4] > 0 ? noth[c] : 0; };   // point terms gather nothing
    // record with consecutive slots for the given clusters (levels 1 and 2)
    auto packed_rec = [&](std::initializer_list<const uint16_t*> rcs) {
      MRec r{{DUMMY, DUMMY, DUMMY}, 0, {}};
      for (const uint16_t* rc : rcs) {
        const int c = rc[3];
        Term t{c, {0, 0, 0}};
        for (int i = 0; i < nreal(c); ++i) { t.slot[i] = r.ns; r.site[r.ns++] = rc[i]; }
        if (nreal(c) == 0 && r.ns < 3) ++r.ns;   // a point term keeps its (dummy) slot
        r.terms.push_back(t);
      }
      return r;
    };
    for (int i = 0; i < N; ++i) {
      std::vector<MRec>& out = merged[i];
      if (level == 3) {
        std::vector<const uint16_t*> recs;
        for (int64_t r = d->site_rec_off[i]; r < d->site_rec_off[i + 1]; ++r) recs.push_back(d->site_rec + r * 4);
        std::stable_sort(recs.begin(), recs.end(), [&](const uint16_t* a, const uint16_t* b) { return nreal(a[3]) > nreal(b[3]); });
        for (const uint16_t* rc : recs) {
          const int c = rc[3], n = nreal(c);
          int best = -1, best_score = -1;
          for (size_t k = 0; k < out.size(); ++k) {
            MRec& r = out[k];
            int missing = 0;
            for (int a = 0; a < n; ++a) {
              bool have = false;
              for (int b = 0; b < r.ns; ++b) have = have || r.site[b] == rc[a];
              for (int b = 0; b < a; ++b) have = have || rc[b] == rc[a];   // repeated site of an aliased cluster
              if (!have) ++missing;
            }
            if (r.ns + missing > 3) continue;
            // prefer records that already hold the sites, then the fullest one
            const int score = (n - missing) * 16 + (missing == 0 ? 8 : 0) + r.ns;
            if (score > best_score) { best_score = score; best = (int)k; }
          }
          if (best < 0) { out.push_back(MRec{{DUMMY, DUMMY, DUMMY}, 0, {}}); best = (int)out.size() - 1; }
          MRec& r = out[best];
          Term t{c, {0, 0, 0}};
          for (int a = 0; a < n; ++a) {
            int at = -1;
            for (int b = 0; b < r.ns; ++b) if (r.site[b] == rc[a]) at = b;
            if (at < 0) { at = r.ns; r.site[r.ns++] = rc[a]; }
            t.slot[a] = at;
          }
          r.terms.push_back(t);
        }
      } else {
        std::vector<std::vector<const uint16_t*>> P(nCls);   // pair / point records by class
        std::vector<const uint16_t*> T;
        for (int64_t r = d->site_rec_off[i]; r < d->site_rec_off[i + 1]; ++r) {
          const uint16_t* rc = d->site_rec + r * 4;
          const int c = rc[3], n = noth[c];
          if (n == 3 || level == 1) out.push_back(packed_rec({rc}));
          else if (n == 2) T.push_back(rc);
          else P[c].push_back(rc);
        }
        std::vector<size_t> head(nCls, 0);
        auto remaining = [&](int c) { return P[c].size() - head[c]; };
        for (const uint16_t* t : T) {   // a 3-site cluster takes a pair term of the fullest class along
          int best = -1;
          for (int c = 0; c < nCls; ++c)
            if (remaining(c) > 0 && (best < 0 || remaining(c) > remaining(best))) best = c;
          if (best >= 0) out.push_back(packed_rec({t, P[best][head[best]++]}));
          else out.push_back(packed_rec({t}));
        }
        for (int c = 0; c < nCls; ++c)
          while (remaining(c) >= 3) {
            out.push_back(packed_rec({P[c][head[c]], P[c][head[c] + 1], P[c][head[c] + 2]}));
            head[c] += 3;
          }
        std::vector<const uint16_t*> rest;
        for (int c = 0; c < nCls; ++c)
          while (remaining(c) > 0) rest.push_back(P[c][head[c]++]);
        for (size_t k = 0; k < rest.size(); k += 3) {
          if (k + 2 < rest.size()) out.push_back(packed_rec({rest[k], rest[k + 1], rest[k + 2]}));
          else if (k + 1 < rest.size()) out.push_back(packed_rec({rest[k], rest[k + 1]}));
          else out.push_back(packed_rec({rest[k]}));
        }
      }
      nq = std::max(nq, out.size());
    }
    const int NQ = (int)((nq + 7) & ~size_t(7));
    std::vector<unsigned char> rec((size_t)std::max(N * NQ * 8, 16), 0);
    bool ok = true;
    for (int i = 0; i < N && ok; ++i) {
      uint32_t* p = reinterpret_cast<uint32_t*>(rec.data() + (size_t)i * NQ * 8);
      for (int k = 0; k < NQ; ++k) {
        if (k < (int)merged[i].size()) {
          MRec& mr = merged[i][k];
          const long tb = store_base[block_of(mr)];
          if (L > 65535 || (size_t)L * NC * 8 > budget) { ok = false; break; }
          p[2 * k] = (uint32_t)mr.site[0] | ((uint32_t)mr.site[1] << 16);
          p[2 * k + 1] = (uint32_t)mr.site[2] | ((uint32_t)tb << 16);
        } else {   // padding: three zero bytes, zero block
          p[2 * k] = (uint32_t)DUMMY | ((uint32_t)DUMMY << 16);
          p[2 * k + 1] = (uint32_t)DUMMY;
        }
      }
    }
    if (!ok) continue;
    sp.dtab.assign((size_t)L * NC, 0.0);
    for (size_t b = 0; b < store.size(); ++b) {
      const size_t len = store[b].size() / NC;
      for (int newc = 0; newc < NC; ++newc)
        memcpy(&sp.dtab[(size_t)newc * L + store_base[b]], &store[b][(size_t)newc * len], len * 8);
    }
    if (want_f && (size_t)L * NC * F * 8 <= (size_t(64) << 20)) {
      sp.ftab.assign((size_t)L * NC * F, 0.0);
      sp.F = F;
      for (size_t b = 0; b < store.size(); ++b) {
        const size_t len = store[b].size() / NC;
        for (int newc = 0; newc < NC; ++newc)
          memcpy(&sp.ftab[((size_t)newc * L + store_base[b]) * F], &store_f[b][(size_t)newc * len * F], len * F * 8);
      }
    }
    sp.rec = std::move(rec);
    sp.ok = 1; sp.NC = NC; sp.L = (int)L; sp.NQ = NQ; sp.nblocks = (int)store.size(); sp.merged = level - 1;
  }
}

// Environment words of the speculative kernel (lmc_spec.cuh, ENV variants).  Every ACTIVE site keeps the species codes
// its merged records gather as one packed word per lane of the 4-lane step group: record i of lane l (records
// 2 (l + 4 q) + e, i = 2 q + e, the order spec_rec2 walks them) owns the bit field [i * 3b, (i + 1) * 3b) of the lane's
// chunk, slot j of the record the b bits at j * b inside it (b = bits per species code).  A rejected step then costs one
// chunk load per flip instead of three byte gathers per record; an ACCEPTED flip of site s xors (old ^ new) into the
// slots of every site that gathers s (reverse map).  The second flip of a swap sees the first one applied through a
// per-pair slot mask (adjacency ordinal + mask table), the PATCH of the gather variant.
struct EnvTables {
  int ok = 0, b = 0, nrl = 0, nrlp = 0, wide = 0, NA = 0, RV = 0, pair_ok = 0;
  std::vector<uint16_t> tb;              // [N][4][nrlp] table base of the lane's i-th record
  // the same lists deduplicated (translation-equivalent sites share theirs): cls[site] -> row of tbc; staged to shared
  // memory with the table blob while there are at most 256 classes / 8 KB of lists (ncls == 0: not built)
  int ncls = 0;
  std::vector<uint8_t> cls;              // [N]
  std::vector<uint16_t> tbc;             // [ncls][4][nrlp]
  std::vector<uint32_t> rev;             // [NA][RV] active index of the gathering site | bit position << 16; ~0 = none
  // [NA][NA][4] x (u32 | u64 when wide): lowest slot bits, per lane chunk, where the COLUMN site sits among the codes the
  // ROW site gathers (all zero for pairs that do not see each other); swaps only, built while it stays below 64 MB
  std::vector<unsigned char> pair;
};

static void build_env_tables(const LmcModelDesc* d, const SpecTables& sp, EnvTables& ev) {
  ev = EnvTables();
  if (!sp.ok || getenv("LMC_SPEC_ENV_OFF")) return;
  const int N = d->num_sites, NQ = sp.NQ, NC = sp.NC;
  if (NC > 4) return;                                   // kernels are instantiated for 1 and 2 bits per code
  const int b = NC <= 2 ? 1 : 2, fb = 3 * b;
  const int nrl = NQ / 4;                               // NQ is a multiple of 8
  if (nrl * fb > 64) return;
  const int wide = nrl * fb > 32 ? 1 : 0;
  const int lane_bits = wide ? 64 : 32;
  const int NA = d->sl_site_off[d->num_sublattices];
  if (NA <= 0 || NA > 65535) return;
  std::vector<int> aidx(N, -1);
  for (int a = 0; a < NA; ++a) {
    const int s = d->sl_sites[a];
    if (s < 0 || s >= N || aidx[s] >= 0) return;        // overlapping sublattices: not served
    aidx[s] = a;
  }
  const int nrlp = (nrl + 7) & ~7;
  const size_t pair_el = wide ? 8 : 4;
  const bool want_pair = (size_t)NA * NA * 4 * pair_el <= (size_t(64) << 20);
  ev.tb.assign((size_t)N * 4 * nrlp, 0);
  if (want_pair) ev.pair.assign((size_t)NA * NA * 4 * pair_el, 0);
  std::vector<std::vector<uint32_t>> rev(NA);
  const uint32_t* rec = reinterpret_cast<const uint32_t*>(sp.rec.data());
  for (int k = 0; k < N; ++k)
    for (int l = 0; l < 4; ++l)
      for (int i = 0; i < nrl; ++i) {
        const int r = 2 * (l + 4 * (i >> 1)) + (i & 1);
        const uint32_t x = rec[((size_t)k * NQ + r) * 2], y = rec[((size_t)k * NQ + r) * 2 + 1];
        ev.tb[((size_t)k * 4 + l) * nrlp + i] = (uint16_t)(y >> 16);
        if (aidx[k] < 0) continue;
        const int site[3] = {(int)(x & 0xffffu), (int)(x >> 16), (int)(y & 0xffffu)};
        for (int j = 0; j < 3; ++j) {
          const int s = site[j];
          if (s >= N || aidx[s] < 0) continue;          // pad slot or a site that never changes
          const int bit = i * fb + j * b;
          rev[aidx[s]].push_back((uint32_t)aidx[k] | ((uint32_t)(l * lane_bits + bit) << 16));
          if (want_pair) {
            unsigned char* p = ev.pair.data() + (((size_t)aidx[k] * NA + aidx[s]) * 4 + l) * pair_el;
            if (wide) { unsigned long long v; memcpy(&v, p, 8); v |= 1ull << bit; memcpy(p, &v, 8); }
            else { uint32_t v; memcpy(&v, p, 4); v |= 1u << bit; memcpy(p, &v, 4); }
          }
        }
      }
  size_t rv = 1;
  for (int a = 0; a < NA; ++a) rv = std::max(rv, rev[a].size());
  ev.RV = (int)((rv + 31) & ~size_t(31));
  ev.rev.assign((size_t)NA * ev.RV, 0xffffffffu);
  for (int a = 0; a < NA; ++a) std::copy(rev[a].begin(), rev[a].end(), ev.rev.begin() + (size_t)a * ev.RV);
  {
    std::map<std::vector<uint16_t>, int> seen;
    std::vector<uint8_t> cls(N, 0);
    std::vector<uint16_t> tbc;
    bool fits = true;
    for (int k = 0; k < N && fits; ++k) {
      std::vector<uint16_t> row(ev.tb.begin() + (size_t)k * 4 * nrlp, ev.tb.begin() + (size_t)(k + 1) * 4 * nrlp);
      auto it = seen.find(row);
      if (it == seen.end()) {
        if (seen.size() >= 256 || (seen.size() + 1) * row.size() * 2 > 8 * 1024) { fits = false; break; }
        it = seen.emplace(row, (int)seen.size()).first;
        tbc.insert(tbc.end(), row.begin(), row.end());
      }
      cls[k] = (uint8_t)it->second;
    }
    if (fits) { ev.ncls = (int)seen.size(); ev.cls = std::move(cls); ev.tbc = std::move(tbc); }
  }
  ev.ok = 1; ev.b = b; ev.nrl = nrl; ev.nrlp = nrlp; ev.wide = wide; ev.NA = NA; ev.pair_ok = want_pair ? 1 : 0;
}

// Compact environment words of lmc_spec_c64.cuh: ONE 64-bit word per active site and walker, in shared memory.  Every
// merged record owns a field of three codes (3 b bits) inside the word of its site; a record whose first gathered site is
// the last one of another record starts on that record's last slot (chains of overlapping fields), which is what brings
// the 22 records of the FCC cluster set from 66 to 64 bits.  Per record the kernel needs the table base and the shift of
// the field: one u32 list per lane of the four-lane step group, deduplicated over sites (translation-equivalent sites
// share theirs) and staged with the table blob.  An accepted flip of site s xors (old ^ new) into the bits that hold s in
// the words of the sites that gather it (reverse map); the second flip of a swap sees the first through the pair mask.
struct C64Tables {
  int ok = 0, b = 0, nrl = 0, nrlp = 0, NA = 0, RV = 0, ncls = 0, bits = 0;
  std::vector<uint32_t> desc;              // [ncls][4][nrlp] table base | shift << 16 of the lane's i-th record
  std::vector<uint8_t> cls;                // [N] class of a site
  std::vector<uint32_t> rev;               // [NA][RV] active index of the gathering site | bit << 16; ~0 = none
  std::vector<unsigned long long> pair;    // [NA][NA] bits of the ROW site's word that hold the COLUMN site (lowest bit of each slot)
};

static void build_c64_tables(const LmcModelDesc* d, const SpecTables& sp, C64Tables& ct) {
  ct = C64Tables();
  if (!sp.ok || getenv("LMC_SPEC_C64_OFF")) return;
  const int N = d->num_sites, NQ = sp.NQ, NC = sp.NC;
  if (NC > 4) return;
  const int b = NC <= 2 ? 1 : 2, fb = 3 * b;
  const int nrl = NQ / 4, nrlp = (nrl + 3) & ~3;
  const int NA = d->sl_site_off[d->num_sublattices];
  if (NA <= 0 || NA > 65535 || (size_t)NA * NA * 8 > (size_t(64) << 20)) return;
  std::vector<int> aidx(N, -1);
  for (int a = 0; a < NA; ++a) {
    const int s = d->sl_sites[a];
    if (s < 0 || s >= N || aidx[s] >= 0) return;
    aidx[s] = a;
  }
  const uint32_t* rec = reinterpret_cast<const uint32_t*>(sp.rec.data());
  std::vector<std::vector<uint32_t>> rev(NA);
  ct.pair.assign((size_t)NA * NA, 0ull);
  ct.cls.assign(N, 0);
  std::map<std::vector<uint32_t>, int> seen;
  int maxbits = 0;
  std::vector<int> S0(NQ), S1(NQ), S2(NQ), TB(NQ), succ(NQ), pred(NQ), pos(NQ);
  for (int k = 0; k < N; ++k) {
    for (int r = 0; r < NQ; ++r) {
      const uint32_t x = rec[((size_t)k * NQ + r) * 2], y = rec[((size_t)k * NQ + r) * 2 + 1];
      S0[r] = (int)(x & 0xffffu); S1[r] = (int)(x >> 16); S2[r] = (int)(y & 0xffffu); TB[r] = (int)(y >> 16);
      succ[r] = pred[r] = -1; pos[r] = 0;
    }
    auto real = [&](int r) { return TB[r] != 0 || S0[r] < N || S1[r] < N || S2[r] < N; };
    // chains: B starts on A's last slot when that is the site B gathers first
    for (int a = 0; a < NQ; ++a) {
      if (!real(a) || S2[a] >= N) continue;
      for (int c = 0; c < NQ && succ[a] < 0; ++c) {
        if (c == a || !real(c) || pred[c] >= 0 || S0[c] != S2[a]) continue;
        bool cycle = false;
        for (int w = c; w >= 0 && !cycle; w = succ[w]) cycle = w == a;
        if (cycle) continue;
        succ[a] = c; pred[c] = a;
      }
    }
    int at = 0;
    for (int h = 0; h < NQ; ++h) {
      if (!real(h) || pred[h] >= 0) continue;
      int w = h;
      pos[w] = at;
      while (succ[w] >= 0) { pos[succ[w]] = pos[w] + 2 * b; w = succ[w]; }
      at = pos[w] + fb;
    }
    maxbits = std::max(maxbits, at);
    if (at > 64) return;
    std::vector<uint32_t> row((size_t)4 * nrlp, 0u);
    for (int l = 0; l < 4; ++l)
      for (int i = 0; i < nrl; ++i) {
        const int r = 2 * (l + 4 * (i >> 1)) + (i & 1);
        row[(size_t)l * nrlp + i] = (uint32_t)TB[r] | ((uint32_t)pos[r] << 16);
      }
    auto it = seen.find(row);
    if (it == seen.end()) {
      if (seen.size() >= 256 || (seen.size() + 1) * row.size() * 4 > 8 * 1024) return;
      it = seen.emplace(row, (int)seen.size()).first;
      ct.desc.insert(ct.desc.end(), row.begin(), row.end());
    }
    ct.cls[k] = (uint8_t)it->second;
    if (aidx[k] < 0) continue;
    for (int r = 0; r < NQ; ++r) {
      if (!real(r)) continue;
      const int site[3] = {S0[r], S1[r], S2[r]};
      for (int j = 0; j < 3; ++j) {
        const int s = site[j];
        if (s >= N || aidx[s] < 0) continue;
        const int bit = pos[r] + j * b;
        unsigned long long& pm = ct.pair[(size_t)aidx[k] * NA + aidx[s]];
        if (pm & (1ull << bit)) continue;           // the shared slot of two chained records: one entry
        pm |= 1ull << bit;
        rev[aidx[s]].push_back((uint32_t)aidx[k] | ((uint32_t)bit << 16));
      }
    }
  }
  size_t rv = 1;
  for (int a = 0; a < NA; ++a) rv = std::max(rv, rev[a].size());
  ct.RV = (int)((rv + 31) & ~size_t(31));
  ct.rev.assign((size_t)NA * ct.RV, 0xffffffffu);
  for (int a = 0; a < NA; ++a) std::copy(rev[a].begin(), rev[a].end(), ct.rev.begin() + (size_t)a * ct.RV);
  ct.ok = 1; ct.b = b; ct.nrl = nrl; ct.nrlp = nrlp; ct.NA = NA; ct.ncls = (int)seen.size(); ct.bits = maxbits;
}

// host-only: build the tables of the speculative kernel for a model description (tests / diagnostics)
// info = {ok, NC, L, NQ, nblocks, merged, table bytes, record bytes}
extern "C" int lmc_spec_tables_host(const LmcModelDesc* d, int32_t* info, double* dtab_out, int64_t dtab_cap,
                                    uint8_t* rec_out, int64_t rec_cap) {
  if (!d || !info) return fail("null argument");
  std::vector<OrbDev> orbs;
  std::vector<double> tabA;
  bool kone = true;
  if (host_orbits(d, orbs, tabA, kone)) return -1;
  SpecTables sp;
  build_spec_tables(d, orbs, tabA, kone, sp);
  info[0] = sp.ok; info[1] = sp.NC; info[2] = sp.L; info[3] = sp.NQ; info[4] = sp.nblocks; info[5] = sp.merged;
  info[6] = (int32_t)(sp.dtab.size() * 8); info[7] = (int32_t)sp.rec.size();
  if (dtab_out && (int64_t)sp.dtab.size() <= dtab_cap) memcpy(dtab_out, sp.dtab.data(), sp.dtab.size() * 8);
  if (rec_out && (int64_t)sp.rec.size() <= rec_cap) memcpy(rec_out, sp.rec.data(), sp.rec.size());
  return 0;
}

// host-only: compact environment-word tables of a model description (tests / diagnostics)
// info = {ok, bits per code, records per lane, padded records per lane, active sites, reverse entries per site, classes, bits used}
extern "C" int lmc_spec_c64_host(const LmcModelDesc* d, int32_t* info, uint32_t* desc_out, int64_t desc_cap, uint8_t* cls_out,
                                 int64_t cls_cap, uint32_t* rev_out, int64_t rev_cap, uint64_t* pair_out, int64_t pair_cap) {
  if (!d || !info) return fail("null argument");
  std::vector<OrbDev> orbs;
  std::vector<double> tabA;
  bool kone = true;
  if (host_orbits(d, orbs, tabA, kone)) return -1;
  SpecTables sp;
  build_spec_tables(d, orbs, tabA, kone, sp);
  C64Tables ct;
  build_c64_tables(d, sp, ct);
  info[0] = ct.ok; info[1] = ct.b; info[2] = ct.nrl; info[3] = ct.nrlp; info[4] = ct.NA; info[5] = ct.RV; info[6] = ct.ncls;
  info[7] = ct.bits;
  if (desc_out && (int64_t)ct.desc.size() <= desc_cap) memcpy(desc_out, ct.desc.data(), ct.desc.size() * 4);
  if (cls_out && (int64_t)ct.cls.size() <= cls_cap) memcpy(cls_out, ct.cls.data(), ct.cls.size());
  if (rev_out && (int64_t)ct.rev.size() <= rev_cap) memcpy(rev_out, ct.rev.data(), ct.rev.size() * 4);
  if (pair_out && (int64_t)ct.pair.size() <= pair_cap) memcpy(pair_out, ct.pair.data(), ct.pair.size() * 8);
  return 0;
}

// host-only: environment-word tables of a model description (tests / diagnostics)
// info = {ok, bits per code, records per lane, padded records per lane, wide, active sites, reverse entries per site, pair table built}
extern "C" int lmc_spec_env_host(const LmcModelDesc* d, int32_t* info, uint16_t* tb_out, int64_t tb_cap, uint32_t* rev_out,
                                 int64_t rev_cap, uint8_t* pair_out, int64_t pair_cap) {
  if (!d || !info) return fail("null argument");
  std::vector<OrbDev> orbs;
  std::vector<double> tabA;
  bool kone = true;
  if (host_orbits(d, orbs, tabA, kone)) return -1;
  SpecTables sp;
  build_spec_tables(d, orbs, tabA, kone, sp);
  EnvTables ev;
  build_env_tables(d, sp, ev);
  C64Tables c64;
  build_c64_tables(d, sp, c64);
  info[0] = ev.ok; info[1] = ev.b; info[2] = ev.nrl; info[3] = ev.nrlp; info[4] = ev.wide; info[5] = ev.NA; info[6] = ev.RV;
  info[7] = ev.pair_ok;
  if (tb_out && (int64_t)ev.tb.size() <= tb_cap) memcpy(tb_out, ev.tb.data(), ev.tb.size() * 2);
  if (rev_out && (int64_t)ev.rev.size() <= rev_cap) memcpy(rev_out, ev.rev.data(), ev.rev.size() * 4);
  if (pair_out && (int64_t)ev.pair.size() <= pair_cap) memcpy(pair_out, ev.pair.data(), ev.pair.size());
  return 0;
}

extern "C" int lmc_model_create(const LmcModelDesc* d, LmcModel** out) {
  if (!d || !out) return fail("null argument");
  if (d->abi_version != LMC_ABI_VERSION) return fail("ABI version mismatch");
  if (d->num_sites <= 0 || d->num_sites > 65535) return fail("num_sites must be in 1..65535 (u16 site indices)");
  if (d->num_classes > 65535) return fail("too many record classes");
  if (d->num_sublattices < 1 || d->num_sublattices > LMC_MAX_SUBLATTICES) return fail("1..8 active sublattices");
  if (d->tf_num_dims > LMC_MAX_DIMS || d->tf_num_flips > LMC_MAX_TABLE_FLIPS) return fail("flip table too large");
  LmcModel* mdl = new LmcModel();
  DevModel& m = mdl->dm;
  memset(&m, 0, sizeof(m));
  cudaGetDevice(&mdl->device);
  cudaDeviceGetAttribute(&mdl->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, mdl->device);
  cudaDeviceGetAttribute(&mdl->num_sms, cudaDevAttrMultiProcessorCount, mdl->device);
  if (cudaHostAlloc((void**)&mdl->stats_host, 2 * sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess &&
      cudaMalloc((void**)&mdl->stats_dev, 2 * sizeof(unsigned long long)) == cudaSuccess) {
    mdl->stats_host[0] = mdl->stats_host[1] = 0;
    cudaMemset(mdl->stats_dev, 0, 2 * sizeof(unsigned long long));
    mdl->allocs.push_back(mdl->stats_dev);
  } else {
    if (mdl->stats_host) cudaFreeHost(mdl->stats_host);
    mdl->stats_host = nullptr;
    mdl->stats_dev = nullptr;
    (void)cudaGetLastError();
  }
  m.N = d->num_sites;
  m.Npad = lmc_row_stride(d->num_sites);
  m.F = d->num_features;
  m.Fce = d->num_ce_features;
  m.nOrb = d->num_orbits;
  m.nCls = d->num_classes;
  m.nSl = d->num_sublattices;
  m.size = d->supercell_size;
  m.feature0 = d->feature0;
  int rc = 0;
#define UP(T, src, cnt, dst)                                         \
  if ((rc = upload<T>(mdl, (const T*)(src), (size_t)(cnt), (const T**)&(dst)))) { \
    lmc_model_destroy(mdl);                                          \
    return rc;                                                       \
  }
  // orbit descriptors + phase-A table
  std::vector<OrbDev> orbs;
  std::vector<double> tabA;
  bool kone = true;
  if (host_orbits(d, orbs, tabA, kone)) { lmc_model_destroy(mdl); return -1; }
  m.kone = kone ? 1 : 0;
  m.Kmax = 1;
  for (const OrbDev& o : orbs) m.Kmax = std::max(m.Kmax, o.K);
  const int tabA_len = (int)tabA.size();
  m.tabA_len = tabA_len;
  std::vector<double> qtab;
  std::vector<double2> qd_host;   // (charge, diagonal) per (site, code) of a factorised Ewald matrix
  // Ewald.  Generic form: keep the TRANSPOSE so that the reference's column gathers become row
  // gathers.  Ewald matrices are charge products times a geometric site kernel,
  // M[i,j] = q_i q_j K[site_i, site_j] (pymatgen EwaldSummation), which turns the two strided E-long
  // row gathers of a flip into ONE contiguous N-long row; detected here, verified entry by entry.
  m.E = d->ewald_size;
  if (m.E >= 65535) { lmc_model_destroy(mdl); return fail("Ewald matrices with 65535 or more rows are not supported (u16 row cache)"); }
  if (m.E > 0) {
    m.ewW = d->ewald_width;
    m.ewF = d->ewald_feature;
    const size_t E = (size_t)m.E, N = (size_t)m.N;
    const double* M = d->ewald_matrix;
    bool fact = true;
    if (const char* e = getenv("LMC_EWALD_FACTORIZE")) fact = atoi(e) != 0;
    std::vector<int> row_site(E, -1), site_ref(N, -1);
    for (size_t k = 0; k < N && fact; ++k)
      for (int c = 0; c < m.ewW; ++c) {
        const int e = d->ewald_inds[k * m.ewW + c];
        if (e < 0) continue;
        if ((size_t)e >= E || row_site[e] >= 0) { fact = false; break; }
        row_site[e] = (int)k;
        if (site_ref[k] < 0) site_ref[k] = e;
      }
    for (size_t e = 0; e < E && fact; ++e) fact = row_site[e] >= 0;
    std::vector<double> q(E, 0.0), dg(E, 0.0), K;
    double mmax = 0.0;
    if (fact) {
      for (size_t i = 0; i < E * E; ++i) mmax = std::max(mmax, fabs(M[i]));
      // relative charge of row i w.r.t. the reference row of its site, from the column where the
      // reference row is largest (on another site)
      for (size_t i = 0; i < E && fact; ++i) {
        const int ref = site_ref[row_site[i]];
        size_t jbest = 0; double best = -1.0;
        for (size_t j = 0; j < E; ++j)
          if (row_site[j] != row_site[i] && fabs(M[(size_t)ref * E + j]) > best) { best = fabs(M[(size_t)ref * E + j]); jbest = j; }
        if (!(best > 1e-12 * mmax)) { fact = false; break; }
        q[i] = M[i * E + jbest] / M[(size_t)ref * E + jbest];
        dg[i] = M[i * E + i];
      }
    }
    if (fact) {
      K.assign(N * N, 0.0);
      for (size_t s = 0; s < N; ++s)
        for (size_t t = 0; t < N; ++t)
          if (s != t && site_ref[s] >= 0 && site_ref[t] >= 0) K[s * N + t] = M[(size_t)site_ref[s] * E + site_ref[t]];
      const double tol = 1e-12 * mmax;
      for (size_t i = 0; i < E && fact; ++i)
        for (size_t j = 0; j < E; ++j) {
          if (row_site[i] == row_site[j]) continue;
          if (fabs(M[i * E + j] - q[i] * q[j] * K[(size_t)row_site[i] * N + row_site[j]]) > tol ||
              fabs(M[i * E + j] - M[j * E + i]) > tol) { fact = false; break; }
        }
    }
    std::vector<uint8_t> qidx(E, 0);
    if (fact) {   // distinct charge values -> byte index (shared-memory table), else stay generic
      for (size_t i = 0; i < E && fact; ++i) {
        size_t f = 0;
        while (f < qtab.size() && qtab[f] != q[i]) ++f;
        if (f == qtab.size()) { if (qtab.size() >= 255) { fact = false; break; } qtab.push_back(q[i]); }
        qidx[i] = (uint8_t)f;
      }
    }
    if (fact) {
      m.ewNQ = (int)qtab.size();
      qtab.push_back(0.0);   // vacancy
      // exactly symmetric site kernel (the check above bounds the asymmetry by 1e-12 max|M|): the potential
      // cache reads rows where the definition has columns
      for (size_t s = 0; s < N; ++s)
        for (size_t t = s + 1; t < N; ++t) K[s * N + t] = K[t * N + s] = 0.5 * (K[s * N + t] + K[t * N + s]);
      std::vector<double2> qd(N * (size_t)m.ewW, make_double2(0.0, 0.0));
      for (size_t k = 0; k < N; ++k)
        for (int c = 0; c < m.ewW; ++c) {
          const int e = d->ewald_inds[k * m.ewW + c];
          if (e >= 0) qd[k * m.ewW + c] = make_double2(q[e], dg[e]);
        }
      UP(double2, qd.data(), qd.size(), m.ewQD);
      qd_host = qd;
      UP(double, K.data(), N * N, m.ewK);
      UP(double, q.data(), E, m.ewQ);
      UP(double, dg.data(), E, m.ewD);
      UP(uint8_t, qidx.data(), E, m.ewQidx);
      m.ewMt = nullptr;
    } else {
      qtab.clear();
      std::vector<double> mt(E * E);
      for (size_t i = 0; i < E; ++i)
        for (size_t j = 0; j < E; ++j) mt[j * E + i] = M[i * E + j];
      UP(double, mt.data(), E * E, m.ewMt);
    }
    UP(int, d->ewald_inds, (size_t)m.N * m.ewW, m.ewInds);
  }
  // speculative-batch kernel tables (merged records + pre-differenced table), see build_spec_tables
  SpecTables sp;
  build_spec_tables(d, orbs, tabA, kone, sp);
  const std::vector<double>& dtab = sp.dtab;
  const std::vector<unsigned char>& sprec = sp.rec;
  m.spOK = sp.ok; m.spNC = sp.NC; m.spL = sp.L; m.spNQ = sp.NQ; m.spSb = sp.NQ * 8;
  EnvTables ev;
  build_env_tables(d, sp, ev);
  C64Tables c64;
  build_c64_tables(d, sp, c64);
  // blob: cls (nCls + 1 entries, the last is the all-zero padding class) | tabA | nat | orb
  {
    const int C = m.nCls + 1;
    size_t off = 0;
    const size_t off_cls = off; off += (size_t)C * 16;
    m.off_tabA = (int)off; off += ((size_t)tabA_len * 8 + 15) & ~size_t(15);
    m.off_nat = (int)off; off += ((size_t)m.F * 8 + 15) & ~size_t(15);
    m.off_orb = (int)off; off += ((size_t)m.nOrb * sizeof(OrbDev) + 15) & ~size_t(15);
    m.off_qtab = (int)off; off += ((size_t)std::max<size_t>(qtab.size(), 2) * 8 + 15) & ~size_t(15);
    m.off_dtab = (int)off; off += (dtab.size() * 8 + 15) & ~size_t(15);
    // environment words: per-class table-base lists and the class of every site (speculative kernels only)
    // per-sublattice chemical-potential / charge / diagonal tables for the speculative kernels (groups of a warp look up
    // DIFFERENT entries at once: shared memory serves that, the constant bank serialises it); filled below
    m.off_ctab = (int)off; off += 3 * LMC_MAX_SUBLATTICES * LMC_MAX_CODES * 8;
    m.blob_bytes = (int)off;      // what every kernel but the environment-word variants stages
    m.off_c64desc = (int)off; off += (c64.desc.size() * 4 + 15) & ~size_t(15);
    m.off_c64cls = (int)off; off += (c64.cls.size() * (c64.ok ? 1 : 0) + 15) & ~size_t(15);
    m.blob_c64_bytes = (int)off;
    m.off_envtb = (int)off; off += (ev.tbc.size() * 2 + 15) & ~size_t(15);
    m.off_envcls = (int)off; off += (ev.cls.size() + 15) & ~size_t(15);
    m.blob_env_bytes = (int)off;
    std::vector<unsigned char> blob(off, 0);
    uint32_t* cls = reinterpret_cast<uint32_t*>(blob.data() + off_cls);
    for (int c = 0; c < m.nCls; ++c) {
      const int orb = d->cls_orbit[c];
      const int* st = d->cls_stride + c * 4;
      for (int i = 0; i < 4; ++i)
        if (st[i] < 0 || st[i] > 255) {
          lmc_model_destroy(mdl);
          return fail("flat tensor strides above 255 are not supported (u8 strides for dp4a)");
        }
      cls[c * 4 + 0] = (uint32_t)st[0] | ((uint32_t)st[1] << 8) | ((uint32_t)st[2] << 16) | ((uint32_t)st[3] << 24);
      cls[c * 4 + 1] = (uint32_t)orbs[orb].atab_off;
      const double coef = kone ? d->natural_parameters[orbs[orb].fidx] * orbs[orb].w : 1.0;
      memcpy(&cls[c * 4 + 2], &coef, 8);
    }
    memcpy(blob.data() + m.off_tabA, tabA.data(), (size_t)tabA_len * 8);
    memcpy(blob.data() + m.off_nat, d->natural_parameters, (size_t)m.F * 8);
    memcpy(blob.data() + m.off_orb, orbs.data(), (size_t)m.nOrb * sizeof(OrbDev));
    if (!qtab.empty()) memcpy(blob.data() + m.off_qtab, qtab.data(), qtab.size() * 8);
    if (!dtab.empty()) memcpy(blob.data() + m.off_dtab, dtab.data(), dtab.size() * 8);
    if (c64.ok) {
      memcpy(blob.data() + m.off_c64desc, c64.desc.data(), c64.desc.size() * 4);
      memcpy(blob.data() + m.off_c64cls, c64.cls.data(), c64.cls.size());
    }
    if (ev.ncls) {
      memcpy(blob.data() + m.off_envtb, ev.tbc.data(), ev.tbc.size() * 2);
      memcpy(blob.data() + m.off_envcls, ev.cls.data(), ev.cls.size());
    }
    UP(unsigned char, blob.data(), blob.size(), m.blob);
    if (m.spOK) UP(unsigned char, sprec.data(), sprec.size(), m.sp_rec);
    m.spFtab = nullptr;
    if (m.spOK && !sp.ftab.empty()) UP(double, sp.ftab.data(), sp.ftab.size(), m.spFtab);
    m.c64OK = c64.ok; m.c64B = c64.b; m.c64NRL = c64.nrl; m.c64NRLP = c64.nrlp; m.c64NA = c64.NA; m.c64RV = c64.RV;
    m.c64NCls = c64.ncls; m.c64Bits = c64.bits;
    m.c64Rev = nullptr; m.c64Pair = nullptr;
    if (c64.ok) {
      UP(uint32_t, c64.rev.data(), c64.rev.size(), m.c64Rev);
      UP(unsigned long long, c64.pair.data(), c64.pair.size(), m.c64Pair);
    }
    m.envNCls = ev.ncls;
    m.envOK = ev.ok; m.envB = ev.b; m.envNRL = ev.nrl; m.envNRLP = ev.nrlp; m.envWide = ev.wide; m.envNA = ev.NA;
    m.envRV = ev.RV;
    m.envTb = nullptr; m.envRev = nullptr; m.envPair = nullptr;
    if (ev.ok) {
      UP(uint16_t, ev.tb.data(), ev.tb.size(), m.envTb);
      UP(uint32_t, ev.rev.data(), ev.rev.size(), m.envRev);
      if (ev.pair_ok) UP(unsigned char, ev.pair.data(), ev.pair.size(), m.envPair);
    }
  }
  UP(OrbDev, orbs.data(), orbs.size(), mdl->orb_dev);
  UP(double, d->natural_parameters, m.F, mdl->nat_dev);
  UP(double, d->ftab, d->ftab_len, m.ftab);
  // records
  {
    int rmax = 0;
    for (int i = 0; i < m.N; ++i) rmax = std::max(rmax, (int)(d->site_rec_off[i + 1] - d->site_rec_off[i]));
    m.Rstride = std::max(32, (rmax + 31) & ~31);  // multiple of every group size
    // fixed-stride table: site s owns records [s*Rstride, (s+1)*Rstride); pads point at site 0
    // with the zero class so that they contribute exactly 0
    std::vector<uint16_t> rec((size_t)m.N * m.Rstride * 4, 0);
    for (int i = 0; i < m.N; ++i) {
      const int64_t a = d->site_rec_off[i], b = d->site_rec_off[i + 1];
      uint16_t* dst = rec.data() + (size_t)i * m.Rstride * 4;
      memcpy(dst, d->site_rec + a * 4, (size_t)(b - a) * 8);
      for (int r = (int)(b - a); r < m.Rstride; ++r) dst[r * 4 + 3] = (uint16_t)m.nCls;
    }
    UP(uint2, rec.data(), (size_t)m.N * m.Rstride, m.site_rec);
    // orbit segments of a site -> lane entries (first, count, orbit, pieces that follow | -1).  A segment longer
    // than SEG_PIECE clusters is cut into up to four pieces on adjacent lanes which flip_features merges with
    // shuffles; a run of pieces never crosses a multiple of four lanes (first fit over 4-lane blocks), so it
    // stays inside one group for every group size.
    constexpr int SEG_PIECE = LMC_SEG_PIECE;
    std::vector<std::vector<int4>> site_entries(m.N);
    int smax = 4;
    for (int i = 0; i < m.N; ++i) {
      struct Run { int first, count, orbit, pieces; };
      std::vector<Run> runs;
      for (int64_t q = d->site_seg_off[i]; q < d->site_seg_off[i + 1]; ++q) {
        const int cnt = d->site_seg[q * 3 + 1];
        if (cnt <= 0) continue;
        runs.push_back({d->site_seg[q * 3], cnt, d->site_seg[q * 3 + 2], std::min(4, (cnt + SEG_PIECE - 1) / SEG_PIECE)});
      }
      std::stable_sort(runs.begin(), runs.end(), [](const Run& x, const Run& y) { return x.pieces > y.pieces; });
      std::vector<int> used;   // lanes taken in each 4-lane block
      std::vector<int4>& ent = site_entries[i];
      for (const Run& r : runs) {
        size_t b = 0;
        while (b < used.size() && used[b] + r.pieces > 4) ++b;
        if (b == used.size()) { used.push_back(0); ent.resize(4 * used.size(), make_int4(0, 0, 0, -1)); }
        const int len = (r.count + r.pieces - 1) / r.pieces;
        for (int k = 0, at = 0; k < r.pieces; ++k, at += len)
          ent[4 * b + used[b] + k] = make_int4(r.first + at, std::max(0, std::min(len, r.count - at)), r.orbit,
                                               k == 0 ? r.pieces - 1 : -1);
        used[b] += r.pieces;
      }
      smax = std::max(smax, (int)ent.size());
    }
    m.Sstride = smax;
    std::vector<int4> segs((size_t)m.N * smax, make_int4(0, 0, 0, -1));
    for (int i = 0; i < m.N; ++i) std::copy(site_entries[i].begin(), site_entries[i].end(), segs.begin() + (size_t)i * smax);
    UP(int4, segs.data(), segs.size(), m.site_seg);
    UP(uint2, d->full_rows, d->orb_row_off[m.nOrb], m.full_rows);
  }
  m.muW = d->mu_width;
  if (m.muW > 0) {
    m.muF = d->mu_feature;
    UP(double, d->mu_table, (size_t)m.N * m.muW, m.mu);
  }
  // sublattices
  {
    double cum = 0.0;
    for (int s = 0; s < m.nSl; ++s) {
      m.sl_off[s] = d->sl_site_off[s];
      m.sl_ncodes[s] = d->sl_ncodes[s];
      if (m.sl_ncodes[s] > LMC_MAX_CODES) { lmc_model_destroy(mdl); return fail("too many species on a sublattice"); }
      for (int c = 0; c < LMC_MAX_CODES; ++c) {
        m.sl_codes[s][c] = d->sl_codes[s * LMC_MAX_CODES + c];
        if (c < m.sl_ncodes[s] && (m.sl_codes[s][c] < 0 || m.sl_codes[s][c] >= LMC_MAX_CODES)) {
          lmc_model_destroy(mdl);
          return fail("species codes must be < 8");
        }
      }
      for (int code = 0; code < LMC_MAX_CODES; ++code) {
        m.sl_code_pos[s][code] = (unsigned char)m.sl_ncodes[s];
        for (int c = m.sl_ncodes[s] - 1; c >= 0; --c) if (m.sl_codes[s][c] == code) m.sl_code_pos[s][code] = (unsigned char)c;
      }
      cum += d->sl_prob[s];
      m.sl_cum[s] = (s == m.nSl - 1) ? 1.0 : cum;
      const int a = d->sl_site_off[s], b = d->sl_site_off[s + 1];
      bool contig = b > a;
      for (int j = a + 1; j < b; ++j) contig = contig && d->sl_sites[j] == d->sl_sites[j - 1] + 1;
      m.sl_first[s] = contig ? d->sl_sites[a] : -1;
    }
    m.sl_off[m.nSl] = d->sl_site_off[m.nSl];
    // per-sublattice chemical-potential / (charge, diagonal) tables (see DevModel): bitwise identical rows only
    m.muC = m.muW > 0 && m.muW <= LMC_MAX_CODES;
    m.qdC = m.E > 0 && m.ewK != nullptr && m.ewW <= LMC_MAX_CODES;
    memset(m.mu_c, 0, sizeof(m.mu_c));
    memset(m.qc_c, 0, sizeof(m.qc_c));
    memset(m.qg_c, 0, sizeof(m.qg_c));
    for (int s = 0; s < m.nSl; ++s) {
      const int a = d->sl_site_off[s], b = d->sl_site_off[s + 1];
      if (b <= a) continue;
      const int s0 = d->sl_sites[a];
      for (int j = a; j < b; ++j) {
        const int k = d->sl_sites[j];
        if (m.muC)
          for (int c = 0; c < m.muW; ++c)
            if (memcmp(&d->mu_table[(size_t)k * m.muW + c], &d->mu_table[(size_t)s0 * m.muW + c], 8) != 0) m.muC = 0;
        if (m.qdC)
          for (int c = 0; c < m.ewW; ++c)
            if (memcmp(&qd_host[(size_t)k * m.ewW + c], &qd_host[(size_t)s0 * m.ewW + c], 16) != 0) m.qdC = 0;
      }
      for (int c = 0; c < LMC_MAX_CODES; ++c) {
        if (m.muC && c < m.muW) m.mu_c[s][c] = d->mu_table[(size_t)s0 * m.muW + c];
        if (m.qdC && c < m.ewW) { m.qc_c[s][c] = qd_host[(size_t)s0 * m.ewW + c].x; m.qg_c[s][c] = qd_host[(size_t)s0 * m.ewW + c].y; }
      }
    }
    if (getenv("LMC_COMPACT_TABLES_OFF")) m.muC = m.qdC = 0;
    {
      double ct[3 * LMC_MAX_SUBLATTICES * LMC_MAX_CODES];
      memcpy(ct, m.mu_c, sizeof(m.mu_c));
      memcpy(ct + LMC_MAX_SUBLATTICES * LMC_MAX_CODES, m.qc_c, sizeof(m.qc_c));
      memcpy(ct + 2 * LMC_MAX_SUBLATTICES * LMC_MAX_CODES, m.qg_c, sizeof(m.qg_c));
      cudaError_t e = cudaMemcpy(const_cast<unsigned char*>(m.blob) + m.off_ctab, ct, sizeof(ct), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { lmc_model_destroy(mdl); return fail(std::string("cudaMemcpy: ") + cudaGetErrorString(e)); }
    }
    int pw = 0, le = 0;
    for (int s = 0; s < m.nSl; ++s) {
      int maxcode = 0;
      for (int c = 0; c < m.sl_ncodes[s]; ++c) maxcode = std::max(maxcode, m.sl_codes[s][c]);
      m.sl_nwords[s] = (m.sl_off[s + 1] - m.sl_off[s] + 31) / 32;
      m.sl_plane_off[s] = pw;
      pw += (maxcode + 1) * m.sl_nwords[s];
      m.sl_nplanes[s] = maxcode + 1;
      m.sl_list_off[s] = le;
      le += (maxcode + 1) * (m.sl_off[s + 1] - m.sl_off[s]);
    }
    m.plane_words = pw;
    m.list_entries = le;
    UP(int, d->sl_sites, d->sl_site_off[m.nSl], m.sl_sites);
  }
  m.tfD = d->tf_num_flips > 0 ? d->tf_num_dims : 0;
  m.tfNF = d->tf_num_flips;
  for (int i = 0; i < m.tfNF; ++i)
    for (int k = 0; k < m.tfD; ++k) m.tf_table[i][k] = d->tf_table[i * m.tfD + k];
  for (int i = 0; i < 2 * m.tfNF; ++i) m.tf_w[i] = d->tf_weights[i];
  for (int k = 0; k < m.tfD; ++k) {
    m.tf_max_n[k] = d->tf_max_n[k];
    m.tf_dim_sl[k] = d->tf_dim_sl[k];
    m.tf_dim_code[k] = d->tf_dim_code[k];
  }
  m.tf_sw = d->tf_swap_weight;
  memset(m.tf_npick, 0, sizeof(m.tf_npick));
  memset(m.tf_pick, 0, sizeof(m.tf_pick));
  for (int i = 0; i < 2 * m.tfNF; ++i) {
    const int sgn = (i & 1) ? -1 : 1;
    int n = 0;
    bool fits = true;
    int d0 = 0;
    while (d0 < m.tfD) {   // dims of one sublattice are consecutive (occu_utils.py:20-25)
      const int sl = m.tf_dim_sl[d0];
      int d1 = d0 + 1;
      while (d1 < m.tfD && m.tf_dim_sl[d1] == sl) ++d1;
      if (sl >= 0) {
        bool first = true;
        for (int pass = 0; pass < 2; ++pass)
          for (int dd = d0; dd < d1; ++dd) {
            const int ud = sgn * m.tf_table[i >> 1][dd];
            const int cnt = pass == 0 ? -ud : ud;
            for (int p = 0; p < cnt; ++p) {
              if (n >= 2 * LMC_MAX_FLIPS) { fits = false; break; }
              m.tf_pick[i][n++] = (uint32_t)pass | ((uint32_t)sl << 1) | ((uint32_t)m.tf_dim_code[dd] << 4) | ((uint32_t)dd << 8) |
                                  ((uint32_t)p << 12) | ((first ? 1u : 0u) << 16);
              first = false;
            }
          }
      }
      d0 = d1;
    }
    m.tf_npick[i] = fits ? n : -1;   // -1: more picks than descriptors (the flip table changes more than 4 sites: refused at lmc_run)
  }
  {
    int maxn = 1;
    for (int k = 0; k < m.tfD; ++k) maxn = std::max(maxn, m.tf_max_n[k]);
    std::vector<double> lg(maxn + 2);
    for (int n = 0; n < maxn + 2; ++n) lg[n] = lgamma((double)n + 1.0);   // gammaln(n + 1), mcusher.py:705-709
    UP(double, lg.data(), lg.size(), m.lgam);
  }
#undef UP
  *out = mdl;
  return 0;
}

extern "C" int lmc_model_info(const LmcModel* mdl, int32_t* info, int n) {
  if (!mdl || !info) return fail("null argument");
  const DevModel& m = mdl->dm;
  const int32_t v[6] = {m.E > 0 && m.ewK != nullptr, m.spOK, m.blob_bytes, m.spNQ,
                        m.envOK ? m.envNA * (m.envWide ? 32 : 16) : 0, m.c64OK ? m.c64Bits : 0};
  for (int i = 0; i < n && i < 6; ++i) info[i] = v[i];
  return 0;
}

extern "C" int lmc_ewald_field(const LmcModel* mdl, const int8_t* occ, int W, double* field, void* stream) {
  if (!mdl || !occ || !field) return fail("null argument");
  if (W <= 0) return 0;
  const DevModel& m = mdl->dm;
  if (m.E <= 0 || !m.ewK) return fail("the Ewald potential cache needs an Ewald matrix of the form q_i q_j K[site_i, site_j]");
  const size_t smem = (size_t)FIELD_WPB * m.N * sizeof(double);
  // many walkers: the tiled product (every element of K read W / 128 times); few walkers or LMC_FIELD_TILED=0: one
  // warp per row of K, four walkers per block
  bool tiled = W >= 64;
  if (const char* e = getenv("LMC_FIELD_TILED")) tiled = atoi(e) != 0;
  if (tiled) {
    dim3 grid((m.N + FT_S - 1) / FT_S, (W + FT_W - 1) / FT_W);
    lmc_ewald_field_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m, occ, W, field);
  } else {
    if ((int)smem > mdl->smem_optin) return fail("too many sites for the Ewald potential cache kernel");
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(lmc_ewald_field_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lmc_ewald_field_kernel<<<(W + FIELD_WPB - 1) / FIELD_WPB, 256, smem, (cudaStream_t)stream>>>(m, occ, W, field);
  }
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int lmc_full_features_field(const LmcModel* mdl, const int8_t* occ, int W, double* feat, double* enth,
                                       double* field, void* stream) {
  if (!mdl) return fail("null model");
  if (W <= 0) return 0;
  const DevModel& m = mdl->dm;
  if (!field || m.E <= 0 || !m.ewK) return lmc_full_features(mdl, occ, W, feat, enth, stream);
  const int rc = lmc_ewald_field(mdl, occ, W, field, stream);
  if (rc != 0) return rc;
  const size_t smem = (size_t)m.Npad + (size_t)m.N * 4 + 16;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(lmc_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lmc_full_kernel<<<W, 256, smem, (cudaStream_t)stream>>>(m, occ, feat, enth, mdl->orb_dev, mdl->nat_dev, field);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int lmc_bias_init(const int8_t* occ, int W, int N, int mode, int bw, int rows, double pen, const double* icpt,
                             const double* tab, double* bias, double* sum, void* stream) {
  if (!occ || !tab || !bias || !sum) return fail("null argument");
  if (mode != LMC_BIAS_TABLE_SUM && mode != LMC_BIAS_SQUARE_SUM) return fail("unknown bias mode");
  if (rows < 1 || rows > LMC_MAX_BIAS_ROWS || (mode == LMC_BIAS_TABLE_SUM && rows != 1)) return fail("bias_rows out of range");
  if (W <= 0) return 0;
  BiasIcpt ic;
  for (int r = 0; r < LMC_MAX_BIAS_ROWS; ++r) ic.v[r] = (icpt && r < rows) ? icpt[r] : 0.0;
  lmc_bias_init_kernel<<<(W + 3) / 4, 128, 0, (cudaStream_t)stream>>>(occ, W, N, lmc_row_stride(N), mode, bw, rows, pen, ic, tab,
                                                                    bias, sum);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int lmc_distance_init(const LmcModel* mdl, int W, double* feat, double* vec, double* enth, const double* target,
                                 double tol, int ngrp, const int32_t* goff, const int32_t* gidx, const double* gdiam,
                                 void* stream) {
  if (!mdl || !feat || !vec || !target || !goff || !gidx || !gdiam) return fail("null argument");
  if (W <= 0) return 0;
  const DevModel& m = mdl->dm;
  lmc_distance_init_kernel<<<(W + 127) / 128, 128, 0, (cudaStream_t)stream>>>(W, m.F, m.size, mdl->nat_dev, feat, vec, enth, target,
                                                                              tol, ngrp, goff, gidx, gdiam);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int lmc_ewald_site_kernel(const double* cart, int nsites, const int32_t* origins, int norig, const double* gv,
                                     const double* gc, int ng, const double* tv, int nt, double eta, double rcut, double vol,
                                     double* out, void* stream) {
  if (!cart || !origins || !gv || !gc || !tv || !out) return fail("null argument");
  if (nsites <= 0 || norig <= 0) return 0;
  if (norig > 65535) return fail("too many origin sites");
  lmc_ewald_site_kernel_k<<<dim3((unsigned)nsites, (unsigned)norig), 256, 0, (cudaStream_t)stream>>>(
      cart, origins, gv, gc, ng, tv, nt, eta, rcut, vol, nsites, out);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int lmc_cast_i32_to_i8(const int32_t* src, int8_t* dst, int W, int N, void* stream) {
  const int Npad = lmc_row_stride(N);
  const long long n = (long long)W * Npad;
  if (n == 0) return 0;
  lmc_cast_i32_i8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, W, N, Npad);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}
extern "C" int lmc_cast_i8_to_i32(const int8_t* src, int32_t* dst, int64_t rows, int N, int stride, void* stream) {
  const long long n = rows * N;
  if (n == 0) return 0;
  lmc_cast_i8_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, N, stride);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int lmc_full_features(const LmcModel* mdl, const int8_t* occ, int W, double* feat, double* enth, void* stream) {
  if (!mdl) return fail("null model");
  if (W <= 0) return 0;
  const DevModel& m = mdl->dm;
  const size_t smem = (size_t)m.Npad + (m.E ? (size_t)m.N * 4 : 0) + 16;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(lmc_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lmc_full_kernel<<<W, 256, smem, (cudaStream_t)stream>>>(m, occ, feat, enth, mdl->orb_dev, mdl->nat_dev, nullptr);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

static size_t delta_walker_smem(const DevModel& m) {
  return (((size_t)m.F * 8 + 15) & ~size_t(15)) + (((size_t)m.Rstride * 8 + 15) & ~size_t(15)) +
         (m.E ? (((size_t)m.N * 2 + 15) & ~size_t(15)) : 0);
}

extern "C" int lmc_delta_features(const LmcModel* mdl, const int8_t* occ, int W, const int32_t* sites, const int32_t* codes,
                                  int nflips, double* out, void* stream) {
  if (!mdl) return fail("null model");
  if (W <= 0) return 0;
  const DevModel& m = mdl->dm;
  const int G = 32, threads = 128, wpb = threads / G;
  const size_t wsm = delta_walker_smem(m);
  const size_t smem = (((size_t)m.off_dtab + 15) & ~size_t(15)) + (size_t)wpb * (m.Npad + wsm);
  if ((int)smem > mdl->smem_optin) return fail("model tables do not fit in shared memory");
  const int grid = (W + wpb - 1) / wpb;
  if (m.kone) {
    CK(cudaFuncSetAttribute(lmc_delta_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lmc_delta_kernel<32, true><<<grid, threads, smem, (cudaStream_t)stream>>>(m, occ, W, sites, codes, nflips, out, wpb, (int)wsm);
  } else {
    CK(cudaFuncSetAttribute(lmc_delta_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lmc_delta_kernel<32, false><<<grid, threads, smem, (cudaStream_t)stream>>>(m, occ, W, sites, codes, nflips, out, wpb, (int)wsm);
  }
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int lmc_run(const LmcModel* mdl, const LmcRunConfig* c, void* stream) {
  if (!mdl || !c) return fail("null argument");
  const DevModel& m = mdl->dm;
  if (c->num_walkers <= 0 || c->num_samples <= 0) return 0;
  if (c->thin_by <= 0) return fail("thin_by must be positive");
  if (c->usher == LMC_USHER_TABLEFLIP && m.tfNF == 0) return fail("model has no flip table");
  if (c->usher == LMC_USHER_COMPOSITE) {
    if (c->comp_num < 1 || c->comp_num > LMC_MAX_COMPOSITE) return fail("a composite usher takes 1..4 sub-ushers");
    for (int i = 0; i < c->comp_num; ++i)
      if (c->comp_usher[i] != LMC_USHER_FLIP && c->comp_usher[i] != LMC_USHER_SWAP)
        return fail("composite sub-ushers must be Flip or Swap");
  }
  if (c->usher == LMC_USHER_MULTISTEP) {
    if (c->ms_usher != LMC_USHER_FLIP && c->ms_usher != LMC_USHER_SWAP) return fail("the multi-step sub-usher must be Flip or Swap");
    if (c->ms_num < 1 || c->ms_num > LMC_MAX_COMPOSITE) return fail("a multi-step usher takes 1..4 step lengths");
    for (int i = 0; i < c->ms_num; ++i)
      if (c->ms_len[i] < 1 || c->ms_len[i] * (c->ms_usher == LMC_USHER_SWAP ? 2 : 1) > LMC_MAX_FLIPS)
        return fail("multi-step lengths must change at most 4 sites per step (<= 4 flips or <= 2 swaps)");
  }
  if (c->usher < 0 || c->usher > LMC_USHER_MULTISTEP) return fail("unknown usher");
  if (c->kernel == LMC_KERNEL_WANGLANDAU && c->wl.num_bins <= 1) return fail("Wang-Landau needs more than one bin");
  const bool ewald = m.E > 0;
  const bool field = ewald && c->ewald_field_dev != nullptr;   // Ewald through the potential cache
  if (field && !m.ewK) return fail("ewald_field_dev needs an Ewald matrix of the form q_i q_j K[site_i, site_j]");
  int G = c->group_size;
  if (const char* e = getenv("LMC_GROUP_SIZE")) { if (G == 0) G = atoi(e); }
  // kernel selection: the speculative-batch kernel (lmc_spec.cuh) wins while fewer than ~1/3 of the
  // steps are accepted; the acceptance ratio comes from the totals of the launches that have completed
  LmcModel* mm = const_cast<LmcModel*>(mdl);
  if (mm->stats_host) {
    const unsigned long long acc = ((volatile unsigned long long*)mm->stats_host)[0];
    const unsigned long long att = ((volatile unsigned long long*)mm->stats_host)[1];
    if (att >= mm->snap[1] + 65536ull && acc >= mm->snap[0]) {
      mm->acc_rate = (double)(acc - mm->snap[0]) / (double)(att - mm->snap[1]);
      mm->snap[0] = acc; mm->snap[1] = att;
    }
  }
  int spec_mode = c->spec_mode;
  if (const char* e = getenv("LMC_SPEC")) { if (spec_mode == 0) spec_mode = atoi(e) ? 2 : 1; }
  if (c->bias_mode != LMC_BIAS_NONE) {
    if (c->kernel != LMC_KERNEL_METROPOLIS) return fail("bias terms are defined for the Metropolis kernel only (wanglandau.py:24)");
    if (c->bias_mode != LMC_BIAS_TABLE_SUM && c->bias_mode != LMC_BIAS_SQUARE_SUM) return fail("unknown bias mode");
    if (!c->bias_table_dev || !c->bias_dev || !c->bias_sum_dev || c->bias_width <= 0) return fail("bias pointers must not be null");
    if (c->bias_rows < 1 || c->bias_rows > LMC_MAX_BIAS_ROWS || (c->bias_mode == LMC_BIAS_TABLE_SUM && c->bias_rows != 1))
      return fail("bias_rows out of range");
  }
  const bool dist = c->dist_mode != 0;
  if (dist) {
    if (c->kernel != LMC_KERNEL_METROPOLIS || (c->usher != LMC_USHER_FLIP && c->usher != LMC_USHER_SWAP) || ewald || m.muW)
      return fail("distance processors run Metropolis flip / swap steps without Ewald or chemical-potential terms");
    if (!c->dist_target_dev || !c->dist_group_off_dev || !c->dist_group_idx_dev || !c->dist_group_diam_dev || !c->dist_vector_dev)
      return fail("distance processor pointers must not be null");
    if (c->bias_mode != LMC_BIAS_NONE) return fail("bias terms are not combined with distance processors");
  }
  const bool multicell = c->walker_mask_dev != nullptr || c->accept_offset_dev != nullptr;
  if (multicell && c->kernel != LMC_KERNEL_METROPOLIS) return fail("walker_mask_dev / accept_offset_dev are for Metropolis kernels");
  const bool spec_ok = m.spOK && !dist && !multicell && c->bias_mode == LMC_BIAS_NONE && (!ewald || field) && c->kernel == LMC_KERNEL_METROPOLIS &&
                       (c->usher == LMC_USHER_FLIP || c->usher == LMC_USHER_SWAP) && (G == 0 || G == 32);
  // table flips: speculative batches of their own (lmc_spec_tf.cuh), one block of up to 16 walkers per SM
  const bool tf_spec_ok = m.spOK && m.spNQ % 8 == 0 && c->usher == LMC_USHER_TABLEFLIP && c->kernel == LMC_KERNEL_METROPOLIS && !dist &&
                          !multicell && c->bias_mode == LMC_BIAS_NONE && (!ewald || field) && (G == 0 || G == 32) && c->block_threads == 0;
  if (spec_mode == 2 && !spec_ok && !tf_spec_ok)
    return fail("the speculative kernels support unbiased Metropolis flip / swap / table-flip steps (an Ewald term through ewald_field_dev only)");
  bool tf_spec = tf_spec_ok && (spec_mode == 2 || ((spec_mode == 3 || (spec_mode == 0 && mm->acc_rate < 0.35)) && G == 0));
  if (const char* e = getenv("LMC_SPEC_TF")) tf_spec = tf_spec && atoi(e) != 0;
  // auto: only while the staged tables leave room for a full complement of resident walkers per SM
  // 3 = speculative whenever this model / run supports it (decided by the caller from ITS acceptance history:
  // deterministic, unlike the asynchronously refreshed acc_rate of mode 0)
  const bool use_spec = spec_ok && (spec_mode == 2 || ((spec_mode == 3 || (spec_mode == 0 && mm->acc_rate < 0.35)) && G == 0 && m.off_ctab <= 40 * 1024));
  if (use_spec) G = 32;
  // lanes per speculated step (lmc_spec.cuh).  Four everywhere: with two or one lane per step (one uses
  // sorted position lists for the swap partner) the scalar work per step shrinks, but 16 / 32 unrelated
  // flip sites per warp instruction multiply the L1 / shared-memory wavefronts of the record loads and
  // occupancy gathers, which is the pipe that bounds this kernel (measured, profiles/r01_variants.md)
  int spec_sg = 4;
  if (const char* e = getenv("LMC_SPEC_SG")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) spec_sg = v; }
  // swap partner from sorted position lists instead of the rank select: measured neutral with four lanes per
  // step (the kernel is bound by the L1 / shared-memory data pipe, not by issue slots; profiles/r01_variants.md),
  // so only the one-lane variant and LMC_SPEC_LISTS=1 use them
  bool spec_lists = use_spec && c->usher == LMC_USHER_SWAP && !field && spec_sg == 1;
  if (const char* e = getenv("LMC_SPEC_LISTS")) spec_lists = use_spec && c->usher == LMC_USHER_SWAP && !field && spec_sg != 2 && (atoi(e) != 0 || spec_sg == 1);
  {
    const size_t per_walker = (size_t)m.Npad + 4096 + (size_t)m.list_entries * 2;   // generous bound of the slab
    if (spec_lists && spec_sg == 4 && 7 * ((((size_t)m.blob_bytes + 15) & ~size_t(15)) + 4 * per_walker + 1024) > 227 * 1024) spec_lists = false;
  }
  if (G == 0) {
    // measured on B200 (profiles/): a full warp per walker wins while all walkers fit in one wave
    // (W <= 32 per SM); beyond that half warps amortise the per-step scalar work better
    if (ewald || c->num_walkers <= mdl->num_sms * 32) G = 32;
    else if (c->num_walkers <= mdl->num_sms * 1024) G = 16;
    else G = 8;
  }
  if (c->usher == LMC_USHER_TABLEFLIP) G = (G >= 16) ? 32 : 8;
  if (c->usher == LMC_USHER_COMPOSITE || c->usher == LMC_USHER_MULTISTEP || dist) G = 32;
  if (G != 4 && G != 8 && G != 16 && G != 32) return fail("group_size must be 4, 8, 16 or 32");
  int threads = c->block_threads;
  if (const char* e = getenv("LMC_BLOCK_THREADS")) { if (threads == 0) threads = atoi(e); }
  const bool relaxed = !use_spec && (ewald || c->kernel == LMC_KERNEL_WANGLANDAU || c->usher >= LMC_USHER_TABLEFLIP);
  const int max_threads = relaxed ? 256 : 128;
  const bool auto_threads = threads == 0;
  if (threads == 0) threads = 128;
  if (threads % 32 || threads > max_threads) return fail("block_threads must be a multiple of 32 (<= 128, or <= 256 for Ewald / Wang-Landau / table-flip kernels)");

  RunArgs a;
  memset(&a, 0, sizeof(a));
  a.W = c->num_walkers; a.walker_base = c->walker_id_base; a.usher = c->usher; a.kernel = c->kernel;
  a.S = c->num_samples; a.thin = c->thin_by; a.step0 = c->step_begin;
  a.seeds = reinterpret_cast<const unsigned long long*>(c->seeds_dev);
  a.beta = c->beta_dev;
  a.occ = c->occ_dev; a.features = c->features_dev; a.enthalpy = c->enthalpy_dev;
  a.tr_occ = c->trace_occ_dev; a.tr_feat = c->trace_features_dev; a.tr_enth = c->trace_enthalpy_dev;
  a.tr_acc = c->trace_accepted_dev; a.tr_nacc = c->trace_naccepted_dev;
  a.wl = c->wl;
  a.bias_mode = c->bias_mode; a.bias_w = c->bias_width; a.bias_rows = c->bias_rows; a.bias_pen = c->bias_penalty;
  a.bias_tab = c->bias_table_dev; a.bias = c->bias_dev; a.bias_sum = c->bias_sum_dev; a.tr_bias = c->trace_bias_dev;
  a.comp_num = c->usher == LMC_USHER_COMPOSITE ? c->comp_num : 0;
  for (int i = 0; i < LMC_MAX_COMPOSITE; ++i) {
    a.comp_usher[i] = c->comp_usher[i];
    a.comp_cum[i] = c->comp_cum[i];
    for (int k = 0; k < LMC_MAX_SUBLATTICES; ++k) a.comp_sl_cum[i][k] = c->comp_sl_cum[i][k];
  }
  a.ms_usher = c->ms_usher; a.ms_num = c->usher == LMC_USHER_MULTISTEP ? c->ms_num : 0;
  for (int i = 0; i < LMC_MAX_COMPOSITE; ++i) { a.ms_len[i] = c->ms_len[i]; a.ms_cum[i] = c->ms_cum[i]; }
  a.stats = multicell ? nullptr : mm->stats_dev;   // (masked launches would skew the acceptance feedback)
  a.mask = c->walker_mask_dev; a.acc_off = c->accept_offset_dev;
  if (!a.seeds || !a.occ || !a.features || !a.enthalpy) return fail("state pointers must not be null");
  if (c->kernel != LMC_KERNEL_WANGLANDAU && !a.beta) return fail("beta_dev must not be null");
  // per-walker shared-memory slab: [features][stash x MAX_FLIPS][counts]
  const size_t stash_el = m.kone ? 8 : 4;
  a.off_feat = 0;
  a.off_stash = (int)(((size_t)m.F * 8 + 15) & ~size_t(15));
  a.max_flips = c->usher == LMC_USHER_FLIP ? 1 : (c->usher == LMC_USHER_MULTISTEP ? LMC_MAX_FLIPS : 2);
  if (c->usher == LMC_USHER_TABLEFLIP)
    for (int i = 0; i < m.tfNF; ++i) {
      int up = 0, dn = 0;
      for (int k = 0; k < m.tfD; ++k) { up += std::max(m.tf_table[i][k], 0); dn += std::max(-m.tf_table[i][k], 0); }
      a.max_flips = std::max(a.max_flips, std::max(up, dn));
    }
  if (const char* e = getenv("LMC_SEQUENTIAL_FLIPS")) a.seq_flips = atoi(e);
  if (a.max_flips > LMC_MAX_FLIPS) return fail("flip table changes more than 4 sites per step");
  if (tf_spec) {
    // one stash slot (the commit evaluates and folds flip by flip) and two rings of 32 random-word blocks; the
    // variant is taken when at least four walkers and the whole table blob fit a block
    const size_t slab = (size_t)m.Npad + (((size_t)m.F * 8 + 15) & ~size_t(15)) +
                        std::max<size_t>(((size_t)m.Rstride * stash_el + 15) & ~size_t(15), 2 * 32 * 16) +
                        LMC_MAX_SUBLATTICES * LMC_MAX_CODES * 4 + ((m.plane_words * 4 + 15) & ~15) + ((m.plane_words * 2 + 15) & ~15) +
                        (6 * LMC_MAX_TABLE_FLIPS + 2) * 8;
    if ((((size_t)m.blob_bytes + 15) & ~size_t(15)) + 4 * slab > (size_t)mdl->smem_optin - 1024) {
      if (spec_mode == 2) return fail("the speculative table-flip kernel does not fit this model in shared memory");
      tf_spec = false;
    }
  }
  if (tf_spec) G = 32;
  // compact environment words in shared memory (lmc_spec_c64.cuh): the default of the speculative flip / swap kernel
  // where the model has the tables and 28 walkers per SM (seven blocks of four) still fit beside their words
  bool spec_c64 = use_spec && m.c64OK && !field && spec_sg == 4 && !spec_lists && c->block_threads == 0 && !getenv("LMC_SPEC_WIDE") &&
                  c->spec_env_dev == nullptr;   // (a caller that hands in the L2 workspace asks for that variant)
  if (const char* e = getenv("LMC_SPEC_C64")) spec_c64 = spec_c64 && atoi(e) != 0;
  if (spec_c64) {
    const size_t slab = (size_t)m.Npad + (((size_t)m.F * 8 + 15) & ~size_t(15)) +
                        std::max<size_t>(((size_t)m.Rstride * stash_el + 15) & ~size_t(15), 32 * 16) +
                        LMC_MAX_SUBLATTICES * LMC_MAX_CODES * 4 + ((m.plane_words * 4 + 15) & ~15) + ((m.plane_words * 2 + 15) & ~15) +
                        (size_t)m.c64NA * 8;
    if (7 * ((((size_t)m.blob_c64_bytes + 15) & ~size_t(15)) + 4 * slab + 1024) > 227 * 1024) spec_c64 = false;
  }
  const bool one_slot = tf_spec || spec_c64;   // one stash slot (the commit evaluates and folds flip by flip) that also holds the ring(s)
  const int stash_slots = one_slot ? 1 : a.max_flips;
  size_t stash_bytes = ((size_t)stash_slots * m.Rstride * stash_el + 15) & ~size_t(15);
  if (tf_spec) stash_bytes = std::max<size_t>(stash_bytes, 2 * 32 * 16);   // the two random-word rings live in the idle stash
  if (spec_c64) stash_bytes = std::max<size_t>(stash_bytes, 32 * 16);
  a.off_cnt = a.off_stash + (int)stash_bytes;
  a.off_plane = a.off_cnt + LMC_MAX_SUBLATTICES * LMC_MAX_CODES * 4;
  a.off_ring = a.off_plane + ((m.plane_words * 4 + 15) & ~15);   // species bit-planes
  a.off_eidx = a.off_ring + (one_slot ? 0 : G * 16);                  // per-lane precomputed proposals
  if (one_slot) a.off_ring = a.off_stash;
  a.ew_field = field ? c->ewald_field_dev : nullptr;
  a.off_lists = a.off_eidx + ((ewald && !field) ? (((m.ewK ? 1 : 2) * m.N + 15) & ~15) : 0);  // per-site Ewald cache (u8 charge index or u16 row)
  if (one_slot) a.off_lists = a.off_eidx + ((m.plane_words * 2 + 15) & ~15);   // (off_eidx: prefix popcounts of the plane words, u16)
  a.off_bias = a.off_lists + (spec_lists ? ((m.list_entries * 2 + 15) & ~15) : 0);   // sorted position lists
  a.off_dist = a.off_bias + (c->bias_mode != LMC_BIAS_NONE ? 16 * ((1 + LMC_MAX_BIAS_ROWS + 1) / 2) : 0);   // running bias value and table sums
  a.walker_smem = a.off_dist + (dist ? ((3 * m.F * 8 + 15) & ~15) : 0);   // distance processor: vector, delta, new distances
  a.off_tfc = a.walker_smem;
  if (c->usher == LMC_USHER_TABLEFLIP) a.walker_smem += (6 * LMC_MAX_TABLE_FLIPS + 2) * 8;
  // Wang-Landau flips: landing zone of the next step's records / segment entries
  a.off_pref = a.walker_smem;
  if (c->kernel == LMC_KERNEL_WANGLANDAU && c->usher == LMC_USHER_FLIP && !dist) a.walker_smem += 2 * (m.Rstride * 8 + m.Sstride * 16);   // double buffered
  // Wang-Landau: entropy + histogram of the walker next to its occupancy while they are small (<= 24 KB)
  a.off_wl = -1;
  if (c->kernel == LMC_KERNEL_WANGLANDAU && c->wl.num_bins > 0 && (size_t)c->wl.num_bins * 16 <= 24 * 1024 && !getenv("LMC_WL_GLOBAL")) {
    a.off_wl = a.walker_smem;
    a.walker_smem += (c->wl.num_bins * 16 + 15) & ~15;
  }
  a.off_env64 = a.walker_smem;
  if (spec_c64) a.walker_smem += (m.c64NA * 8 + 15) & ~15;
  a.dist_ngrp = c->dist_num_groups; a.dist_tol = c->dist_tol; a.dist_target = c->dist_target_dev;
  a.dist_grp_off = c->dist_group_off_dev; a.dist_grp_idx = c->dist_group_idx_dev; a.dist_grp_diam = c->dist_group_diam_dev;
  a.dist_vec = c->dist_vector_dev;
  // staged tables: the speculative kernel takes the whole blob, the classic kernels stop before its difference table
  // environment words: caller's workspace + tables built + the default four-lane rank-select variant
  bool spec_env = use_spec && c->spec_env_dev != nullptr && m.envOK && spec_sg == 4 && !spec_lists &&
                  (c->usher == LMC_USHER_FLIP || m.envPair != nullptr);
  if (const char* e = getenv("LMC_SPEC_ENV")) spec_env = spec_env && atoi(e) != 0;
  a.env = spec_env ? reinterpret_cast<uint32_t*>(c->spec_env_dev) : nullptr;
  if (spec_c64) spec_env = false;
  const size_t blob = ((size_t)(use_spec ? (spec_c64 ? m.blob_c64_bytes : (spec_env ? m.blob_env_bytes : m.blob_bytes)) : m.off_dtab) + 15) & ~size_t(15);
  size_t smem = 0;
  if (auto_threads && relaxed) {
    // shared memory limits residency here: take the block size with the most resident walkers per SM
    // (fewest waves); 16 warps/SM is the register ceiling of these variants
    const size_t sm_smem = 227 * 1024;
    // threads per SM the register budget of these variants allows: 512 at <= 128 registers; the table-flip + Ewald
    // variants are compiled for LMC_TF_MINB resident 256-thread blocks
    int reg_threads_sm = 512;
    if (c->usher == LMC_USHER_TABLEFLIP && ewald && c->kernel == LMC_KERNEL_METROPOLIS) reg_threads_sm = 256 * LMC_TF_MINB;
    int best_t = 0; long best_w = -1;
    for (int t = G; t <= max_threads; t += G) {
      if (t % 32) continue;
      const size_t bs = blob + (size_t)(t / G) * (m.Npad + a.walker_smem);
      if ((int)bs > mdl->smem_optin - 1024) break;
      long blocks = (long)(sm_smem / (bs + 1024));
      blocks = std::min(blocks, (long)(reg_threads_sm / t));
      const long wsm = blocks * (t / G);
      if (wsm > best_w) { best_w = wsm; best_t = t; }
    }
    if (best_t) threads = best_t;
  }
  bool spec_wide = false;
  if (use_spec && auto_threads) {
    // seven blocks of four walkers per SM while they fit; with a larger table blob two blocks of fourteen
    // walkers (448 threads) keep the same 28 walkers resident
    const size_t b4 = blob + 4 * (size_t)(m.Npad + a.walker_smem) + 1024, b14 = blob + 14 * (size_t)(m.Npad + a.walker_smem) + 1024;
    if (!spec_lists && 7 * b4 > 227 * 1024 && 2 * b14 <= 227 * 1024 && (int)b14 <= mdl->smem_optin) { spec_wide = true; threads = 448; }
  }
  if (const char* e = getenv("LMC_SPEC_WIDE")) { if (use_spec && !spec_lists) { spec_wide = atoi(e) != 0; threads = spec_wide ? 448 : 128; } }
  for (;;) {
    a.wpb = threads / G;
    smem = blob + (size_t)a.wpb * (m.Npad + a.walker_smem);
    if ((int)smem <= mdl->smem_optin - 1024 || threads <= 32) break;
    threads -= 32;
  }
  if ((int)smem > mdl->smem_optin - 1024) return fail("model + walker state do not fit in shared memory");
  // Wang-Landau flips with few walkers: the warp-specialised kernel (decision warps + a bookkeeping warp per walker,
  // lmc_wl.cuh) while every walker is resident at seven blocks per SM; LMC_WL2 = 0 classic, 1 one decision warp,
  // 3 depth-2 speculation
  if (c->kernel == LMC_KERNEL_WANGLANDAU && c->usher == LMC_USHER_FLIP && !dist && !ewald && a.off_wl >= 0 && !multicell &&
      c->group_size == 0 && c->block_threads == 0 && !getenv("LMC_GROUP_SIZE") && c->num_walkers <= 7 * mdl->num_sms) {
    int ne = 4;     // 4: merged-record variant where the model has the tables, else one decision warp over classic records
    bool forced = false;   // an explicit LMC_WL2 takes the variant whenever it fits a block (tests), not only at seven blocks per SM
    if (const char* e = getenv("LMC_WL2")) { ne = atoi(e); forced = true; }
    if (ne == 4) {
      const size_t sm3 = m.spOK && m.spFtab && m.kone ? wl3_smem_bytes(m, c->wl.num_bins) : 0;
      if (sm3 && m.spNQ % 8 == 0 && (forced || 7 * (sm3 + 1024) <= 227 * 1024) && (int)sm3 <= mdl->smem_optin - 1024) {
        a.wpb = 1;
        LaunchCfg lc3{a.W, 96, sm3, (cudaStream_t)stream};
        const int rc3 = launch_wl3(m, a, lc3);
        g_launches++;
        if (rc3 != 0) return fail(std::string("launch failed: ") + cudaGetErrorString((cudaError_t)rc3));
        return 0;
      }
      ne = 1;
    }
    if (ne == 1 || ne == 3) {
      const size_t sm2 = wl2_smem_bytes(m, c->wl.num_bins, ne);
      if ((forced || 7 * (sm2 + 1024) <= 227 * 1024) && (int)sm2 <= mdl->smem_optin - 1024) {
        a.wpb = 1;
        LaunchCfg lc2{a.W, 32 * (ne + 1), sm2, (cudaStream_t)stream};
        const int rc2 = launch_wl2(m, a, m.kone != 0, ne, lc2);
        g_launches++;
        if (rc2 != 0) return fail(std::string("launch failed: ") + cudaGetErrorString((cudaError_t)rc2));
        return 0;
      }
    }
  }
  if (tf_spec) {
    // as many walkers per block as fit (<= 16 warps), then evened out over the waves the launch needs anyway
    const size_t blob_all = ((size_t)m.blob_bytes + 15) & ~size_t(15);
    int wpb = 16;
    while (wpb > 1 && blob_all + (size_t)wpb * (m.Npad + a.walker_smem) > (size_t)mdl->smem_optin - 1024) --wpb;
    const long per_wave = (long)wpb * mdl->num_sms;
    const long waves = (a.W + per_wave - 1) / per_wave;
    wpb = (int)std::min<long>(wpb, std::max<long>(1, (a.W + waves * mdl->num_sms - 1) / (waves * mdl->num_sms)));
    a.wpb = wpb;
    LaunchCfg lct{(a.W + wpb - 1) / wpb, 32 * wpb, blob_all + (size_t)wpb * (m.Npad + a.walker_smem), (cudaStream_t)stream};
    const int rct = launch_spec_tf(m, a, m.kone != 0, field, lct);
    g_launches++;
    if (rct != 0) return fail(std::string("launch failed: ") + cudaGetErrorString((cudaError_t)rct));
    if (mm->stats_host) cudaMemcpyAsync(mm->stats_host, mm->stats_dev, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return 0;
  }
  const int grid = (a.W + a.wpb - 1) / a.wpb;
  LaunchCfg lc{grid, threads, smem, (cudaStream_t)stream};
  int rc = -2;
  const bool wl = c->kernel == LMC_KERNEL_WANGLANDAU;
  const int ewmode = !ewald ? 0 : (field ? 2 : 1);   // Ewald path of the classic kernels
  spec_wide = spec_wide && threads == 448;
  if (spec_c64 && threads == 128 && !spec_wide) { rc = launch_spec_c64(m, a, m.kone != 0, c->usher, lc); g_c64_launches++; }
  else if (spec_c64) return fail("internal: the compact-word kernel was planned for blocks of four walkers");
  else if (spec_env) { rc = launch_spec_env(m, a, m.kone != 0, c->usher, field, spec_wide, lc); g_env_launches++; }
  else if (use_spec && (field || spec_wide)) rc = launch_spec_x(m, a, m.kone != 0, c->usher, field, spec_wide, lc);
  else if (use_spec) rc = launch_spec(m, a, m.kone != 0, c->usher, spec_sg, spec_lists, lc);
  else if (dist) rc = launch_run_dist(m, a, m.kone != 0, c->usher, lc);
  else switch (G) {
    case 4: rc = (wl ? launch_run_wl_g4 : launch_run_g4)(m, a, m.kone != 0, ewmode, c->usher, lc); break;
    case 8: rc = (wl ? launch_run_wl_g8 : launch_run_g8)(m, a, m.kone != 0, ewmode, c->usher, lc); break;
    case 16: rc = (wl ? launch_run_wl_g16 : launch_run_g16)(m, a, m.kone != 0, ewmode, c->usher, lc); break;
    case 32: rc = (wl ? launch_run_wl_g32 : launch_run_g32)(m, a, m.kone != 0, ewmode, c->usher, lc); break;
  }
  if (rc == -2) return fail("no kernel instantiated for this group size / usher");
  g_launches++;
  if (rc != 0) return fail(std::string("launch failed: ") + cudaGetErrorString((cudaError_t)rc));
  // acceptance totals follow the launch in stream order; the host looks at them at the next call
  if (mm->stats_host && c->kernel == LMC_KERNEL_METROPOLIS)
    cudaMemcpyAsync(mm->stats_host, mm->stats_dev, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  return 0;
}
