/*
 * CPU ORACLE (C restatement) -- TEST INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Restates smol's per-flip hot path in the reference's own loop order so that it can be
 * compared bit-for-bit with oracle/lmc_oracle.py (compile with -ffp-contract=off, no fast-math):
 *   delta_*_from_occupancies   smol/utils/cluster/evaluator.pyx:211-317
 *   *_from_occupancy           smol/utils/cluster/evaluator.pyx:121-209
 *   delta_ewald_single_flip    smol/utils/cluster/ewald.pyx:9-59
 *   processors                 smol/moca/processor/expansion.py:191-231, 420-464, ewald.py:128-182
 *   ensemble mu term           smol/moca/ensemble.py:323-376
 *   Flip / Swap proposals      smol/moca/kernel/mcusher.py:154-200
 *   Metropolis / Wang-Landau   smol/moca/kernel/metropolis.py:31-49, wanglandau.py:175-266
 *   sampler loop + trace       smol/moca/sampler/sampler.py:195-210
 * RNG: Philox4x32-10 counter stream shared with the CUDA kernels (see oracle/lmc_oracle.py).
 * TableFlip is restated in Python only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t N, F, Fce, n_orb, size; /* Fce = features of the cluster part */
  double feature0;
  const double* nat;
  /* orbits */
  const int32_t* orb_fidx; const int32_t* orb_K; const int32_t* orb_T; const int32_t* orb_I;
  const int32_t* orb_stride; /* [n_orb][8] */
  const int64_t* orb_tab_off; const double* tab;
  const int64_t* full_off; /* [n_orb+1] row offsets */
  const int64_t* full_idx_off; /* [n_orb] offset into full_rows (ints) */
  const int32_t* full_rows;
  /* per-site local evaluators (reference layout: rows containing the site, ratio) */
  const int64_t* site_ptr; /* [N+1] */
  const int32_t* ent_orb; const int64_t* ent_row_off; const int32_t* ent_J; const double* ent_ratio;
  const int32_t* local_rows;
  /* ewald */
  int32_t E, ewW, ewF; const double* ewM; const int32_t* ewInds;
  /* mu */
  int32_t muW, muF; const double* mu;
  /* sublattices (active) */
  int32_t n_sl; const int32_t* sl_off; const int32_t* sl_sites; const int32_t* sl_ncodes;
  const int32_t* sl_codes; /* [n_sl][8] */
  const double* sl_cum;
} OModel;

typedef struct {
  int32_t W, walker_base, usher, kernel, thin;
  int64_t S;
  uint64_t step0;
  const uint64_t* seeds; const double* beta;
  int32_t* occ;        /* [W][N] in/out */
  double* features;    /* [W][F] in/out */
  double* enthalpy;    /* [W] in/out */
  int32_t* tr_occ; double* tr_feat; double* tr_enth; uint8_t* tr_acc; int32_t* tr_nacc; /* may be NULL */
  /* Wang-Landau */
  double wl_min, wl_max, wl_bin, wl_flat, wl_modupd;
  int32_t wl_nb, wl_check, wl_update;
  double* wl_S; int64_t* wl_H; int64_t* wl_O; double* wl_M; double* wl_m; int64_t* wl_cnt;
  int32_t nthreads;
} ORun;

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static inline uint32_t mulhi32(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }
static inline double u01(uint32_t r) { return ((double)r + 0.5) * 2.3283064365386963e-10; }

/* ------------------------------------------------------------------ evaluators */
void o_full_features(const OModel* m, const int32_t* occ, double* out) {
  for (int f = 0; f < m->F; ++f) out[f] = 0.0;
  if (m->n_orb > 0) out[0] = m->feature0; /* evaluator.pyx:143 / :191 */
  for (int n = 0; n < m->n_orb; ++n) {
    const int I = m->orb_I[n], K = m->orb_K[n], T = m->orb_T[n];
    const int64_t J = m->full_off[n + 1] - m->full_off[n];
    const int32_t* rows = m->full_rows + m->full_idx_off[n];
    const int32_t* st = m->orb_stride + n * 8;
    for (int k = 0; k < K; ++k) {
      const double* tk = m->tab + m->orb_tab_off[n] + (int64_t)k * T;
      double p = 0.0;
      for (int64_t j = 0; j < J; ++j) {
        int idx = 0;
        for (int i = 0; i < I; ++i) idx += st[i] * occ[rows[j * I + i]];
        p = p + tk[idx];
      }
      out[m->orb_fidx[n] + k] = p / (double)J * (double)m->size; /* evaluator.pyx:165, expansion.py:184 */
    }
  }
  if (m->E > 0) { /* ewald.py:128-145: sum over occupied x occupied */
    double s = 0.0;
    for (int a = 0; a < m->N; ++a) {
      const int ia = m->ewInds[a * m->ewW + occ[a]];
      if (ia < 0) continue;
      for (int b = 0; b < m->N; ++b) {
        const int ib = m->ewInds[b * m->ewW + occ[b]];
        if (ib >= 0) s += m->ewM[(int64_t)ia * m->E + ib];
      }
    }
    out[m->ewF] = s;
  }
  if (m->muW > 0) { /* ensemble.py:344-349 */
    double s = 0.0;
    for (int a = 0; a < m->N; ++a) s += m->mu[a * m->muW + occ[a]];
    out[m->muF] = s;
  }
}

/* delta of ONE flip, accumulated into dcorr (not yet scaled by size); occ is the pre-flip occupancy */
static void delta_one(const OModel* m, const int32_t* occ, int site, int newc, double* dcorr) {
  const int oldc = occ[site];
  for (int64_t e = m->site_ptr[site]; e < m->site_ptr[site + 1]; ++e) {
    const int n = m->ent_orb[e];
    const int I = m->orb_I[n], K = m->orb_K[n], T = m->orb_T[n], J = m->ent_J[e];
    const int32_t* rows = m->local_rows + m->ent_row_off[e];
    const int32_t* st = m->orb_stride + n * 8;
    for (int k = 0; k < K; ++k) {
      const double* tk = m->tab + m->orb_tab_off[n] + (int64_t)k * T;
      double p = 0.0;
      for (int j = 0; j < J; ++j) {
        int ii = 0, ff = 0;
        for (int i = 0; i < I; ++i) {
          const int s = rows[j * I + i];
          const int o = occ[s];
          ii += st[i] * o;
          ff += st[i] * (s == site ? newc : o);
        }
        p = p + (tk[ff] - tk[ii]);
      }
      dcorr[m->orb_fidx[n] + k] += p / m->ent_ratio[e] / (double)J; /* evaluator.pyx:262 */
    }
  }
  (void)oldc;
}

static double delta_ewald(const OModel* m, const int32_t* occ, int site, int newc) {
  /* ewald.pyx:38-58 */
  const int add = m->ewInds[site * m->ewW + newc], sub = m->ewInds[site * m->ewW + occ[site]];
  double out = 0.0;
  for (int k = 0; k < m->N; ++k) {
    const int i = m->ewInds[k * m->ewW + (k == site ? newc : occ[k])];
    const int j = m->ewInds[k * m->ewW + occ[k]];
    double out_k = 0.0;
    if (i != -1 && add != -1) {
      if (i != add) out_k = out_k + 2 * m->ewM[(int64_t)i * m->E + add];
      else out_k = out_k + m->ewM[(int64_t)i * m->E + add];
    }
    if (j != -1 && sub != -1) {
      if (j != sub) out_k = out_k - 2 * m->ewM[(int64_t)j * m->E + sub];
      else out_k = out_k - m->ewM[(int64_t)j * m->E + sub];
    }
    out += out_k;
  }
  return out;
}

/* feature change of a step (flips applied sequentially); occ restored on return */
void o_delta_features(const OModel* m, int32_t* occ, const int32_t* sites, const int32_t* codes, int nflips, double* out) {
  int32_t saved[16];
  for (int f = 0; f < m->F; ++f) out[f] = 0.0;
  double dmu = 0.0, dew = 0.0;
  for (int f = 0; f < nflips; ++f) { /* mu against the PRE-step occupancy, ensemble.py:369-373 */
    if (m->muW > 0) dmu += m->mu[sites[f] * m->muW + codes[f]] - m->mu[sites[f] * m->muW + occ[sites[f]]];
  }
  for (int f = 0; f < nflips; ++f) {
    if (m->n_orb > 0) delta_one(m, occ, sites[f], codes[f], out);
    if (m->E > 0) dew += delta_ewald(m, occ, sites[f], codes[f]);
    saved[f] = occ[sites[f]];
    occ[sites[f]] = codes[f];
  }
  for (int f = nflips - 1; f >= 0; --f) occ[sites[f]] = saved[f];
  for (int f = 0; f < m->Fce; ++f) out[f] = out[f] * (double)m->size; /* expansion.py:231 */
  if (m->E > 0) out[m->ewF] = dew;
  if (m->muW > 0) out[m->muF] = dmu;
}

static double dot_seq(const double* a, const double* b, int n) {
  double p = 0.0;
  for (int i = 0; i < n; ++i) p = p + a[i] * b[i];
  return p;
}

static int choose_sl(const OModel* m, uint32_t r0) {
  if (m->n_sl == 1) return 0;
  const double u = u01(r0);
  int s = 0;
  while (s < m->n_sl - 1 && !(m->sl_cum[s] > u)) ++s;
  return s;
}

static double py_floordiv(double a, double b) { /* CPython float // */
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

static void run_walker(const OModel* m, const ORun* r, int w, double* dfeat) {
  const int N = m->N, F = m->F;
  int32_t* occ = r->occ + (int64_t)w * N;
  double* feat = r->features + (int64_t)w * F;
  double enth = r->enthalpy[w];
  const uint64_t seed = r->seeds[w];
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), wid = (uint32_t)(r->walker_base + w);
  const int wl = r->kernel == 1;
  const double beta = wl ? 0.0 : r->beta[w];
  double* S = 0; int64_t* H = 0; int64_t* Oc = 0; double* Mf = 0; double wl_m = 0.0; int64_t wl_cnt = 0;
  const int nb = r->wl_nb;
  if (wl) {
    S = r->wl_S + (int64_t)w * nb; H = r->wl_H + (int64_t)w * nb; Oc = r->wl_O + (int64_t)w * nb;
    Mf = r->wl_M + (int64_t)w * nb * F; wl_m = r->wl_m[w]; wl_cnt = r->wl_cnt[w];
  }
  uint64_t step = r->step0;
  for (int64_t s = 0; s < r->S; ++s) {
    int nacc = 0, accepted = 1;
    for (int it = 0; it < r->thin; ++it, ++step) {
      uint32_t q[4];
      philox((uint32_t)step, (uint32_t)(step >> 32), 0u, wid, k0, k1, q);
      int32_t sites[2], codes[2];
      int nfl = 0;
      const int sl = choose_sl(m, q[0]);
      const int off = m->sl_off[sl], nact = m->sl_off[sl + 1] - off;
      const int site1 = m->sl_sites[off + mulhi32(q[1], (uint32_t)nact)];
      if (r->usher == 0) { /* Flip, mcusher.py:154-170 */
        const int cur = occ[site1], nc = m->sl_ncodes[sl];
        int ci = (int)mulhi32(q[2], (uint32_t)(nc - 1)), pos = nc;
        for (int c = 0; c < nc; ++c) if (m->sl_codes[sl * 8 + c] == cur) { pos = c; break; }
        if (ci >= pos) ++ci;
        sites[0] = site1; codes[0] = m->sl_codes[sl * 8 + ci]; nfl = 1;
      } else { /* Swap, mcusher.py:176-200 */
        const int s1 = occ[site1];
        int ndiff = 0;
        for (int j = 0; j < nact; ++j) ndiff += occ[m->sl_sites[off + j]] != s1;
        if (ndiff > 0) {
          int k = (int)mulhi32(q[2], (uint32_t)ndiff), site2 = -1;
          for (int j = 0; j < nact; ++j) {
            const int sj = m->sl_sites[off + j];
            if (occ[sj] != s1) { if (k == 0) { site2 = sj; break; } --k; }
          }
          sites[0] = site1; codes[0] = occ[site2]; sites[1] = site2; codes[1] = s1; nfl = 2;
        }
      }
      o_delta_features(m, occ, sites, codes, nfl, dfeat);
      const double dH = dot_seq(m->nat, dfeat, F);
      if (!wl) {
        const double exponent = -beta * dH + 0.0;
        accepted = exponent >= 0.0 ? 1 : (exponent > log(u01(q[3])));
      } else {
        const double e_new = enth + dH;
        if (e_new < r->wl_min || e_new >= r->wl_max) accepted = 0;
        else {
          const int bin = (int)py_floordiv(enth - r->wl_min, r->wl_bin);
          const int nbin = (int)py_floordiv(e_new - r->wl_min, r->wl_bin);
          const double so = (bin >= 0 && bin < nb) ? S[bin] : 0.0, sn = (nbin >= 0 && nbin < nb) ? S[nbin] : 0.0;
          const double exponent = (so - sn) + 0.0;
          accepted = exponent >= 0.0 ? 1 : (exponent > log(u01(q[3])));
        }
      }
      if (accepted) {
        for (int f = 0; f < nfl; ++f) occ[sites[f]] = codes[f];
        for (int f = 0; f < F; ++f) feat[f] += dfeat[f]; /* sampler.py:204-207 */
        enth += dH;
        ++nacc;
      }
      if (wl) { /* wanglandau.py:222-266 */
        const double fb = py_floordiv(enth - r->wl_min, r->wl_bin);
        if (fb >= 0.0 && fb < (double)nb) {
          const int bin = (int)fb;
          ++wl_cnt;
          const int64_t total = Oc[bin];
          for (int f = 0; f < F; ++f)
            Mf[(int64_t)bin * F + f] = 1.0 / (double)(total + 1) * (feat[f] + (double)total * Mf[(int64_t)bin * F + f]);
          if (wl_cnt % r->wl_update == 0) { S[bin] += wl_m; H[bin] += 1; Oc[bin] += 1; }
        }
        if (wl_cnt % r->wl_check == 0) {
          int nvis = 0; double hsum = 0.0, hmin = 1e300;
          for (int b = 0; b < nb; ++b) if (S[b] > 0.0) { ++nvis; hsum += (double)H[b]; if ((double)H[b] < hmin) hmin = (double)H[b]; }
          if (nvis >= 2 && hmin > r->wl_flat * (hsum / (double)nvis)) {
            for (int b = 0; b < nb; ++b) H[b] = 0;
            wl_m = wl_m / r->wl_modupd;
          }
        }
      }
    }
    const int64_t sw = s * r->W + w;
    if (r->tr_occ) memcpy(r->tr_occ + sw * N, occ, sizeof(int32_t) * N);
    if (r->tr_feat) memcpy(r->tr_feat + sw * F, feat, sizeof(double) * F);
    if (r->tr_enth) r->tr_enth[sw] = enth;
    if (r->tr_acc) r->tr_acc[sw] = (uint8_t)accepted;
    if (r->tr_nacc) r->tr_nacc[sw] = nacc;
  }
  r->enthalpy[w] = enth;
  if (wl) { r->wl_m[w] = wl_m; r->wl_cnt[w] = wl_cnt; }
}

int o_run(const OModel* m, const ORun* r) {
  int nt = r->nthreads > 0 ? r->nthreads : 1;
#pragma omp parallel num_threads(nt)
  {
    double* dfeat = (double*)malloc(sizeof(double) * (m->F + 1));
#pragma omp for schedule(dynamic, 1)
    for (int w = 0; w < r->W; ++w) run_walker(m, r, w, dfeat);
    free(dfeat);
  }
  return 0;
}

void o_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  philox(c0, c1, c2, c3, k0, k1, out);
}
