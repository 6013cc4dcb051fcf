"""Kernel throughput of BASELINE.json configs 3, 4, 5 (config 2 is bench.py). Prints one JSON per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import smol_b200 as S
from smol_b200 import lattice as L
from tests import models as M

try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0      # B200_PROFILING.md fallback
# (superseded by `bench.py --config {3,4,5}`; kept for quick kernel-only A/B runs)


def timed(smp, nsteps, occ0, thin, reps=3):
    smp.run(thin * 2, occ0, thin_by=thin)
    smp.clear_samples()
    best = None
    for _ in range(reps):
        smp.run(nsteps, thin_by=thin)
        ms = smp.last_kernel_ms
        best = ms if best is None else min(best, ms)
        acc = smp.samples.step_efficiency()
        smp.clear_samples()
    return best, acc


def config3(W=4096):
    sub = M.rocksalt_subspace(); scm = np.eye(3, dtype=int) * 8
    rng = np.random.default_rng(3)
    coefs = rng.normal(0, 0.02, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    ewm, ewi = L.ewald_matrix(sub, scm)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.1, ewald_matrix=ewm, ewald_inds=ewi))
    ens = S.Ensemble(comp, chemical_potentials={"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2})
    occ0 = M.random_occupancies(sub, scm, W, seed=5)
    smp = S.Sampler.from_ensemble(ens, 1500.0, step_type="flip", nwalkers=W, seeds=list(range(W)), record_occupancy=True)
    nsteps = 512 * 4
    ms, acc = timed(smp, nsteps, occ0, 512)
    rate = W * nsteps / ms * 1e3
    bytes_step = 214 + 2 * 1024 * 8 + 1024
    return dict(config="3: ternary rocksalt 8x8x8 + Ewald, semigrand flip, %d walkers" % W, steps_per_s=rate, ms=ms,
                acceptance=acc, algorithmic_bytes_per_step=bytes_step, achieved_GBps=rate * bytes_step / 1e9,
                frac_of_hbm_peak=rate * bytes_step / 1e9 / PEAK)


def config4(W=1024):
    sub = M.fcc_subspace(); scm = np.eye(3, dtype=int) * 8
    # AFM Ising-like: point h = 2, NN pair J = 1 (x multiplicity), all else 0 (wang-landau-ising notebook), kB := 1
    coefs = np.zeros(sub.num_corr_functions)
    mult = sub.function_total_multiplicities
    coefs[1] = 2.0 * mult[1]; coefs[2] = 1.0 * mult[2]
    it = L.cluster_interaction_tensors(sub, coefs)
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it))
    occ0 = M.random_occupancies(sub, scm, W, seed=2)
    e0 = ens.compute_feature_vector_batch(occ0[:64]) @ ens.natural_parameters
    lo, hi = float(e0.mean() - 5 * e0.std() - 200), float(e0.mean() + 5 * e0.std() + 200)
    bin_size = 4.0
    lo = np.floor(lo / bin_size) * bin_size - 2.0     # levels centred on the lattice energies (multiples of 4)
    smp = S.Sampler.from_ensemble(ens, lo, hi, bin_size, step_type="flip", kernel_type="WangLandau", nwalkers=W,
                                  seeds=list(range(W)), flatness=0.8, check_period=1000, record_occupancy=False)
    nsteps = 512 * 40
    ms, acc = timed(smp, nsteps, occ0, 512, reps=2)
    rate = W * nsteps / ms * 1e3
    st = smp.wang_landau_state
    bytes_step = 214 + 16 + 48 + 8 * 16
    return dict(config="4: binary FCC 8x8x8 Wang-Landau flip, %d independent walkers" % W, steps_per_s=rate, ms=ms,
                acceptance=acc, bins=int(len(st["levels"])), min_mod_factor=float(st["mod_factor"].min()),
                max_mod_factor=float(st["mod_factor"].max()), visited_bins_mean=float((st["entropy"] > 0).sum(1).mean()),
                algorithmic_bytes_per_step=bytes_step, achieved_GBps=rate * bytes_step / 1e9,
                frac_of_hbm_peak=rate * bytes_step / 1e9 / PEAK)


def config5(W=4096, n=12):
    sub = M.rocksalt_subspace(anions=("O2-", "F-")); scm = np.eye(3, dtype=int) * n
    rng = np.random.default_rng(21)
    coefs = rng.normal(0, 0.01, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    t0 = time.time()
    ewm, ewi = L.ewald_matrix(sub, scm)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.1, ewald_matrix=ewm, ewald_inds=ewi))
    mus = {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0}
    ens = S.Ensemble(comp, chemical_potentials=mus)
    nc = n ** 3
    # charge neutral start: nLi + 3 nMn + 4 nTi = 2 nO + nF
    nMn, nTi = nc // 6, nc // 8
    nLi = nc - nMn - nTi
    q = nLi + 3 * nMn + 4 * nTi
    nO = q - nc
    cat = np.array([0] * nLi + [1] * nMn + [2] * nTi); ani = np.array([0] * nO + [1] * (nc - nO))
    occ0 = np.zeros((W, 2 * nc), dtype=np.int32)
    for w in range(W):
        occ0[w, :nc] = rng.permutation(cat); occ0[w, nc:] = rng.permutation(ani)
    table = [[-1, 1, 0, 2, -2], [0, -1, 1, 1, -1]]
    smp = S.Sampler.from_ensemble(ens, 1500.0, step_type="table_flip", nwalkers=W, seeds=list(range(W)),
                                  flip_table=table, swap_weight=0.1, record_occupancy=False)
    N = 2 * nc
    nsteps = 256
    ms, acc = timed(smp, nsteps, occ0, 128, reps=2)
    rate = W * nsteps / ms * 1e3
    bytes_step = 2.5 * (250 + 2 * N * 8 + N)
    return dict(config="5: 5-species rocksalt %dx%dx%d + Ewald, charge-neutral table flips, %d walkers (1 GPU)" % (n, n, n, W),
                steps_per_s=rate, ms=ms, acceptance=acc, setup_s=time.time() - t0, algorithmic_bytes_per_step=bytes_step,
                achieved_GBps=rate * bytes_step / 1e9, frac_of_hbm_peak=rate * bytes_step / 1e9 / PEAK)


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4", "5"]
    for c in which:
        out = {"3": config3, "4": config4, "5": config5}[c]()
        print(json.dumps(out), flush=True)
