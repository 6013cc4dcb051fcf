"""Parity at BASELINE.json's full sizes: bit-exact trajectories against the C oracle on walker
subsets, plus size-independent properties over all walkers."""
import warnings

import numpy as np
import pytest

from tests import models as M
from tests import workloads as WK

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _c_oracle(ens_g, ora_p, mus=None):
    from oracle import c_oracle as CO
    from oracle import lmc_oracle as O
    return CO.COracle(O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus))


def test_config2_full_size_canonical_swap(cuda_device):
    """binary FCC 8x8x8, 4096 walkers: 2 sweeps; 48 walkers checked bit-exact against the C oracle,
    ALL walkers: composition conserved, recorded features == full re-evaluation of the recorded
    occupancy (tests/test_moca/test_sampler.py:59-85), resume == one long run."""
    import smol_b200 as S
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 8
    coefs = M.fcc_coefs(sub)
    it = L.cluster_interaction_tensors(sub, coefs)
    proc = S.ClusterDecompositionProcessor(sub, scm, it)
    ens = S.Ensemble(proc)
    W, N = 4096, 512
    occ0 = M.random_occupancies(sub, scm, W, seed=0, balanced=True)
    seeds = np.arange(W) * 7919 + 13
    smp = S.Sampler.from_ensemble(ens, 1000.0, step_type="swap", nwalkers=W, seeds=list(seeds))
    smp.run(2 * N, occ0, thin_by=N // 2)
    s = smp.samples
    occ_tr = s.get_occupancies(flat=False)
    assert occ_tr.shape == (4, W, N) and occ_tr.dtype == np.int32
    assert np.all(occ_tr.sum(axis=2) == 256)                                   # canonical: composition fixed
    feats = s.get_feature_vectors(flat=False)
    full_last = proc.compute_feature_vector_batch(occ_tr[-1])
    np.testing.assert_allclose(feats[-1], full_last, rtol=0, atol=5e-13 * 512)
    np.testing.assert_allclose(s.get_enthalpies(flat=False)[-1, :, 0], full_last @ ens.natural_parameters,
                               rtol=RTOL, atol=1e-9)
    pick = np.random.default_rng(1).choice(W, size=48, replace=False)
    co = _c_oracle(ens, O.ClusterDecompositionProcessor(sub, scm, it))
    for w in pick[:48:4]:      # walker ids matter: run them one by one with their global id
        out, _ = co.run(occ0[w:w + 1], 2 * N, N // 2, seeds[w:w + 1], usher="swap", temperature=1000.0,
                        walker_base=int(w))
        np.testing.assert_array_equal(out["occupancy"][:, 0], occ_tr[:, w])
        np.testing.assert_array_equal(out["n_accepted"][:, 0], s.get_trace_value("n_accepted", flat=False)[:, w])
        np.testing.assert_allclose(out["features"][:, 0], feats[:, w], rtol=RTOL, atol=RTOL * np.abs(feats).max())
    # resume (run(None)) continues the same Markov chains: counter-based RNG
    smp2 = S.Sampler.from_ensemble(ens, 1000.0, step_type="swap", nwalkers=W, seeds=list(seeds))
    smp2.run(N, occ0, thin_by=N // 2)
    smp2.run(N, thin_by=N // 2)
    np.testing.assert_array_equal(smp2.samples.get_occupancies(flat=False), occ_tr)


def test_config3_semigrand_ewald_full_size(cuda_device):
    """ternary rocksalt 8x8x8 + Ewald, semigrand flips, 512 walkers, 1 sweep of the cation sublattice."""
    import smol_b200 as S
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 8
    rng = np.random.default_rng(3)
    coefs = rng.normal(0, 0.02, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    ewm, ewi = L.ewald_matrix(sub, scm)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.1, ewald_matrix=ewm, ewald_inds=ewi))
    mus = {"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2}
    ens = S.Ensemble(comp, chemical_potentials=mus)
    W, N = 512, 1024
    occ0 = M.random_occupancies(sub, scm, W, seed=5)
    seeds = np.arange(W) + 1000
    smp = S.Sampler.from_ensemble(ens, 1500.0, step_type="flip", nwalkers=W, seeds=list(seeds))
    smp.run(512, occ0, thin_by=256)
    s = smp.samples
    occ_tr = s.get_occupancies(flat=False)
    assert np.all(occ_tr[:, :, 512:] == 0)                                      # O2- sublattice never touched
    feats = s.get_feature_vectors(flat=False)
    full_last = ens.compute_feature_vector_batch(occ_tr[-1])
    scale = np.abs(full_last).max()
    np.testing.assert_allclose(feats[-1], full_last, rtol=RTOL, atol=RTOL * scale)
    ora_p = O.CompositeProcessor([O.ClusterDecompositionProcessor(sub, scm, it), O.EwaldProcessor(ewm, ewi, 0.1)])
    co = _c_oracle(ens, ora_p, mus)
    for w in (0, 17, 511):
        out, _ = co.run(occ0[w:w + 1], 512, 256, seeds[w:w + 1], usher="flip", temperature=1500.0, walker_base=w)
        np.testing.assert_array_equal(out["occupancy"][:, 0], occ_tr[:, w])
        np.testing.assert_allclose(out["features"][:, 0], feats[:, w], rtol=RTOL, atol=RTOL * scale)


@pytest.mark.parametrize("kernel", ["classic", "warp-specialised"])
def test_config4_wang_landau_full_size(cuda_device, kernel, monkeypatch):
    """binary FCC 8x8x8 Wang-Landau (AFM Ising coefficients of the wang-landau notebook), 1024 independent
    walkers: walkers checked bit-exact against the C oracle incl. their entropy / histogram; ALL walkers:
    occurrences count every step inside the window, histogram <= occurrences, enthalpy == full re-evaluation,
    visited levels inside the window."""
    import smol_b200 as S
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    if kernel == "classic":
        monkeypatch.setenv("LMC_WL2", "0")
    else:
        monkeypatch.delenv("LMC_WL2", raising=False)      # default: lmc_wl.cuh (1024 walkers: all resident)
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 8
    coefs = np.zeros(sub.num_corr_functions)
    mult = sub.function_total_multiplicities
    coefs[1], coefs[2] = 2.0 * mult[1], 1.0 * mult[2]
    it = L.cluster_interaction_tensors(sub, coefs)
    proc = S.ClusterDecompositionProcessor(sub, scm, it)
    ens = S.Ensemble(proc)
    W, N = 1024, 512
    occ0 = M.random_occupancies(sub, scm, W, seed=2)
    e0 = ens.compute_feature_vector_batch(occ0[:64]) @ ens.natural_parameters
    lo = float(np.floor((e0.mean() - 5 * e0.std() - 200) / 4.0) * 4.0 - 2.0)
    hi = float(e0.mean() + 5 * e0.std() + 200)
    seeds = np.arange(W) * 31 + 5
    nsteps, thin = 4096, 1024
    smp = S.Sampler.from_ensemble(ens, lo, hi, 4.0, step_type="flip", kernel_type="WangLandau", nwalkers=W,
                                  seeds=list(seeds), flatness=0.8, check_period=1000)
    smp.run(nsteps, occ0, thin_by=thin)
    st = smp.wang_landau_state
    enth = smp.samples.get_enthalpies(flat=False)[:, :, 0]
    assert (enth >= lo).all() and (enth < hi).all()
    assert (st["occurrences"].sum(axis=1) == nsteps).all()                   # every step ends inside the window
    assert (st["histogram"] <= st["occurrences"]).all() and (st["entropy"] >= 0).all()
    assert ((st["entropy"] > 0) == (st["occurrences"] > 0)).all()
    occ_last = smp.samples.get_occupancies(flat=False)[-1]
    full = ens.compute_feature_vector_batch(occ_last) @ ens.natural_parameters
    np.testing.assert_allclose(enth[-1], full, rtol=RTOL, atol=1e-9)
    co = _c_oracle(ens, O.ClusterDecompositionProcessor(sub, scm, it))
    for w in (0, 333, 1023):
        out, state = co.run(occ0[w:w + 1], nsteps, thin, seeds[w:w + 1], usher="flip", walker_base=w,
                            wl=dict(min=lo, max=hi, bin=4.0, flatness=0.8, check=1000))
        np.testing.assert_array_equal(out["occupancy"][:, 0], smp.samples.get_occupancies(flat=False)[:, w])
        np.testing.assert_array_equal(state["histogram"][0], st["histogram"][w])
        np.testing.assert_array_equal(state["occurrences"][0], st["occurrences"][w])
        np.testing.assert_allclose(state["entropy"][0], st["entropy"][w], rtol=1e-12, atol=0)
        assert state["mod_factor"][0] == st["mod_factor"][w]


def test_config5_table_flip_ewald_full_cell(cuda_device):
    """5-species rocksalt 12x12x12 (N = 3456, Ewald E = 8640) + charge-neutral table flips: properties that do
    not depend on the size -- charge neutrality and site conservation of every sample, running features ==
    full re-evaluation (Ewald row-gather AND potential-cache paths give the same chains), 64 walkers."""
    import smol_b200 as S
    from smol_b200 import lattice as L
    n = 12
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * n
    rng = np.random.default_rng(21)
    it = L.cluster_interaction_tensors(sub, rng.normal(0, 0.01, sub.num_corr_functions))
    ewm, ewi = L.ewald_matrix(sub, scm)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.1, ewald_matrix=ewm, ewald_inds=ewi))
    mus = {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0}
    ens = S.Ensemble(comp, chemical_potentials=mus)
    nc, W = n ** 3, 64
    nMn, nTi = nc // 6, nc // 8
    nLi = nc - nMn - nTi
    nO = nLi + 3 * nMn + 4 * nTi - nc
    cat = np.array([0] * nLi + [1] * nMn + [2] * nTi)
    ani = np.array([0] * nO + [1] * (nc - nO))
    occ0 = np.zeros((W, 2 * nc), dtype=np.int32)
    for w in range(W):
        occ0[w, :nc], occ0[w, nc:] = rng.permutation(cat), rng.permutation(ani)
    table = [[-1, 1, 0, 2, -2], [0, -1, 1, 1, -1]]
    res = []
    for field in (False, True):
        smp = S.Sampler.from_ensemble(ens, 1500.0, step_type="table_flip", nwalkers=W, seeds=list(range(W)),
                                      flip_table=table, swap_weight=0.1, ewald_field=field)
        smp.run(600, occ0, thin_by=200)
        occ = smp.samples.get_occupancies(flat=False)
        q = np.array([1, 3, 4])[occ[:, :, :nc]].sum(axis=2) + np.array([-2, -1])[occ[:, :, nc:]].sum(axis=2)
        assert (q == 0).all()                                               # charge neutral at every sample
        assert smp.samples.step_efficiency() > 0.1
        feats = smp.samples.get_feature_vectors(flat=False)
        full = ens.compute_feature_vector_batch(occ[-1])
        np.testing.assert_allclose(feats[-1], full, rtol=RTOL, atol=RTOL * np.abs(full).max())
        res.append(occ)
    np.testing.assert_array_equal(res[0], res[1])
    assert (res[0][-1] != occ0).any()


def test_config5_speculative_table_flips_equal_classic_chains_over_many_sweeps(cuda_device):
    """BASELINE config 5 (N = 3456), 256 walkers x 12 sweeps (1.06e7 steps) through the equilibration -- acceptance from
    ~18 % of the first sweep down to a few per cent: the speculative table-flip kernel (csrc/lmc_spec_tf.cuh) and the classic kernel walk
    the SAME chains (occupancies and accepted-step counts identical at every sample, features / enthalpies to rounding),
    and the running features equal a full re-evaluation at the end"""
    wk = WK.get(5)
    ens = wk.product_ensemble()
    W = 256
    occ0 = wk.initial_occupancies(W, seed=5)
    seeds = list(range(9000, 9000 + W))
    runs = []
    for spec_mode in (1, 2):
        smp = wk.sampler(ens, W, seeds, ewald_field=True, spec_mode=spec_mode)
        smp.run(12 * wk.thin_by, occ0, thin_by=wk.thin_by)
        runs.append(smp.samples)
    a, b = runs
    np.testing.assert_array_equal(a.get_occupancies(flat=False), b.get_occupancies(flat=False))
    np.testing.assert_array_equal(a.get_trace_value("n_accepted", flat=False), b.get_trace_value("n_accepted", flat=False))
    nacc = b.get_trace_value("n_accepted", flat=False)
    assert nacc[0].mean() > 0.1 * wk.thin_by and 0 < nacc[-1].mean() < 0.5 * nacc[0].mean()     # hot start, then cold
    fa, fb = a.get_feature_vectors(flat=False), b.get_feature_vectors(flat=False)
    scale = np.abs(fa).max()
    np.testing.assert_allclose(fb, fa, rtol=1e-10, atol=1e-10 * scale)
    np.testing.assert_allclose(b.get_enthalpies(flat=False), a.get_enthalpies(flat=False), rtol=1e-10,
                               atol=1e-10 * np.abs(a.get_enthalpies(flat=False)).max())
    full = ens.compute_feature_vector_batch(b.get_occupancies(flat=False)[-1])
    np.testing.assert_allclose(fb[-1], full, rtol=1e-9, atol=1e-9 * np.abs(full).max())


def test_edge_cases(cuda_device):
    """empty swap (mcusher.py:194-199), restricted sites, sublattice probabilities, single walker,
    thin_by remainder warning (sampler.py:183-188), wrong shapes, anneal."""
    import smol_b200 as S
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 3
    rng = np.random.default_rng(9)
    coefs = rng.normal(0, 0.03, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    proc = S.ClusterDecompositionProcessor(sub, scm, it)
    ens = S.Ensemble(proc)
    # all cations Li, all anions O: no swap possible -> empty steps, recorded as accepted
    occ = np.zeros((1, 54), dtype=np.int32)
    smp = S.Sampler.from_ensemble(ens, 1000.0, step_type="swap", nwalkers=1, seeds=[3])
    with pytest.warns(RuntimeWarning):
        smp.run(25, occ[0], thin_by=10)
    assert smp.samples.num_samples == 2
    assert np.all(smp.samples.get_occupancies() == 0) and smp.samples.sampling_efficiency() == 1.0
    assert np.all(smp.samples.get_trace_value("n_accepted") == 10)
    # restricted sites (non-contiguous active list) + sublattice probabilities vs the oracle
    ens.restrict_sites([0, 3, 4, 30, 31])
    probs = [0.8, 0.2]
    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=2)
    smp = S.Sampler.from_ensemble(ens, 1200.0, step_type="flip", nwalkers=W, seeds=[1, 2, 3, 4],
                                  sublattice_probabilities=probs)
    smp.run(300, occ0, thin_by=10)
    subl = M.oracle_sublattices(O, ens.sublattices)
    ora_p = O.ClusterDecompositionProcessor(sub, scm, it)
    ks = [O.Metropolis(O.Ensemble(ora_p, subl), O.Flip(subl, probs), 1200.0, seed=w + 1, walker=w) for w in range(W)]
    ref = O.run_sampler(ks, occ0, 300, 10)
    np.testing.assert_array_equal(smp.samples.get_occupancies(flat=False), ref["occupancy"])
    tr = smp.samples.get_occupancies(flat=False)
    for site in (0, 3, 4, 30, 31):
        assert np.all(tr[:, :, site] == occ0[None, :, site])
    with pytest.raises(AttributeError):
        smp.run(10, np.zeros((3, 54), dtype=np.int32))
    with pytest.raises(ValueError):
        S.Sampler.from_ensemble(ens, 1000.0, nwalkers=2, seeds=[1])
    # anneal: temperatures recorded, chains continue (sampler.py:303-384)
    ens.reset_restricted_sites()
    smp = S.Sampler.from_ensemble(ens, 2000.0, step_type="swap", nwalkers=2, seeds=[5, 6])
    smp.anneal([2000.0, 1000.0, 500.0], 100, occ0[:2], thin_by=50)
    t = smp.samples.get_temperatures()           # flat like the reference's (container.py:231-233): [samples x walkers]
    assert t.shape == (12,) and list(t.reshape(6, 2)[:, 0]) == [2000.0, 2000.0, 1000.0, 1000.0, 500.0, 500.0]
    with pytest.raises(ValueError):
        smp.anneal([500.0, 1000.0], 10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e = smp.samples.get_enthalpies(flat=False)
    assert e.shape == (6, 2, 1)
