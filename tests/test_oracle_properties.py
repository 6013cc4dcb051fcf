"""Property tests of the oracle in the style of the reference's own suite
(tests/test_moca/test_processor.py:170-231, test_sampler.py:59-85, test_mcushers.py:124-196)."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import lmc_oracle as O
from smol_b200 import lattice as L
from tests import models as M

RTOL = 1e-12
ATOL = 2e4 * np.finfo(float).eps


def _fcc(n, kind="cd", seed=3):
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * n
    coefs = M.fcc_coefs(sub, seed=seed)
    if kind == "cd":
        p = O.ClusterDecompositionProcessor(sub, scm, L.cluster_interaction_tensors(sub, coefs))
    else:
        p = O.ClusterExpansionProcessor(sub, scm, coefs)
    subl = [O.Sublattice(("A", "B"), np.arange(n ** 3))]
    return sub, scm, p, subl


def test_lattice_geometry_counts():
    """S_fcc multiplicities 1,6,3,12,6,8,2; clusters per site 1,12,6,24,12,24,8."""
    sub = M.fcc_subspace()
    assert [o.multiplicity for o in sub.orbits] == [1, 6, 3, 12, 6, 8, 2]
    assert sub.num_orbits == 8 and sub.num_corr_functions == 8
    idx = sub.get_orbit_indices(np.eye(3, dtype=int) * 8).arrays
    assert [int((a == 0).any(axis=1).sum()) for a in idx] == [1, 12, 6, 24, 12, 24, 8]
    assert len(L.fcc_prim().space_group()) == 48


def test_madelung_constant():
    nacl = L.rocksalt_prim(a=5.64, cations=("Na+",), anions=("Cl-",), charges={"Na+": 1, "Cl-": -1})
    m, _ = L.ewald_matrix(L.ClusterSubspace(nacl, []), np.eye(3, dtype=int) * 2)
    assert -m.sum() / 8 * 2.82 / 14.39964547842567 == pytest.approx(1.747564594633, rel=1e-9)
    assert np.array_equal(m, m.T)


@pytest.mark.parametrize("kind", ["cd", "ce"])
@pytest.mark.parametrize("n", [2, 3])
def test_delta_equals_full_difference(kind, n):
    """test_processor.py:175-231: delta == full(new) - full(old), reverse == -forward."""
    sub, scm, p, _ = _fcc(n, kind)
    rng = np.random.default_rng(0)
    occ = M.random_occupancies(sub, scm, 1, seed=2)[0]
    for _ in range(30):
        s = int(rng.integers(len(occ)))
        new = 1 - occ[s]
        d = p.compute_feature_vector_change(occ, [(s, new)])
        occ2 = occ.copy()
        occ2[s] = new
        np.testing.assert_allclose(d, p.compute_feature_vector(occ2) - p.compute_feature_vector(occ), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(p.compute_feature_vector_change(occ2, [(s, occ[s])]), -d, rtol=RTOL, atol=ATOL)
        occ = occ2


def test_ce_and_cd_give_the_same_energy():
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(4)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    ce = O.ClusterExpansionProcessor(sub, scm, coefs)
    cd = O.ClusterDecompositionProcessor(sub, scm, L.cluster_interaction_tensors(sub, coefs))
    for occ in M.random_occupancies(sub, scm, 5, seed=6):
        assert ce.compute_property(occ) == pytest.approx(cd.compute_property(occ), rel=1e-12)


def test_ewald_delta_equals_full_difference():
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 2
    ewm, ewi = L.ewald_matrix(sub, scm)
    ew = O.EwaldProcessor(ewm, ewi)
    rng = np.random.default_rng(1)
    occ = M.random_occupancies(sub, scm, 1, seed=3)[0]
    for _ in range(20):
        s = int(rng.integers(8))
        new = int((occ[s] + 1 + rng.integers(2)) % 3)
        d = ew.compute_feature_vector_change(occ, [(s, new)])
        occ2 = occ.copy()
        occ2[s] = new
        assert d == pytest.approx(ew.compute_feature_vector(occ2) - ew.compute_feature_vector(occ), rel=1e-10, abs=1e-9)
        occ = occ2


def test_sampled_traces_recomputable_from_occupancies():
    """test_sampler.py:59-85: recorded features == compute_feature_vector(recorded occupancy)."""
    sub, scm, p, subl = _fcc(3)
    W = 3
    occ0 = M.random_occupancies(sub, scm, W, seed=1, balanced=False)
    ks = [O.Metropolis(O.Ensemble(p, subl), O.Swap(subl), 2000.0, seed=w, walker=w) for w in range(W)]
    tr = O.run_sampler(ks, occ0, 600, 10)
    for s in (0, 17, 59):
        for w in range(W):
            np.testing.assert_allclose(tr["features"][s, w], p.compute_feature_vector(tr["occupancy"][s, w]), atol=5e-13 * 27)
    assert 0 <= tr["n_accepted"].sum() / 1800 <= 1


def test_c_oracle_trajectories_bitwise_vs_python():
    sub, scm, p, subl = _fcc(3, "ce")
    W = 3
    occ0 = M.random_occupancies(sub, scm, W, seed=1)
    seeds = [11, 12, 13]
    for usher, U in (("swap", O.Swap), ("flip", O.Flip)):
        ks = [O.Metropolis(O.Ensemble(p, subl), U(subl), 1500.0, seed=seeds[w], walker=w) for w in range(W)]
        ref = O.run_sampler(ks, occ0, 300, 10)
        out, _ = CO.COracle(O.Ensemble(p, subl)).run(occ0, 300, 10, seeds, usher=usher, temperature=1500.0)
        for k in ("occupancy", "features", "enthalpy", "accepted", "n_accepted"):
            np.testing.assert_array_equal(out[k], ref[k])


def test_c_oracle_wang_landau_bitwise_vs_python():
    sub, scm, p, subl = _fcc(3)
    W = 2
    occ0 = M.random_occupancies(sub, scm, W, seed=5)
    e0 = [p.compute_property(o) for o in occ0]
    wl = dict(min=min(e0) - 2.0, max=max(e0) + 2.0, bin=0.11, check=40, flatness=0.3)
    ks = [O.WangLandau(O.Ensemble(p, subl), O.Flip(subl), wl["min"], wl["max"], wl["bin"], flatness=0.3,
                       check_period=40, seed=w + 1, walker=w) for w in range(W)]
    ref = O.run_sampler(ks, occ0, 800, 40)
    out, st = CO.COracle(O.Ensemble(p, subl)).run(occ0, 800, 40, [1, 2], usher="flip", wl=wl)
    np.testing.assert_array_equal(out["occupancy"], ref["occupancy"])
    for w, k in enumerate(ks):
        np.testing.assert_array_equal(st["histogram"][w], k._histogram)
        np.testing.assert_array_equal(st["entropy"][w], k._entropy)
        np.testing.assert_array_equal(st["mean_features"][w], k._mean_features)
        assert st["mod_factor"][w] == k._m


def test_proposal_statistics():
    """test_mcushers.py:124-196: valid proposals, uniform site visits, new != old species."""
    sub, scm, p, subl = _fcc(2)
    occ = np.array([0, 1, 0, 1, 0, 1, 0, 1])
    flip, swap = O.Flip(subl), O.Swap(subl)
    visits = np.zeros(8)
    for step in range(4000):
        rnd = O.StepRandom(5, 0, step)
        (s, c), = flip.propose_step(occ, rnd)
        assert c != occ[s]
        visits[s] += 1
        sw = swap.propose_step(occ, rnd)
        assert len(sw) == 2 and occ[sw[0][0]] != occ[sw[1][0]] and sw[0][1] == occ[sw[1][0]]
    assert np.all(np.abs(visits / 4000 - 1 / 8) < 0.02)
    assert swap.propose_step(np.zeros(8, dtype=int), O.StepRandom(1, 0, 0)) == []  # mcusher.py:194-199


def test_table_flip_samples_neutral_states_uniformly():
    """test_mcushers.py:237-319 (shortened): detailed balance of the a-priori factor."""
    sl = [O.Sublattice(("Li+", "Zr4+", "Mn3+"), [0, 1, 2]), O.Sublattice(("O2-", "F-"), [3, 4, 5])]
    tf = O.TableFlip(sl, [[0, -1, 1, -1, 1], [1, 0, -1, -2, 2]])
    occ = np.array([0, 0, 1, 0, 0, 0])
    from collections import Counter
    cnt = Counter()
    nsteps = 30000
    for step in range(nsteps):
        rnd = O.StepRandom(99, 0, step)
        st = tf.propose_step(occ, rnd)
        lp = tf.compute_log_priori_factor(occ, st)
        if lp >= 0 or lp > np.log(O.u01(rnd.word(3))):
            for s, c in st:
                occ[s] = c
        q = np.array([1, 4, 3])[occ[:3]].sum() + np.array([-2, -1])[occ[3:]].sum()
        assert q == 0
        cnt[tuple(occ)] += 1
    freq = np.array(list(cnt.values())) / nsteps * len(cnt)
    # compositions (Li,Zr,Mn|O,F): (2,1,0|3,0) 3 states, (2,0,1|2,1) 9 states, (3,0,0|0,3) 1 state
    assert len(cnt) == 13
    np.testing.assert_allclose(freq, 1.0, atol=0.15)


def test_distance_processors_match_the_compiled_reference_evaluators():
    """oracle restatement of evaluator.pyx:319-435 + processor/distance.py against the reference's own compiled
    corr_distances / interaction_distances_from_occupancies (oracle/_ref); change == difference of vectors"""
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(1)
    occs = M.random_occupancies(sub, scm, 4, seed=3)
    tv = rng.normal(0, 0.2, sub.num_corr_functions)
    tv[0] = 1.0
    it = L.cluster_interaction_tensors(sub, rng.normal(0, 0.05, sub.num_corr_functions))
    tvi = rng.normal(0, 0.05, sub.num_orbits)
    have_ref = O.load_ref() is not None
    for make in (lambda r: O.CorrelationDistanceProcessor(sub, scm, target_vector=tv, match_weight=1.0, use_ref=r),
                 lambda r: O.ClusterInteractionDistanceProcessor(sub, scm, it, target_vector=tvi, use_ref=r)):
        port = make(False)
        ref = make(True) if have_ref else None
        for occ in occs:
            flips = [(0, int((occ[0] + 1) % 3)), (3, int((occ[3] + 2) % 3))]
            fv, dv = port.compute_feature_vector(occ), port.compute_feature_vector_change(occ, flips)
            nxt = occ.copy()
            for s_, c_ in flips:
                nxt[s_] = c_
            np.testing.assert_allclose(dv, port.compute_feature_vector(nxt) - fv, rtol=0, atol=1e-13)
            if ref is not None:
                np.testing.assert_allclose(fv, ref.compute_feature_vector(occ), rtol=1e-12, atol=1e-14)
                np.testing.assert_allclose(dv, ref.compute_feature_vector_change(occ, flips), rtol=0, atol=1e-13)
    # exact matches: the target taken from an occupancy gives L = the largest diameter
    base = O.ClusterExpansionProcessor(sub, scm, np.zeros(sub.num_corr_functions))
    p = O.CorrelationDistanceProcessor(sub, scm, target_vector=base.compute_feature_vector(occs[0]) / base.size)
    assert p.compute_feature_vector(occs[0])[0] == max(O.orbits_by_diameter(sub))
    assert p.compute_feature_vector(occs[1])[0] < max(O.orbits_by_diameter(sub))


def test_ewald_term_matrices_add_up():
    """EwaldTerm.use_term (cofe/extern/ewald.py:168-177): total = reciprocal + real + point; the point matrix is the
    diagonal self term; each part keeps the q_i q_j structure the engine factorises"""
    from types import SimpleNamespace
    import smol_b200 as S
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 2
    parts = {t: L.ewald_matrix(sub, scm, backend="numpy", term=t)[0] for t in ("total", "reciprocal", "real", "point")}
    np.testing.assert_allclose(parts["reciprocal"] + parts["real"] + parts["point"], parts["total"], rtol=0,
                               atol=1e-12 * np.abs(parts["total"]).max())
    assert np.count_nonzero(parts["point"] - np.diag(np.diag(parts["point"]))) == 0 and (np.diag(parts["point"]) < 0).all()
    assert np.abs(parts["real"]).max() > 0 and np.abs(parts["reciprocal"]).max() > 0
    with pytest.raises(ValueError, match="term"):
        L.ewald_matrix(sub, scm, backend="numpy", term="madelung")
    # the processor reads the term off an EwaldTerm-like object
    term = SimpleNamespace(eta=None, real_space_cut=None, recip_space_cut=None, use_term="real")
    proc = S.EwaldProcessor(sub, scm, ewald_term=term, coefficient=1.0)
    np.testing.assert_allclose(proc.ewald_matrix, L.ewald_matrix(sub, scm, term="real")[0], rtol=0, atol=1e-12)
