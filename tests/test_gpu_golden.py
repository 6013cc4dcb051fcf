"""CUDA path vs golden vectors produced by the reference's own compiled Cython evaluators."""
import os

import numpy as np
import pytest

from tests import models as M

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))
CASES = {"fcc2": (lambda: M.fcc_subspace(), 2), "fcc3": (lambda: M.fcc_subspace(), 3),
         "rs2": (lambda: M.rocksalt_subspace(anions=("O2-", "F-")), 2)}


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_reference_golden_vectors(cuda_device, name):
    import smol_b200 as S
    from smol_b200 import lattice as L
    mk, n = CASES[name]
    sub = mk()
    scm = np.eye(3, dtype=int) * n
    coefs = GOLD[f"{name}_coefs"]
    occ, sites, codes = GOLD[f"{name}_occ"], GOLD[f"{name}_sites"], GOLD[f"{name}_codes"]
    ce = S.ClusterExpansionProcessor(sub, scm, coefs)
    cd = S.ClusterDecompositionProcessor(sub, scm, L.cluster_interaction_tensors(sub, coefs))
    for proc, key in ((ce, "corr"), (cd, "inter")):
        full = GOLD[f"{name}_full_{key}"]
        scale = np.abs(full).max()
        np.testing.assert_allclose(proc.compute_feature_vector_batch(occ), full, rtol=1e-10, atol=1e-10 * scale)
        np.testing.assert_allclose(proc.compute_feature_vector_change_batch(occ, sites, codes),
                                   GOLD[f"{name}_delta_{key}"], rtol=1e-10, atol=1e-10 * scale)
    if name == "rs2":
        ew = S.EwaldProcessor(sub, scm, ewald_matrix=GOLD["rs2_ewald_matrix"], ewald_inds=GOLD["rs2_ewald_inds"])
        scale = np.abs(GOLD["rs2_full_ewald"]).max()
        np.testing.assert_allclose(ew.compute_feature_vector_batch(occ)[:, 0], GOLD["rs2_full_ewald"], rtol=1e-10)
        np.testing.assert_allclose(ew.compute_feature_vector_change_batch(occ, sites, codes)[:, 0],
                                   GOLD["rs2_delta_ewald"], rtol=1e-10, atol=1e-10 * scale)


def test_processor_api_single_calls_and_errors(cuda_device):
    """Reference-style single calls (tests/test_moca/test_processor.py:175-231)."""
    import smol_b200 as S
    from smol_b200 import lattice as L
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 3
    coefs = M.fcc_coefs(sub)
    proc = S.ClusterDecompositionProcessor(sub, scm, L.cluster_interaction_tensors(sub, coefs))
    occ = M.random_occupancies(sub, scm, 1, seed=1)[0]
    rng = np.random.default_rng(0)
    for _ in range(5):
        s = int(rng.integers(27))
        new = 1 - occ[s]
        occ2 = occ.copy()
        occ2[s] = new
        dprop = proc.compute_property_change(occ, [(s, new)])
        assert dprop == pytest.approx(proc.compute_property(occ2) - proc.compute_property(occ), rel=1e-10, abs=1e-11)
        assert proc.compute_property_change(occ2, [(s, occ[s])]) == pytest.approx(-dprop, rel=1e-10, abs=1e-12)
        occ = occ2
    with pytest.raises(ValueError):
        proc.compute_feature_vector(["a"] * 27)
    fwd, rev = proc.compute_average_drift(iterations=200, seed=1)
    assert abs(fwd) < 1e-12 and abs(rev) < 1e-12
    assert np.array_equal(proc.compute_feature_vector_change(occ, []), np.zeros(8))
    with pytest.raises(ValueError):
        S.ClusterExpansionProcessor(sub, scm, coefs[:-1])


def test_cuda_matches_reference_known_answer_licabr(cuda_device):
    """the CUDA full-vector kernel on the reference's stored correlation vector (test_clusterspace.py:663-725)"""
    import smol_b200 as S
    from tests.test_oracle_golden import LICABR_EXPECTED, licabr_case
    sub, scm, occ = licabr_case()
    proc = S.ClusterExpansionProcessor(sub, scm, np.ones(sub.num_corr_functions))
    corr = proc.compute_feature_vector(occ) / sub.supercell_size(scm)
    np.testing.assert_allclose(corr, LICABR_EXPECTED, rtol=1e-12, atol=1e-14)


# (green on a B200 since the round-1 driver run, GPUTEST_r01.json)
def test_cuda_reproduces_trajectory_recorded_from_the_reference_python_stack(cuda_device):
    """semigrand flips on the 5-species rocksalt cell: the CUDA sampler against the step record of the reference's OWN
    Ensemble + ClusterDecompositionProcessor + Metropolis + Flip classes (tests/golden/ref_python_steps.npz,
    `fullref_rs2of_flip_*`, made by tests/golden/make_reference_python_golden.py with the generator scripted to the
    engine's Philox word positions)"""
    import importlib.util
    import smol_b200 as S
    path = os.path.join(os.path.dirname(__file__), "golden", "make_reference_python_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_python_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_steps.npz"))
    _, occ0 = mod.table_flip_model()
    sub, scm, _ = mod.processor_cases()["rs2of"]
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, mod.TF_INTERACTIONS()), chemical_potentials=dict(mod.TF_MUS))
    W = len(occ0)
    seeds = [int(gold[f"fullref_rs2of_flip_w{w}_meta"][0]) for w in range(W)]
    T = float(gold["fullref_rs2of_flip_w0_meta"][1])
    smp = S.Sampler.from_ensemble(ens, T, step_type="flip", nwalkers=W, seeds=seeds, spec_mode=1)
    smp.run(300, occ0, thin_by=25)
    occ = smp.samples.get_occupancies(flat=False)
    acc = smp.samples.get_trace_value("accepted", flat=False)
    nacc = smp.samples.get_trace_value("n_accepted", flat=False)
    enth = smp.samples.get_enthalpies(flat=False)
    for w in range(W):
        key = f"fullref_rs2of_flip_w{w}"
        np.testing.assert_array_equal(occ[:, w], gold[key + "_snaps"])                       # bit exact
        np.testing.assert_array_equal(acc[:, w, 0], gold[key + "_acc"][24::25])
        np.testing.assert_array_equal(nacc[:, w], gold[key + "_acc"].reshape(12, 25).sum(axis=1))
        h0 = float(np.dot(gold[key + "_natural"], gold[key + "_feat0"]))
        want = h0 + np.cumsum(gold[key + "_dh"] * gold[key + "_acc"])[24::25]
        np.testing.assert_allclose(enth[:, w, 0], want, rtol=1e-10, atol=1e-10 * max(1.0, abs(h0)))


@pytest.mark.parametrize("name", ["fcc3", "fcc421", "rs2of"])
def test_cuda_matches_records_of_the_reference_python_processors(cuda_device, name):
    """CUDA full vectors and 1..3-flip changes against the outputs recorded from the reference's OWN
    ClusterExpansionProcessor / ClusterDecompositionProcessor (tests/golden/ref_python_steps.npz `proc_*`; the oracle
    reproduces the same records on CPU and the CUDA path equals the oracle in test_gpu_parity.py)"""
    import importlib.util
    import smol_b200 as S
    from smol_b200 import lattice as L
    path = os.path.join(os.path.dirname(__file__), "golden", "make_reference_python_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_python_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_steps.npz"))
    sub, scm, coefs = mod.processor_cases()[name]
    occs, flips = mod.processor_flips(sub, scm, seed=3)
    it = L.cluster_interaction_tensors(sub, coefs)
    atol = 2e4 * np.finfo(float).eps * sub.supercell_size(scm)
    for tag, proc in (("ce", S.ClusterExpansionProcessor(sub, scm, coefs)),
                      ("cd", S.ClusterDecompositionProcessor(sub, scm, it))):
        key = f"proc_{name}_{tag}"
        np.testing.assert_allclose(proc.compute_feature_vector_batch(occs), gold[key + "_full"], rtol=1e-10, atol=atol)
        delta = np.array([proc.compute_feature_vector_change(o, f) for o, f in zip(occs, flips)])
        np.testing.assert_allclose(delta, gold[key + "_delta"], rtol=1e-10, atol=atol)
