"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

from tests import models as M

pytestmark = pytest.mark.gpu

RTOL = 1e-10   # north-star float tolerance (relative); ATOL scales with the feature magnitude


def _oracle():
    from oracle import lmc_oracle as O
    return O


def _processors(kind, subspace, scm, coefs):
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    if kind == "expansion":
        return (S.ClusterExpansionProcessor(subspace, scm, coefs),
                O.ClusterExpansionProcessor(subspace, scm, coefs))
    it = L.cluster_interaction_tensors(subspace, coefs)
    return (S.ClusterDecompositionProcessor(subspace, scm, it),
            O.ClusterDecompositionProcessor(subspace, scm, it))


@pytest.mark.parametrize("kind", ["expansion", "decomposition"])
@pytest.mark.parametrize("case", ["fcc2", "fcc4", "rs3"])
def test_full_and_delta_features(cuda_device, kind, case):
    if case.startswith("fcc"):
        sub = M.fcc_subspace()
        n = int(case[3:])
    else:
        sub = M.rocksalt_subspace(anions=("O2-", "F-"))
        n = 3
    scm = np.eye(3, dtype=int) * n
    rng = np.random.default_rng(5)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    gpu, ora = _processors(kind, sub, scm, coefs)
    W = 24
    occ = M.random_occupancies(sub, scm, W, seed=3)
    full_g = gpu.compute_feature_vector_batch(occ)
    full_o = np.array([ora.compute_feature_vector(o) for o in occ])
    scale = np.abs(full_o).max()
    np.testing.assert_allclose(full_g, full_o, rtol=RTOL, atol=RTOL * scale)
    spaces = sub.allowed_species(scm)
    active = [i for i, s in enumerate(spaces) if len(s) > 1]
    for k in (1, 2, 3):
        sites = rng.choice(active, size=(W, k))
        codes = np.zeros((W, k), dtype=np.int32)
        for w in range(W):
            cur = occ[w].copy()
            for j in range(k):
                s = sites[w, j]
                ch = [c for c in range(len(spaces[s])) if c != cur[s]]
                codes[w, j] = rng.choice(ch)
                cur[s] = codes[w, j]
        d_g = gpu.compute_feature_vector_change_batch(occ, sites, codes)
        d_o = np.array([ora.compute_feature_vector_change(occ[w], list(zip(sites[w], codes[w])))
                        for w in range(W)])
        np.testing.assert_allclose(d_g, d_o, rtol=RTOL, atol=RTOL * scale)


def _run_both(ens_g, ens_o_factory, usher_name, W, nsteps, thin, occ0, seeds, T=None, wl=None,
              usher_kwargs=None, group_size=0, bias=None):
    """bias: (name, kwargs) of an MCBias for both sides"""
    import smol_b200 as S
    O = _oracle()
    usher_kwargs = dict(usher_kwargs or {})
    if bias is not None:
        usher_kwargs.update(bias_type=bias[0], bias_kwargs=bias[1])
    oracle_composite = usher_kwargs.pop("oracle_composite", None)
    oracle_multistep = usher_kwargs.pop("oracle_multistep", None)
    if wl is None:
        smp = S.Sampler.from_ensemble(ens_g, T, step_type=usher_name, nwalkers=W, seeds=list(seeds),
                                      group_size=group_size, **usher_kwargs)
    else:
        smp = S.Sampler.from_ensemble(ens_g, wl["min"], wl["max"], wl["bin"], step_type=usher_name,
                                      kernel_type="WangLandau", nwalkers=W, seeds=list(seeds),
                                      check_period=wl["check"], flatness=wl["flatness"],
                                      group_size=group_size, **usher_kwargs)
    smp.run(nsteps, occ0, thin_by=thin)
    if oracle_composite is not None:
        usher_kwargs["oracle_composite"] = oracle_composite
    if oracle_multistep is not None:
        usher_kwargs["oracle_multistep"] = oracle_multistep
    kernels = []
    for w in range(W):
        ens_o = ens_o_factory()
        subl = ens_o.sublattices
        if usher_name == "swap":
            ush = O.Swap(subl, usher_kwargs.get("sublattice_probabilities"))
        elif usher_name == "flip":
            ush = O.Flip(subl, usher_kwargs.get("sublattice_probabilities"))
        elif usher_name == "multistep":
            sub_name, lens, probs = usher_kwargs["oracle_multistep"]
            ush = O.MultiStep(subl, (O.Flip if sub_name == "flip" else O.Swap)(subl), lens, probs)
        elif usher_name == "composite":
            spec = usher_kwargs["oracle_composite"]     # [(name, sublattice indices or None, probabilities)], weights
            subs = []
            for name, idx, probs in spec[0]:
                sl = subl if idx is None else [subl[i] for i in idx]
                subs.append((O.Flip if name == "flip" else O.Swap)(sl, probs))
            ush = O.Composite(subl, subs, spec[1])
        else:
            ush = O.TableFlip(subl, usher_kwargs["flip_table"],
                              swap_weight=usher_kwargs.get("swap_weight", 0.1))
        if wl is None:
            ob = None
            if bias is not None:
                ob = (O.FugacityBias(subl, bias[1]["fugacity_fractions"]) if "fugacity" in bias[0].lower()
                      else O.SquareHyperplaneBias(subl, **bias[1]) if "hyperplane" in bias[0].lower()
                      else O.SquareChargeBias(subl, **bias[1]))
            kernels.append(O.Metropolis(ens_o, ush, T, seed=int(seeds[w]), walker=w, bias=ob))
        else:
            kernels.append(O.WangLandau(ens_o, ush, wl["min"], wl["max"], wl["bin"],
                                        flatness=wl["flatness"], check_period=wl["check"],
                                        seed=int(seeds[w]), walker=w))
    ref = O.run_sampler(kernels, occ0, nsteps, thin)
    return smp, ref, kernels


def _compare_traces(smp, ref):
    s = smp.samples
    np.testing.assert_array_equal(s.get_occupancies(flat=False), ref["occupancy"])   # bit exact
    np.testing.assert_array_equal(s.get_trace_value("accepted", flat=False), ref["accepted"])
    np.testing.assert_array_equal(s.get_trace_value("n_accepted", flat=False), ref["n_accepted"])
    scale = np.abs(ref["features"]).max()
    np.testing.assert_allclose(s.get_feature_vectors(flat=False), ref["features"], rtol=RTOL,
                               atol=RTOL * scale)
    np.testing.assert_allclose(s.get_enthalpies(flat=False), ref["enthalpy"], rtol=RTOL,
                               atol=RTOL * np.abs(ref["enthalpy"]).max())


@pytest.mark.parametrize("kind", ["decomposition", "expansion"])
@pytest.mark.parametrize("n,group", [(2, 0), (4, 8), (4, 32)])
def test_canonical_swap_trajectory(cuda_device, kind, n, group):
    import smol_b200 as S
    O = _oracle()
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * n
    coefs = M.fcc_coefs(sub)
    gpu_p, ora_p = _processors(kind, sub, scm, coefs)
    ens_g = S.Ensemble(gpu_p)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices))

    W = 6
    occ0 = M.random_occupancies(sub, scm, W, seed=1, balanced=True)
    seeds = np.arange(100, 100 + W)
    smp, ref, _ = _run_both(ens_g, ens_o, "swap", W, 400, 20, occ0, seeds, T=1000.0, group_size=group)
    _compare_traces(smp, ref)
    assert 0 < smp.samples.step_efficiency() <= 1


@pytest.mark.parametrize("factorize", ["1", "gather", "spec", "spec-wide", "spec-env", "spec-wide-env", "0", "random"])
def test_semigrand_ewald_flip_trajectory(cuda_device, factorize, monkeypatch):
    """factorize=1: M = q q^T x K, Ewald through the per-walker potential cache (the default); gather: the
    same matrix with one row of K gathered per flip; spec / spec-wide: potential cache inside the speculative
    kernel (128- and 448-thread blocks); 0: generic transposed-row gather; random: a symmetric matrix that
    does NOT factorise (the library must detect that and fall back)."""
    monkeypatch.setenv("LMC_EWALD_FACTORIZE", "0" if factorize == "0" else "1")
    kw = {}
    if factorize == "gather":
        kw = dict(ewald_field=False)
    elif factorize.startswith("spec"):
        kw = dict(spec_mode=2, ewald_field=True, spec_env=factorize.endswith("env"))   # environment words / gathers
        monkeypatch.setenv("LMC_SPEC_WIDE", "1" if "wide" in factorize else "0")
    else:
        kw = dict(spec_mode=1)
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(11)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    ewm, ewi = L.ewald_matrix(sub, scm)
    if factorize == "random":
        r = np.random.default_rng(0).normal(0, 1.0, ewm.shape)
        ewm = 0.5 * (r + r.T)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.1, ewald_matrix=ewm, ewald_inds=ewi))
    mus = {"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2}
    ens_g = S.Ensemble(comp, chemical_potentials=mus)
    ora_p = O.CompositeProcessor([O.ClusterDecompositionProcessor(sub, scm, it),
                                  O.EwaldProcessor(ewm, ewi, 0.1)])

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)

    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=2)
    seeds = np.arange(7, 7 + W)
    smp, ref, _ = _run_both(ens_g, ens_o, "flip", W, 300, 10, occ0, seeds, T=1500.0, usher_kwargs=kw)
    _compare_traces(smp, ref)
    assert smp.ewald_cache_in_use == (factorize == "1" or factorize.startswith("spec"))
    # a second run continues the chains (potential cache rebuilt from the occupancies, then kept current)
    smp.run(100, thin_by=10)
    assert smp.samples.num_samples == 40


def test_ewald_potential_cache_stays_consistent(cuda_device):
    """after thousands of accepted flips the incrementally updated cache equals the one rebuilt from the
    occupancies, and the running Ewald feature equals the full evaluation (drift check in the spirit of
    Processor.compute_average_drift, processor/base.py:208-243)"""
    import smol_b200 as S
    from smol_b200 import lattice as L
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 4
    rng = np.random.default_rng(23)
    it = L.cluster_interaction_tensors(sub, rng.normal(0, 0.02, sub.num_corr_functions))
    ewm, ewi = L.ewald_matrix(sub, scm)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.05, ewald_matrix=ewm, ewald_inds=ewi))
    ens = S.Ensemble(comp, chemical_potentials={"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2})
    W = 64
    occ0 = M.random_occupancies(sub, scm, W, seed=3)
    for spec in (1, 2):
        smp = S.Sampler.from_ensemble(ens, 8000.0, step_type="flip", nwalkers=W, seeds=list(range(W)),
                                      spec_mode=spec, ewald_field=True)
        smp.run(20000, occ0, thin_by=2000)
        assert smp.samples.step_efficiency() > 0.05         # more than a thousand cache updates per walker
        eng = smp.engine
        fresh = eng.ewald_field(smp._occ_dev).cpu().numpy()
        kept = smp._ew_field.cpu().numpy()
        np.testing.assert_allclose(kept, fresh, rtol=0, atol=1e-10 * np.abs(fresh).max())
        feat, enth = eng.full_features(smp._occ_dev)
        last = smp.samples.get_feature_vectors(flat=False)[-1]
        np.testing.assert_allclose(last, feat.cpu().numpy(), rtol=1e-10, atol=1e-10 * np.abs(last).max())


@pytest.mark.parametrize("mode", ["classic", "spec", "spec-env"])
def test_canonical_ewald_swap_trajectory(cuda_device, mode):
    """canonical swaps with an Ewald term through the potential cache: flip 2 of a swap sees the cache
    shifted by flip 1 (one element of the site kernel); classic and speculative kernels (environment words / gathers)"""
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 3
    rng = np.random.default_rng(17)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    ewm, ewi = L.ewald_matrix(sub, scm)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.1, ewald_matrix=ewm, ewald_inds=ewi))
    ens_g = S.Ensemble(comp)
    ora_p = O.CompositeProcessor([O.ClusterDecompositionProcessor(sub, scm, it), O.EwaldProcessor(ewm, ewi, 0.1)])

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices))

    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=6)
    seeds = np.arange(40, 40 + W)
    smp, ref, _ = _run_both(ens_g, ens_o, "swap", W, 520, 13, occ0, seeds, T=2500.0,
                            usher_kwargs=dict(spec_mode=2 if mode.startswith("spec") else 1, spec_env=mode.endswith("env")))
    _compare_traces(smp, ref)
    assert smp.ewald_cache_in_use and 0 < smp.samples.step_efficiency() < 1


@pytest.mark.parametrize("wl_arrays", ["smem", "global", "pipeline", "speculative", "merged"])
def test_wang_landau_flip_trajectory(cuda_device, wl_arrays, monkeypatch):
    """smem: classic kernel, the walker's entropy / histogram in shared memory during a launch; global: kept in
    HBM / L2 (LMC_WL_GLOBAL, the path of very fine windows); pipeline / speculative: the warp-specialised kernel
    of lmc_wl.cuh with one decision warp or three (depth-2 speculation) over the classic records; merged: its
    merged-record form (one decision warp evaluates the three candidates of the speculation, features from the
    per-feature difference table: the default for few walkers)"""
    if wl_arrays == "global":
        monkeypatch.setenv("LMC_WL_GLOBAL", "1")
    monkeypatch.setenv("LMC_WL2", {"pipeline": "1", "speculative": "3", "merged": "4"}.get(wl_arrays, "0"))
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 3
    coefs = M.fcc_coefs(sub, seed=7)
    it = L.cluster_interaction_tensors(sub, coefs)
    gpu_p = S.ClusterDecompositionProcessor(sub, scm, it)
    ora_p = O.ClusterDecompositionProcessor(sub, scm, it)
    ens_g = S.Ensemble(gpu_p)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices))

    W = 5
    occ0 = M.random_occupancies(sub, scm, W, seed=4)
    e0 = np.array([ora_p.compute_property(o) for o in occ0])
    lo, hi = e0.min() - 2.0, e0.max() + 2.0
    wl = dict(min=lo, max=hi, bin=(hi - lo) / 37.3, check=50, flatness=0.3)
    seeds = np.arange(40, 40 + W)
    smp, ref, kernels = _run_both(ens_g, ens_o, "flip", W, 1500, 50, occ0, seeds, wl=wl)
    _compare_traces(smp, ref)
    st = smp.wang_landau_state
    for w, k in enumerate(kernels):
        np.testing.assert_array_equal(st["histogram"][w], k._histogram)
        np.testing.assert_array_equal(st["occurrences"][w], k._occurrences)
        np.testing.assert_allclose(st["entropy"][w], k._entropy, rtol=1e-13, atol=0)
        assert st["mod_factor"][w] == k._m
        np.testing.assert_allclose(st["mean_features"][w], k._mean_features, rtol=RTOL,
                                   atol=RTOL * np.abs(k._mean_features).max())
    assert (st["mod_factor"] < 1.0).any(), "flatness was never reached; weak test"
    # per-sample Wang-Landau traces (wanglandau.py:247-251): the arrays after the sampled step, the
    # modification factor as it was before that step's flatness check
    g = smp.samples.get_trace_value
    np.testing.assert_array_equal(g("histogram", flat=False), ref["histogram"])
    np.testing.assert_array_equal(g("occurrences", flat=False), ref["occurrences"])
    np.testing.assert_allclose(g("entropy", flat=False), ref["entropy"], rtol=1e-13, atol=0)
    np.testing.assert_array_equal(g("mod_factor", flat=False), ref["mod_factor"])
    np.testing.assert_allclose(g("cumulative_mean_features", flat=False), ref["cumulative_mean_features"], rtol=RTOL,
                               atol=RTOL * np.abs(ref["cumulative_mean_features"]).max())
    assert len(np.unique(ref["mod_factor"][:, 0, 0])) > 1 and ref["histogram"].shape == (30, W, len(kernels[0]._levels))
    # the light variants keep the container free of the big arrays
    smp2 = S.Sampler.from_ensemble(ens_g, wl["min"], wl["max"], wl["bin"], step_type="flip", kernel_type="WangLandau",
                                   nwalkers=W, seeds=list(seeds), check_period=50, flatness=0.3, wl_trace="no_means")
    smp2.run(500, occ0, thin_by=50)
    assert "cumulative_mean_features" not in smp2.samples.traced_values and "entropy" in smp2.samples.traced_values
    np.testing.assert_array_equal(smp2.samples.get_trace_value("histogram", flat=False), ref["histogram"][:10])


@pytest.mark.parametrize("group,factorize", [(8, "1"), (32, "1"), (32, "gather"), (32, "0"), (0, "spec"), (0, "spec-noewald"),
                                             (0, "spec-expansion"), (0, "spec-thin3"), (0, "classic")])
def test_table_flip_ewald_semigrand_trajectory(cuda_device, group, factorize, monkeypatch):
    """factorize=1: potential cache (flips of one step chained through elements of the site kernel);
    gather: one row of K per flip; 0: generic matrix rows; spec: speculative batches (csrc/lmc_spec_tf.cuh: eight steps
    per warp, flips of a step patched into the gathers of the later ones), with and without the Ewald term; classic:
    the automatic choice switched off; spec-expansion: correlation-function basis (several functions per orbit);
    spec-thin3: sampling intervals shorter than a batch"""
    monkeypatch.setenv("LMC_EWALD_FACTORIZE", "0" if factorize == "0" else "1")
    spec = factorize.startswith("spec")
    noew = factorize == "spec-noewald"
    expansion = factorize == "spec-expansion"
    nsteps, thin = (120, 3) if factorize == "spec-thin3" else (400, 20)
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 3
    rng = np.random.default_rng(21)
    coefs = rng.normal(0, 0.03, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    ewm, ewi = L.ewald_matrix(sub, scm)
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterExpansionProcessor(sub, scm, coefs) if expansion else S.ClusterDecompositionProcessor(sub, scm, it))
    if not noew:
        comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.05, ewald_matrix=ewm, ewald_inds=ewi))
    mus = {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0}
    ens_g = S.Ensemble(comp, chemical_potentials=mus)
    ora_p = O.CompositeProcessor([O.ClusterExpansionProcessor(sub, scm, coefs) if expansion
                                  else O.ClusterDecompositionProcessor(sub, scm, it)] +
                                 ([] if noew else [O.EwaldProcessor(ewm, ewi, 0.05)]))

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)

    table = [[-1, 1, 0, 2, -2], [0, -1, 1, 1, -1]]
    W, ncell = 3, 27
    occ0 = np.zeros((W, 2 * ncell), dtype=np.int32)
    for w in range(W):
        cat = np.array([0] * 20 + [1] * 4 + [2] * 3)   # 20 Li+, 4 Mn3+, 3 Ti4+  (+44)
        ani = np.array([0] * 17 + [1] * 10)            # 17 O2-, 10 F-          (-44)
        occ0[w, :ncell] = rng.permutation(cat)
        occ0[w, ncell:] = rng.permutation(ani)
    seeds = np.arange(900, 900 + W)
    kw = dict(flip_table=table, swap_weight=0.2, ewald_field=factorize != "gather" if factorize != "0" else "auto")
    if spec:
        kw.update(spec_mode=2, ewald_field="auto" if noew else True)
    elif factorize == "classic":
        kw.update(spec_mode=1, ewald_field=True)
    smp, ref, _ = _run_both(ens_g, ens_o, "table_flip", W, nsteps, thin, occ0, seeds, T=2000.0, usher_kwargs=kw, group_size=group)
    _compare_traces(smp, ref)
    # charge neutrality is conserved by construction of the table
    occ = smp.samples.get_occupancies(flat=True)
    q_cat = np.array([1, 3, 4])[occ[:, :ncell]].sum(axis=1)
    q_ani = np.array([-2, -1])[occ[:, ncell:]].sum(axis=1)
    assert np.all(q_cat + q_ani == 0)
    assert smp.samples.step_efficiency() > 0


def test_ewald_matrix_gpu_backend_equals_numpy(cuda_device):
    """lmc_ewald_site_kernel (reciprocal + real space sums per site pair) against the numpy evaluation of the
    same sums (cofe/extern/ewald.py:102-177 gets them from pymatgen); a skewed supercell and two sublattices"""
    from smol_b200 import lattice as L
    for anions, scm in ((("O2-",), np.eye(3, dtype=int) * 3),
                        (("O2-", "F-"), np.array([[2, 1, 0], [0, 2, 0], [0, 0, 3]]))):
        sub = M.rocksalt_subspace(anions=anions)
        a, ia = L.ewald_matrix(sub, scm, backend="numpy")
        b, ib = L.ewald_matrix(sub, scm, backend="gpu")
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_allclose(b, a, rtol=0, atol=1e-12 * np.abs(a).max())
        assert np.abs(b - b.T).max() == 0.0


def test_page_locked_tensor_input_equals_array_input(cuda_device):
    """initial occupancies given as a page-locked int32 torch tensor are copied straight from the caller's
    buffer; the chains are those of the ndarray path and the caller's data is left untouched (sampler.py:401)"""
    import torch
    import smol_b200 as S
    from smol_b200 import lattice as L
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 4
    it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub))
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it))
    W = 6
    occ0 = M.random_occupancies(sub, scm, W, seed=2, balanced=True)
    pinned = torch.from_numpy(occ0.astype(np.int32)).pin_memory()
    out = []
    for src in (occ0, pinned):
        smp = S.Sampler.from_ensemble(ens, 1500.0, step_type="swap", nwalkers=W, seeds=list(range(W)))
        smp.run(640, src, thin_by=64)
        out.append((smp.samples.get_occupancies(flat=False), smp.samples.get_enthalpies(flat=False)))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(pinned.numpy(), occ0)


# ---------------------------------------------------------------------------------------------
# distance processors (processor/distance.py): the sampler minimises the distance to a target vector
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,usher", [("correlation", "swap"), ("correlation", "flip"), ("interaction", "swap")])
def test_distance_processor_trajectory(cuda_device, kind, usher):
    """features [L, |f_i - target_i| ...] per supercell; the target is the vector of another occupancy of the
    same composition, so exact matches (L > 0) occur; single-call API and the whole chain against the oracle
    restatement of distance.py / evaluator.pyx:319-435"""
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(12)
    W = 6
    occ0 = M.random_occupancies(sub, scm, W + 1, seed=21, balanced=True)
    tgt_occ, occ0 = occ0[-1], occ0[:W]
    if kind == "correlation":
        base = O.ClusterExpansionProcessor(sub, scm, np.zeros(sub.num_corr_functions))
        target = base.compute_feature_vector(tgt_occ) / base.size
        weights = rng.uniform(0.5, 1.5, len(target) - 1)
        gpu_p = S.CorrelationDistanceProcessor(sub, scm, target_vector=target, match_weight=0.7, match_tol=1e-8,
                                               target_weights=weights)
        ora_p = O.CorrelationDistanceProcessor(sub, scm, target_vector=target, match_weight=0.7, match_tol=1e-8,
                                               target_weights=weights)
    else:
        it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub))
        base = O.ClusterDecompositionProcessor(sub, scm, it)
        target = base.compute_feature_vector(tgt_occ) / base.size
        weights = rng.uniform(0.5, 1.5, len(target) - 1)
        gpu_p = S.ClusterInteractionDistanceProcessor(sub, scm, it, target_vector=target, match_weight=0.7,
                                                      target_weights=weights)
        ora_p = O.ClusterInteractionDistanceProcessor(sub, scm, it, target_vector=target, match_weight=0.7,
                                                      target_weights=weights)
    # single calls
    for o in (occ0[0], tgt_occ):
        np.testing.assert_allclose(gpu_p.compute_feature_vector(o), ora_p.compute_feature_vector(o), rtol=0, atol=1e-12)
    assert gpu_p.compute_feature_vector(tgt_occ)[0] > 0            # the target occupancy matches every diameter
    flips = [(1, int(1 - occ0[0][1])), (5, int(1 - occ0[0][5]))]
    np.testing.assert_allclose(gpu_p.compute_feature_vector_change(occ0[0], flips),
                               ora_p.compute_feature_vector_change(occ0[0], flips), rtol=0, atol=1e-12)
    ens_g = S.Ensemble(gpu_p)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices))

    seeds = np.arange(800, 800 + W)
    # kB T comparable to the distance differences: accepts and rejects both occur, matches are found
    smp, ref, _ = _run_both(ens_g, ens_o, usher, W, 1200, 40, occ0, seeds, T=400.0)
    _compare_traces(smp, ref)
    feats = smp.samples.get_feature_vectors(flat=False)
    assert 0 < smp.samples.step_efficiency() < 1
    if usher == "swap":
        assert (feats[:, :, 0] > 0).any(), "no exact match of the point/pair terms was ever reached; weak test"
    with pytest.raises(ValueError):
        S.CorrelationDistanceProcessor(sub, scm, target_vector=np.zeros(sub.num_corr_functions), match_weight=-1.0)


# ---------------------------------------------------------------------------------------------
# composite usher (mcusher.py:307-394)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hybrid", [False, True])
def test_composite_flip_swap_trajectory(cuda_device, hybrid):
    """mixed flip / swap steps; hybrid: flips only on the anion sublattice, swaps only on the cations (each
    sub-usher built on its own sublattice, 'hybrid ensembles' of mcusher.py:310-311)"""
    import smol_b200 as S
    from smol_b200 import usher as U
    O = _oracle()
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 3
    rng = np.random.default_rng(8)
    coefs = rng.normal(0, 0.03, sub.num_corr_functions)
    gpu_p, ora_p = _processors("expansion", sub, scm, coefs)
    mus = {"Li+": 0.0, "Mn3+": 0.2, "Ti4+": -0.1, "O2-": 0.05, "F-": 0.0}
    ens_g = S.Ensemble(gpu_p, chemical_potentials=mus)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)

    sls = ens_g.sublattices
    act = [s for s in sls if s.is_active]
    icat = next(i for i, s in enumerate(act) if "Li+" in s.species)
    iani = 1 - icat
    if hybrid:
        ushers = [U.Flip([act[iani]]), U.Swap([act[icat]])]
        ospec = ([("flip", [iani], None), ("swap", [icat], None)], [1, 3])
        weights = [1, 3]
    else:
        ushers = ["flip", U.Swap(sls, sublattice_probabilities=[0.25, 0.75])]
        ospec = ([("flip", None, None), ("swap", None, [0.25, 0.75])], [2, 1])
        weights = [2, 1]
    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=14)
    seeds = np.arange(300, 300 + W)
    smp, ref, _ = _run_both(ens_g, ens_o, "composite", W, 420, 14, occ0, seeds, T=2500.0,
                            usher_kwargs=dict(mcushers=ushers, mcusher_weights=weights, oracle_composite=ospec))
    _compare_traces(smp, ref)
    occ = smp.samples.get_occupancies(flat=False)                    # [S, W, N]
    counts = np.stack([(occ[:, :, act[icat].sites] == c).sum(2) for c in range(3)], 2)   # [S, W, 3]
    if hybrid:   # swaps conserve every walker's cation composition, flips change the anions
        assert (counts == counts[:1]).all()
        assert ((occ[:, :, act[iani].sites] == 0).sum(2) != (occ[:1, :, act[iani].sites] == 0).sum(2)).any()
    else:
        assert (counts != counts[:1]).any()
    assert 0 < smp.samples.step_efficiency() < 1


@pytest.mark.parametrize("sub_usher,lens,probs", [("flip", [1, 3, 4], [0.2, 0.5, 0.3]), ("swap", 2, None),
                                                  ("swap", [1, 2], None)], ids=["flip134", "swap2", "swap12"])
def test_multistep_trajectory(cuda_device, sub_usher, lens, probs):
    """chained proposals (mcusher.py:284-304) on a SMALL cell so that collisions with already changed sites --
    dropped proposals -- and swap partners picked in the swapped configuration occur often"""
    import smol_b200 as S
    O = _oracle()
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(10)
    coefs = rng.normal(0, 0.03, sub.num_corr_functions)
    gpu_p, ora_p = _processors("expansion", sub, scm, coefs)
    mus = {"Li+": 0.0, "Mn3+": 0.2, "Ti4+": -0.1, "O2-": 0.05, "F-": 0.0}
    ens_g = S.Ensemble(gpu_p, chemical_potentials=mus)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)

    W = 5
    occ0 = M.random_occupancies(sub, scm, W, seed=15)
    seeds = np.arange(700, 700 + W)
    smp, ref, kernels = _run_both(ens_g, ens_o, "multistep", W, 600, 15, occ0, seeds, T=3000.0,
                                  usher_kwargs=dict(mcusher=sub_usher, step_lengths=lens, step_probabilities=probs,
                                                    oracle_multistep=(sub_usher, lens, probs)))
    _compare_traces(smp, ref)
    # the chain really produced multi-site steps and dropped colliding proposals
    k = O.Metropolis(ens_o(), O.MultiStep(ens_o().sublattices, (O.Flip if sub_usher == "flip" else O.Swap)(ens_o().sublattices),
                                          lens, probs), 3000.0, seed=1, walker=0)
    occ = occ0[0].copy()
    sizes = [len(k.single_step(occ).step) for _ in range(300)]
    per = 1 if sub_usher == "flip" else 2
    top = max(lens) if isinstance(lens, list) else lens
    assert max(sizes) == top * per and len(set(sizes)) > 1
    assert 0 < smp.samples.step_efficiency() < 1


# ---------------------------------------------------------------------------------------------
# bias terms (smol/moca/kernel/bias.py) in the Metropolis exponent
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bias", [("square-charge", dict(penalty=0.5)), ("SquareChargeBias", dict(penalty=0.02)),
                                  ("fugacity", None),
                                  ("square-hyperplane", dict(hyperplane_normals=[[1, 3, 4, -2, -1], [1, -2, 0, 0, 0]],
                                                             hyperplane_intercepts=[0, 3], penalty=0.05))],
                         ids=["charge0.5", "charge0.02", "fugacity", "hyperplanes"])
@pytest.mark.parametrize("group", [0, 8])
def test_biased_semigrand_flip_trajectory(cuda_device, bias, group):
    """single flips on cation AND anion sublattices with a charge penalty / fugacity fractions: the bias
    change enters the exponent (metropolis.py:43-44) and the running bias is traced (base.py:362-363)"""
    import smol_b200 as S
    O = _oracle()
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 3
    rng = np.random.default_rng(4)
    coefs = rng.normal(0, 0.03, sub.num_corr_functions)
    gpu_p, ora_p = _processors("expansion", sub, scm, coefs)
    mus = {"Li+": 0.0, "Mn3+": 0.2, "Ti4+": -0.1, "O2-": 0.05, "F-": 0.0}
    ens_g = S.Ensemble(gpu_p, chemical_potentials=mus)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)

    name, kw = bias
    if kw is None:
        kw = dict(fugacity_fractions=[{"Li+": 0.5, "Mn3+": 0.25, "Ti4+": 0.25}, {"O2-": 0.75, "F-": 0.25}])
    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=31)
    seeds = np.arange(60, 60 + W)
    smp, ref, _ = _run_both(ens_g, ens_o, "flip", W, 360, 12, occ0, seeds, T=3000.0, group_size=group,
                            usher_kwargs=dict(spec_mode=1) if group == 0 else None, bias=(name, kw))
    _compare_traces(smp, ref)
    got = smp.samples.get_trace_value("bias", flat=False)
    scale = max(1.0, np.abs(ref["bias"]).max())
    np.testing.assert_allclose(got, ref["bias"], rtol=RTOL, atol=RTOL * scale)
    assert np.abs(np.diff(ref["bias"][:, 0, 0])).max() > 0, "bias never changed; weak test"
    # single-call API against the oracle's restatement
    ob = ref_bias = (O.FugacityBias(ens_o().sublattices, kw["fugacity_fractions"]) if "fug" in name
                     else O.SquareHyperplaneBias(ens_o().sublattices, **kw) if "hyper" in name
                     else O.SquareChargeBias(ens_o().sublattices, **kw))
    step = [(0, int((occ0[0, 0] + 1) % 3)), (30, int((occ0[0, 30] + 1) % 2))]
    np.testing.assert_allclose(smp.bias.compute_bias(occ0[0]), ob.compute_bias(occ0[0]), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(smp.bias.compute_bias_change(occ0[0], step), ref_bias.compute_bias_change(occ0[0], step),
                               rtol=1e-10, atol=1e-12)


def test_bias_rejected_where_the_reference_rejects_it(cuda_device):
    import smol_b200 as S
    from smol_b200 import lattice as L
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 2
    it = L.cluster_interaction_tensors(sub, np.random.default_rng(1).normal(0, 0.03, sub.num_corr_functions))
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it), chemical_potentials={"Li+": 0, "Mn3+": 0, "Ti4+": 0})
    with pytest.raises(ValueError, match="not a valid MCBias"):       # wanglandau.py:24
        S.Sampler.from_ensemble(ens, -1.0, 1.0, 0.1, kernel_type="WangLandau", step_type="flip", nwalkers=1, seeds=[1],
                                bias_type="square-charge")
    with pytest.raises(ValueError, match="Penalty"):
        S.Sampler.from_ensemble(ens, 1000.0, step_type="flip", nwalkers=1, seeds=[1], bias_type="square-charge",
                                bias_kwargs=dict(penalty=0.0))
    with pytest.raises(ValueError, match="add to one"):
        S.Sampler.from_ensemble(ens, 1000.0, step_type="flip", nwalkers=1, seeds=[1], bias_type="fugacity",
                                bias_kwargs=dict(fugacity_fractions=[{"Li+": 0.5, "Mn3+": 0.25, "Ti4+": 0.5}]))
    smp = S.Sampler.from_ensemble(ens, 1000.0, step_type="flip", nwalkers=2, seeds=[1, 2], bias_type="square-charge",
                                  spec_mode=2)
    with pytest.raises(RuntimeError, match="speculative"):
        smp.run(40, M.random_occupancies(sub, scm, 2, seed=1), thin_by=10)


# ---------------------------------------------------------------------------------------------
# speculative-batch kernel (csrc/lmc_spec.cuh): same chain as the classic kernel and the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["decomposition", "expansion"])
@pytest.mark.parametrize("n,T,thin", [(2, 1000.0, 20), (4, 300.0, 7), (4, 1000.0, 64), (3, 1e5, 13)])
@pytest.mark.parametrize("sg", ["4", "4G", "4E", "4L", "2", "1"])
def test_speculative_swap_trajectory(cuda_device, kind, n, T, thin, sg, monkeypatch):
    """low / medium / near-infinite temperature (acceptance ~0 .. ~1), sampling intervals that are not
    multiples of the batch, aliased 2x2x2 cell; spec_mode=2 forces the speculative kernel, 1 the classic
    one: both must reproduce the oracle chain bit for bit.  sg = lanes per speculated step, L = swap partner
    from sorted position lists instead of the rank select, E = environment words in L2 instead of occupancy gathers
    (``Sampler(spec_env=True)``, where the model has the tables); "4" takes the compact environment words in shared memory
    (csrc/lmc_spec_c64.cuh) where the model's records fit 64 bits per site, "4G" keeps the gathers."""
    import smol_b200 as S
    monkeypatch.setenv("LMC_SPEC_SG", sg[0])
    monkeypatch.setenv("LMC_SPEC_LISTS", "1" if sg.endswith("L") else "0")
    monkeypatch.setenv("LMC_SPEC_C64", "0" if sg.endswith("G") else "1")
    O = _oracle()
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * n
    coefs = M.fcc_coefs(sub)
    gpu_p, ora_p = _processors(kind, sub, scm, coefs)
    ens_g = S.Ensemble(gpu_p)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices))

    W = 5
    occ0 = M.random_occupancies(sub, scm, W, seed=3, balanced=True)
    seeds = np.arange(500, 500 + W)
    nsteps = thin * 30
    eng0 = S.Sampler.from_ensemble(ens_g, T, step_type="swap", nwalkers=W, seeds=list(seeds)).engine
    env0, c640 = eng0.env_launch_count(), eng0.c64_launch_count()
    smp, ref, _ = _run_both(ens_g, ens_o, "swap", W, nsteps, thin, occ0, seeds, T=T,
                            usher_kwargs=dict(spec_mode=2, spec_env=sg.endswith("E")))
    _compare_traces(smp, ref)
    # the compact-word kernel ran exactly where it is meant to: the default of the four-lane variant for these models
    assert smp.engine.model_info()[5] > 0
    assert (smp.engine.c64_launch_count() > c640) == (sg == "4")
    # the environment-word variant ran exactly where it is meant to (decomposition: merged records fit the lane chunks)
    took_env = smp.engine.env_launch_count() > env0
    assert took_env == (sg == "4E" and smp.engine.model_info()[4] > 0)
    if sg == "4E" and kind == "decomposition":
        assert took_env
    smp1 = S.Sampler.from_ensemble(ens_g, T, step_type="swap", nwalkers=W, seeds=list(seeds), spec_mode=1)
    smp1.run(nsteps, occ0, thin_by=thin)
    np.testing.assert_array_equal(smp1.samples.get_occupancies(flat=False),
                                  smp.samples.get_occupancies(flat=False))
    np.testing.assert_allclose(smp1.samples.get_enthalpies(flat=False), smp.samples.get_enthalpies(flat=False),
                               rtol=1e-12, atol=1e-12 * np.abs(ref["enthalpy"]).max())


@pytest.mark.parametrize("T", [400.0, 3000.0])
@pytest.mark.parametrize("sg", ["4", "4E", "2", "1"])
def test_speculative_semigrand_flip_trajectory(cuda_device, T, sg, monkeypatch):
    """ternary rocksalt cations, chemical potentials, single flips (no Ewald term): speculative kernel
    (4E = environment words with two bits per code and 64-bit lane chunks instead of occupancy gathers)"""
    import smol_b200 as S
    monkeypatch.setenv("LMC_SPEC_SG", sg[0])
    from smol_b200 import lattice as L
    O = _oracle()
    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 3
    rng = np.random.default_rng(5)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    mus = {"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2}
    ens_g = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it), chemical_potentials=mus)
    ora_p = O.ClusterDecompositionProcessor(sub, scm, it)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)

    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=8)
    seeds = np.arange(70, 70 + W)
    env0 = S.Sampler.from_ensemble(ens_g, T, step_type="flip", nwalkers=W, seeds=list(seeds)).engine.env_launch_count()
    smp, ref, _ = _run_both(ens_g, ens_o, "flip", W, 330, 11, occ0, seeds, T=T,
                            usher_kwargs=dict(spec_mode=2, spec_env=sg.endswith("E")))
    _compare_traces(smp, ref)
    assert (smp.engine.env_launch_count() > env0) == (sg == "4E")


@pytest.mark.parametrize("kind,n", [("decomposition", 4), ("expansion", 3)])
def test_speculative_semigrand_flip_binary_compact_words(cuda_device, kind, n):
    """binary FCC, chemical potentials, single flips: the FLIP instantiations of the compact-word kernel
    (csrc/lmc_spec_c64.cuh; three record pairs per lane as straight-line code on the 4x4x4 cell, the loop on 3x3x3)"""
    import smol_b200 as S
    O = _oracle()
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * n
    coefs = M.fcc_coefs(sub)
    gpu_p, ora_p = _processors(kind, sub, scm, coefs)
    mus = {"A": 0.0, "B": 0.15}
    ens_g = S.Ensemble(gpu_p, chemical_potentials=mus)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)

    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=14)
    seeds = np.arange(210, 210 + W)
    c640 = S.Sampler.from_ensemble(ens_g, 800.0, step_type="flip", nwalkers=W, seeds=list(seeds)).engine.c64_launch_count()
    smp, ref, _ = _run_both(ens_g, ens_o, "flip", W, 390, 13, occ0, seeds, T=800.0, usher_kwargs=dict(spec_mode=2))
    _compare_traces(smp, ref)
    assert smp.engine.c64_launch_count() > c640
    assert 0 < smp.samples.step_efficiency() < 1


def test_speculative_two_sublattice_swap(cuda_device):
    """cation AND anion sublattices active (different record counts per site class), correlation basis"""
    import smol_b200 as S
    O = _oracle()
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 3
    rng = np.random.default_rng(9)
    coefs = rng.normal(0, 0.03, sub.num_corr_functions)
    gpu_p, ora_p = _processors("expansion", sub, scm, coefs)
    ens_g = S.Ensemble(gpu_p)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices))

    W = 4
    occ0 = M.random_occupancies(sub, scm, W, seed=12)
    seeds = np.arange(33, 33 + W)
    smp, ref, _ = _run_both(ens_g, ens_o, "swap", W, 600, 25, occ0, seeds, T=1200.0, usher_kwargs=dict(spec_mode=2))
    _compare_traces(smp, ref)


def test_speculative_rejects_unsupported(cuda_device):
    import smol_b200 as S
    from smol_b200 import lattice as L
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 3
    it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub))
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it))
    occ0 = M.random_occupancies(sub, scm, 2, seed=4)
    e0 = 0.0
    smp = S.Sampler.from_ensemble(ens, e0 - 50.0, e0 + 50.0, 1.0, step_type="flip", kernel_type="WangLandau",
                                  nwalkers=2, seeds=[1, 2], spec_mode=2)
    with pytest.raises(RuntimeError, match="speculative"):
        smp.run(100, occ0, thin_by=10)


# ---------------------------------------------------------------------------------------------
# the kernel x usher matrix of tests/test_moca/test_kernel.py:22-94 under Wang-Landau
# ---------------------------------------------------------------------------------------------
def _wl_fcc_case(W=4, seed=4):
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 3
    it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub, seed=7))
    ens_g = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it))
    ora_p = O.ClusterDecompositionProcessor(sub, scm, it)

    def ens_o():
        return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices))
    occ0 = M.random_occupancies(sub, scm, W, seed=seed)
    e0 = np.array([ora_p.compute_property(o) for o in occ0])
    lo, hi = e0.min() - 2.0, e0.max() + 2.0
    return ens_g, ens_o, occ0, dict(min=lo, max=hi, bin=(hi - lo) / 29.3, check=40, flatness=0.3)


@pytest.mark.parametrize("usher", ["composite", "multistep", "swap"])
def test_wang_landau_usher_matrix_trajectory(cuda_device, usher):
    """WangLandau x {Composite(Flip, Swap), MultiStep(Flip), Swap}: occupancies bit-exact, entropy / histogram of
    every walker as the oracle's (kernel/wanglandau.py:186-266 over mcusher.py:203-394)"""
    ens_g, ens_o, occ0, wl = _wl_fcc_case()
    W = len(occ0)
    seeds = np.arange(60, 60 + W)
    kw = {}
    if usher == "composite":
        kw = dict(mcushers=["flip", "swap"], mcusher_weights=[2, 1],
                  oracle_composite=([("flip", None, None), ("swap", None, None)], [2, 1]))
    elif usher == "multistep":
        kw = dict(mcusher="flip", step_lengths=[1, 2, 3], oracle_multistep=("flip", [1, 2, 3], None))
    smp, ref, kernels = _run_both(ens_g, ens_o, usher, W, 1200, 40, occ0, seeds, wl=wl, usher_kwargs=kw)
    _compare_traces(smp, ref)
    st = smp.wang_landau_state
    for w, k in enumerate(kernels):
        np.testing.assert_array_equal(st["histogram"][w], k._histogram)
        np.testing.assert_array_equal(st["occurrences"][w], k._occurrences)
        np.testing.assert_allclose(st["entropy"][w], k._entropy, rtol=1e-13, atol=0)
        assert st["mod_factor"][w] == k._m
    assert (st["mod_factor"] < 1.0).any() and 0 < smp.samples.step_efficiency() < 1


@pytest.mark.parametrize("kernel", ["classic", "merged"])
def test_wang_landau_callable_mod_update(cuda_device, kernel, monkeypatch):
    """a callable mod_update (wanglandau.py:100-105): tabulated on the host, stepped through on the device"""
    import smol_b200 as S
    O = _oracle()
    monkeypatch.setenv("LMC_WL2", "0" if kernel == "classic" else "4")
    ens_g, ens_o, occ0, wl = _wl_fcc_case()
    W = len(occ0)
    seeds = np.arange(80, 80 + W)

    def update(m):
        return 0.4 * m + 1e-4          # not a division: the device cannot guess it

    smp = S.Sampler.from_ensemble(ens_g, wl["min"], wl["max"], wl["bin"], step_type="flip", kernel_type="WangLandau",
                                  nwalkers=W, seeds=list(seeds), check_period=wl["check"], flatness=wl["flatness"],
                                  mod_update=update, mod_factor=0.7)
    smp.run(1500, occ0, thin_by=50)
    kernels = [O.WangLandau(ens_o(), O.Flip(ens_o().sublattices), wl["min"], wl["max"], wl["bin"], flatness=wl["flatness"],
                            check_period=wl["check"], mod_factor=0.7, mod_update=update, seed=int(seeds[w]), walker=w)
               for w in range(W)]
    ref = O.run_sampler(kernels, occ0, 1500, 50)
    _compare_traces(smp, ref)
    st = smp.wang_landau_state
    for w, k in enumerate(kernels):
        assert st["mod_factor"][w] == k._m
        np.testing.assert_allclose(st["entropy"][w], k._entropy, rtol=1e-13, atol=0)
    np.testing.assert_array_equal(smp.samples.get_trace_value("mod_factor", flat=False), ref["mod_factor"])
    assert len(np.unique(st["mod_factor"])) >= 1 and (st["mod_factor"] < 0.7).any()


def test_ewald_term_matrices_on_the_gpu(cuda_device):
    """EwaldTerm.use_term parts from lmc_ewald_site_kernel (empty reciprocal / real sums) == the numpy sums"""
    from smol_b200 import lattice as L
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 2
    for term in ("reciprocal", "real", "point", "total"):
        g = L.ewald_matrix(sub, scm, backend="gpu", term=term)[0]
        c = L.ewald_matrix(sub, scm, backend="numpy", term=term)[0]
        np.testing.assert_allclose(g, c, rtol=0, atol=1e-12 * max(np.abs(c).max(), 1e-300))


@pytest.mark.parametrize("variant", ["thin1", "odd-periods", "rocksalt-semigrand"])
@pytest.mark.parametrize("kernel", ["pipeline", "speculative", "merged"])
def test_wang_landau_warp_specialised_edge_cases(cuda_device, kernel, variant, monkeypatch):
    """corners of lmc_wl.cuh: a sample boundary after every step (thin_by = 1: every batch is a single step followed
    by the trace rendezvous), thin_by / check_period / update_period that are odd and mutually prime (batches cut
    short by checks; update_period 2 = running means instead of sums), and a two-sublattice five-species cell with
    chemical potentials (ternary code radix, 152 merged records per site: the run-time record loop, chemical work)"""
    import smol_b200 as S
    from smol_b200 import lattice as L
    O = _oracle()
    monkeypatch.setenv("LMC_WL2", {"pipeline": "1", "speculative": "3", "merged": "4"}[kernel])
    if variant == "rocksalt-semigrand":
        sub = M.rocksalt_subspace(anions=("O2-", "F-"))
        scm = np.eye(3, dtype=int) * 2
        it = L.cluster_interaction_tensors(sub, np.random.default_rng(12).normal(0, 0.04, sub.num_corr_functions))
        mus = {"Li+": 0.0, "Mn3+": 0.2, "Ti4+": -0.1, "O2-": 0.05, "F-": 0.0}
        ens_g = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it), chemical_potentials=mus)
        ora_p = O.ClusterDecompositionProcessor(sub, scm, it)

        def ens_o():
            return O.Ensemble(ora_p, M.oracle_sublattices(O, ens_g.sublattices), chemical_potentials=mus)
        W = 3
        occ0 = M.random_occupancies(sub, scm, W, seed=9)
        e0 = np.array([np.dot(ens_o().natural_parameters, ens_o().compute_feature_vector(o)) for o in occ0])
        lo, hi = e0.min() - 1.5, e0.max() + 1.5
        wl = dict(min=lo, max=hi, bin=(hi - lo) / 17.2, check=30, flatness=0.2)
        nsteps, thin, kw = 900, 30, {}
    else:
        ens_g, ens_o, occ0, wl = _wl_fcc_case(W=3, seed=11)
        W = 3
        if variant == "thin1":
            nsteps, thin, kw = 240, 1, {}
        else:
            wl = dict(wl, check=7)
            nsteps, thin, kw = 13 * 60, 13, dict(update_period=2)
    seeds = np.arange(500, 500 + W)
    smp = S.Sampler.from_ensemble(ens_g, wl["min"], wl["max"], wl["bin"], step_type="flip", kernel_type="WangLandau",
                                  nwalkers=W, seeds=list(seeds), check_period=wl["check"], flatness=wl["flatness"], **kw)
    smp.run(nsteps, occ0, thin_by=thin)
    kernels = [O.WangLandau(ens_o(), O.Flip(ens_o().sublattices), wl["min"], wl["max"], wl["bin"], flatness=wl["flatness"],
                            check_period=wl["check"], seed=int(seeds[w]), walker=w, **kw) for w in range(W)]
    ref = O.run_sampler(kernels, occ0, nsteps, thin)
    _compare_traces(smp, ref)
    st = smp.wang_landau_state
    g = smp.samples.get_trace_value
    for w, k in enumerate(kernels):
        np.testing.assert_array_equal(st["histogram"][w], k._histogram)
        np.testing.assert_array_equal(st["occurrences"][w], k._occurrences)
        np.testing.assert_allclose(st["entropy"][w], k._entropy, rtol=1e-13, atol=0)
        assert st["mod_factor"][w] == k._m
        np.testing.assert_allclose(st["mean_features"][w], k._mean_features, rtol=RTOL, atol=RTOL * max(np.abs(k._mean_features).max(), 1.0))
    np.testing.assert_array_equal(g("histogram", flat=False), ref["histogram"])
    np.testing.assert_array_equal(g("mod_factor", flat=False), ref["mod_factor"])
    np.testing.assert_allclose(g("entropy", flat=False), ref["entropy"], rtol=1e-13, atol=0)
    assert 0 < smp.samples.step_efficiency() < 1
