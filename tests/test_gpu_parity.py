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
              usher_kwargs=None, group_size=0):
    import smol_b200 as S
    O = _oracle()
    usher_kwargs = usher_kwargs or {}
    if wl is None:
        smp = S.Sampler.from_ensemble(ens_g, T, step_type=usher_name, nwalkers=W, seeds=list(seeds),
                                      group_size=group_size, **usher_kwargs)
    else:
        smp = S.Sampler.from_ensemble(ens_g, wl["min"], wl["max"], wl["bin"], step_type=usher_name,
                                      kernel_type="WangLandau", nwalkers=W, seeds=list(seeds),
                                      check_period=wl["check"], flatness=wl["flatness"],
                                      group_size=group_size, **usher_kwargs)
    smp.run(nsteps, occ0, thin_by=thin)
    kernels = []
    for w in range(W):
        ens_o = ens_o_factory()
        subl = ens_o.sublattices
        if usher_name == "swap":
            ush = O.Swap(subl, usher_kwargs.get("sublattice_probabilities"))
        elif usher_name == "flip":
            ush = O.Flip(subl, usher_kwargs.get("sublattice_probabilities"))
        else:
            ush = O.TableFlip(subl, usher_kwargs["flip_table"],
                              swap_weight=usher_kwargs.get("swap_weight", 0.1))
        if wl is None:
            kernels.append(O.Metropolis(ens_o, ush, T, seed=int(seeds[w]), walker=w))
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


@pytest.mark.parametrize("factorize", ["1", "gather", "spec", "spec-wide", "0", "random"])
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
        kw = dict(spec_mode=2, ewald_field=True)
        monkeypatch.setenv("LMC_SPEC_WIDE", "1" if factorize == "spec-wide" else "0")
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
    assert (smp._ew_field is not None) == (factorize in ("1", "spec", "spec-wide"))
    # a second run continues the chains (potential cache rebuilt from the occupancies, then kept current)
    smp.run(100, thin_by=10)
    assert smp.samples.num_samples == 40


@pytest.mark.parametrize("mode", ["classic", "spec"])
def test_canonical_ewald_swap_trajectory(cuda_device, mode):
    """canonical swaps with an Ewald term through the potential cache: flip 2 of a swap sees the cache
    shifted by flip 1 (one element of the site kernel); classic and speculative kernels"""
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
                            usher_kwargs=dict(spec_mode=2 if mode == "spec" else 1))
    _compare_traces(smp, ref)
    assert smp._ew_field is not None and 0 < smp.samples.step_efficiency() < 1


def test_wang_landau_flip_trajectory(cuda_device):
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


@pytest.mark.parametrize("group,factorize", [(8, "1"), (32, "1"), (32, "gather"), (32, "0")])
def test_table_flip_ewald_semigrand_trajectory(cuda_device, group, factorize, monkeypatch):
    """factorize=1: potential cache (flips of one step chained through elements of the site kernel);
    gather: one row of K per flip; 0: generic matrix rows"""
    monkeypatch.setenv("LMC_EWALD_FACTORIZE", "0" if factorize == "0" else "1")
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
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.05, ewald_matrix=ewm, ewald_inds=ewi))
    mus = {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0}
    ens_g = S.Ensemble(comp, chemical_potentials=mus)
    ora_p = O.CompositeProcessor([O.ClusterDecompositionProcessor(sub, scm, it),
                                  O.EwaldProcessor(ewm, ewi, 0.05)])

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
    smp, ref, _ = _run_both(ens_g, ens_o, "table_flip", W, 400, 20, occ0, seeds, T=2000.0,
                            usher_kwargs=dict(flip_table=table, swap_weight=0.2, ewald_field=factorize != "gather"
                                              if factorize != "0" else "auto"), group_size=group)
    _compare_traces(smp, ref)
    # charge neutrality is conserved by construction of the table
    occ = smp.samples.get_occupancies(flat=True)
    q_cat = np.array([1, 3, 4])[occ[:, :ncell]].sum(axis=1)
    q_ani = np.array([-2, -1])[occ[:, ncell:]].sum(axis=1)
    assert np.all(q_cat + q_ani == 0)
    assert smp.samples.step_efficiency() > 0


# ---------------------------------------------------------------------------------------------
# speculative-batch kernel (csrc/lmc_spec.cuh): same chain as the classic kernel and the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["decomposition", "expansion"])
@pytest.mark.parametrize("n,T,thin", [(2, 1000.0, 20), (4, 300.0, 7), (4, 1000.0, 64), (3, 1e5, 13)])
@pytest.mark.parametrize("sg", ["4", "2", "1"])
def test_speculative_swap_trajectory(cuda_device, kind, n, T, thin, sg, monkeypatch):
    """low / medium / near-infinite temperature (acceptance ~0 .. ~1), sampling intervals that are not
    multiples of the batch, aliased 2x2x2 cell; spec_mode=2 forces the speculative kernel, 1 the classic
    one: both must reproduce the oracle chain bit for bit.  sg = lanes per speculated step."""
    import smol_b200 as S
    monkeypatch.setenv("LMC_SPEC_SG", sg)
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
    smp, ref, _ = _run_both(ens_g, ens_o, "swap", W, nsteps, thin, occ0, seeds, T=T,
                            usher_kwargs=dict(spec_mode=2))
    _compare_traces(smp, ref)
    smp1 = S.Sampler.from_ensemble(ens_g, T, step_type="swap", nwalkers=W, seeds=list(seeds), spec_mode=1)
    smp1.run(nsteps, occ0, thin_by=thin)
    np.testing.assert_array_equal(smp1.samples.get_occupancies(flat=False),
                                  smp.samples.get_occupancies(flat=False))
    np.testing.assert_allclose(smp1.samples.get_enthalpies(flat=False), smp.samples.get_enthalpies(flat=False),
                               rtol=1e-12, atol=1e-12 * np.abs(ref["enthalpy"]).max())


@pytest.mark.parametrize("T", [400.0, 3000.0])
@pytest.mark.parametrize("sg", ["4", "2", "1"])
def test_speculative_semigrand_flip_trajectory(cuda_device, T, sg, monkeypatch):
    """ternary rocksalt cations, chemical potentials, single flips (no Ewald term): speculative kernel"""
    import smol_b200 as S
    monkeypatch.setenv("LMC_SPEC_SG", sg)
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
    smp, ref, _ = _run_both(ens_g, ens_o, "flip", W, 330, 11, occ0, seeds, T=T, usher_kwargs=dict(spec_mode=2))
    _compare_traces(smp, ref)


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
