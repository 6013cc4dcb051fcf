"""Multicell sampling (MulticellMetropolis, smol/moca/kernel/base.py:439-722): oracle restatement pinned against the
plain oracle chain, host schedule against numpy's ``Generator.choice``, GPU path against the oracle."""
import numpy as np
import pytest

from tests import models as M

RTOL = 1e-10

SHAPES = [np.diag([2, 2, 2]), np.diag([4, 2, 1]), np.array([[2, 1, 0], [0, 2, 0], [0, 0, 2]])]


def _oracle():
    from oracle import lmc_oracle as O
    return O


def _oracle_chain(O, sub, coefs, shapes, product_ensembles, T, step, seed, kseeds, walker, **kw):
    from smol_b200 import lattice as L
    it = L.cluster_interaction_tensors(sub, coefs)
    kernels = []
    for k, scm in enumerate(shapes):
        ens = O.Ensemble(O.ClusterDecompositionProcessor(sub, scm, it),
                         M.oracle_sublattices(O, product_ensembles[k].sublattices))
        ush = (O.Swap if step == "swap" else O.Flip)(ens.sublattices)
        kernels.append(O.Metropolis(ens, ush, T, seed=int(kseeds[k]), walker=walker))
    return O.MulticellMetropolis(kernels, T, seed=seed, **kw)


def _product_ensembles(sub, coefs, shapes):
    import smol_b200 as S
    from smol_b200 import lattice as L
    it = L.cluster_interaction_tensors(sub, coefs)
    return [S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it)) for scm in shapes]


def test_choice_stream_is_numpy_choice():
    """the pre-drawn schedule consumes the generator exactly like the reference's rng.choice(..., p=...) calls"""
    from smol_b200.multicell import _ChoiceStream, _cdf
    periods, hp = np.array([3, 5, 9]), np.array([0.2, 0.5, 0.3])
    kp = np.array([0.1, 0.6, 0.3])
    rng = np.random.default_rng(77)
    st = _ChoiceStream(77)
    for i in range(600):
        if i % 2 == 0:
            assert periods[np.searchsorted(_cdf(hp), st.uniform(), side="right")] == rng.choice(periods, p=hp)
        else:
            assert np.searchsorted(_cdf(kp), st.uniform(), side="right") == rng.choice(3, p=kp)


def test_oracle_single_shape_is_plain_metropolis():
    """one shape: a hop is one more step of the same chain judged through full feature vectors, so the multicell
    chain must be the plain Metropolis chain with the same Philox key (counter = step index)"""
    O = _oracle()
    sub = M.fcc_subspace()
    coefs = M.fcc_coefs(sub)
    shapes = [SHAPES[0]]
    pens = _product_ensembles(sub, coefs, shapes)     # host-only objects: their sublattices feed the oracle
    occ0 = M.random_occupancies(sub, shapes[0], 1, seed=3, balanced=True)
    chain = _oracle_chain(O, sub, coefs, shapes, pens, 2000.0, "swap", 5, [123], 0, kernel_hop_periods=3)
    got = O.run_multicell([chain], occ0[:, None, :], 240, 8)
    plain = chain._kernels[0].__class__(chain._kernels[0].ensemble, chain._kernels[0].usher, 2000.0, seed=123, walker=0)
    ref = O.run_sampler([plain], occ0, 240, 8)
    np.testing.assert_array_equal(got["occupancy"], ref["occupancy"])
    np.testing.assert_array_equal(got["accepted"], ref["accepted"])
    np.testing.assert_allclose(got["enthalpy"], ref["enthalpy"], rtol=1e-12, atol=1e-12)
    assert 0 < got["n_accepted"].sum() < 240


def test_oracle_multicell_tracks_full_features():
    O = _oracle()
    sub = M.fcc_subspace()
    coefs = M.fcc_coefs(sub)
    pens = _product_ensembles(sub, coefs, SHAPES)
    occ0 = np.stack([M.random_occupancies(sub, scm, 1, seed=10 + k, balanced=True)[0] for k, scm in enumerate(SHAPES)])
    chain = _oracle_chain(O, sub, coefs, SHAPES, pens, 3000.0, "swap", 9, [1, 2, 3], 0,
                          kernel_hop_periods=[2, 4], kernel_hop_probabilities=[0.5, 0.5],
                          kernel_probabilities=[0.25, 0.25, 0.5])
    out = O.run_multicell([chain], occ0[None], 300, 5)
    assert len(np.unique(out["kernel_index"])) == 3, "the chain never visited every shape; weak test"
    cur = chain._current_kernel_index     # (features of the other shapes are as of their last visit)
    full = chain._kernels[cur].ensemble.compute_feature_vector(chain.current_occupancy)
    np.testing.assert_allclose(chain._features[cur], full, rtol=1e-11, atol=1e-9)
    # the sampled state is always the current shape's
    np.testing.assert_array_equal(out["occupancy"][-1, 0], chain.current_occupancy)


# share=True = the reference's aliasing semantics: torch row copies and lmc_full_features calls in front of the same
# masked launches; both modes green on a B200 (GPUTEST_r01.json)
@pytest.mark.gpu
@pytest.mark.parametrize("share", [False, True], ids=["own-rows", "shared-live-row"])
@pytest.mark.parametrize("step", ["swap", "flip"])
def test_multicell_trajectory_vs_oracle(cuda_device, step, share):
    from smol_b200.multicell import MulticellSampler
    O = _oracle()
    sub = M.fcc_subspace()
    coefs = M.fcc_coefs(sub)
    pens = _product_ensembles(sub, coefs, SHAPES)
    W, K = 5, len(SHAPES)
    occ0 = np.stack([np.stack([M.random_occupancies(sub, scm, 1, seed=100 * w + k, balanced=True)[0]
                               for k, scm in enumerate(SHAPES)]) for w in range(W)])
    seeds = [11 + w for w in range(W)]
    kseeds = np.array([[1000 * (k + 1) + w for w in range(W)] for k in range(K)], dtype=np.uint64)
    kw = dict(kernel_hop_periods=[3, 5], kernel_hop_probabilities=[0.5, 0.5], kernel_probabilities=[0.25, 0.25, 0.5])
    T = 3000.0
    smp = MulticellSampler(pens, T, step_type=step, nwalkers=W, seeds=seeds, kernel_seeds=kseeds, share_visited=share, **kw)
    smp.run(240, occ0, thin_by=6)
    smp.run(120, thin_by=4)            # resumes: schedule, current shapes and states carry over
    chains = [_oracle_chain(O, sub, coefs, SHAPES, pens, T, step, seeds[w], kseeds[:, w], w, share_visited=share, **kw)
              for w in range(W)]
    ref1 = O.run_multicell(chains, occ0, 240, 6)
    ref2 = {k: [] for k in ref1}
    S2 = 120 // 4
    for s in range(S2):                 # continue the same chains
        for i, c in enumerate(chains):
            nacc = 0
            for _ in range(4):
                acc, _ = c.single_step()
                nacc += bool(acc)
            cur = c._current_kernel_index
            ref2["occupancy"].append(c.current_occupancy.copy()); ref2["kernel_index"].append(cur)
            ref2["accepted"].append(acc); ref2["n_accepted"].append(nacc)
            ref2["enthalpy"].append(float(np.dot(c.natural_params, c._features[cur])))
    s = smp.samples
    ki = s.get_trace_value("kernel_index", flat=False)
    assert ki.shape == (40 + S2, W, 1)
    np.testing.assert_array_equal(ki[:40], ref1["kernel_index"])
    np.testing.assert_array_equal(s.get_occupancies(flat=False)[:40], ref1["occupancy"])
    np.testing.assert_array_equal(s.get_trace_value("accepted", flat=False)[:40], ref1["accepted"])
    np.testing.assert_array_equal(s.get_trace_value("n_accepted", flat=False)[:40], ref1["n_accepted"])
    scale = np.abs(ref1["enthalpy"]).max()
    np.testing.assert_allclose(s.get_enthalpies(flat=False)[:40], ref1["enthalpy"], rtol=RTOL, atol=RTOL * scale)
    np.testing.assert_allclose(s.get_feature_vectors(flat=False)[:40], ref1["features"], rtol=RTOL,
                               atol=RTOL * np.abs(ref1["features"]).max())
    np.testing.assert_array_equal(ki[40:, :, 0], np.array(ref2["kernel_index"]).reshape(S2, W))
    np.testing.assert_array_equal(s.get_occupancies(flat=False)[40:], np.array(ref2["occupancy"]).reshape(S2, W, -1))
    np.testing.assert_array_equal(s.get_trace_value("n_accepted", flat=False)[40:],
                                  np.array(ref2["n_accepted"]).reshape(S2, W))
    np.testing.assert_allclose(s.get_enthalpies(flat=False)[40:, :, 0], np.array(ref2["enthalpy"]).reshape(S2, W),
                               rtol=RTOL, atol=RTOL * scale)
    assert len(np.unique(ki)) == K and 0 < s.get_trace_value("n_accepted", flat=False).sum()


# ---- host logic of MulticellSampler without a GPU: an engine stand-in whose run() is the oracle ------------------
class _OracleEngine:
    """Stands in for LmcEngine on CPU tensors: `run(cfg)` advances the walkers selected by walker_mask_dev with the
    oracle's Metropolis step (accept_offset_dev inside the exponent), reading / writing the state through the raw
    pointers of the run configuration exactly as the CUDA library would."""
    queue = []          # (oracle ensemble factory, usher class) per constructed engine

    def __init__(self, packed, device=None):
        import torch
        self.ens_factory, self.usher_cls, self.T_of = _OracleEngine.queue.pop(0)
        ens = self.ens_factory()
        self.N, self.F = int(packed.desc.num_sites), int(packed.desc.num_features)
        assert self.F == len(ens.natural_parameters)
        self.row_stride = (self.N + 16) // 16 * 16
        self.device = torch.device("cpu")

    def upload_occupancy(self, occ):
        import torch
        out = torch.zeros((occ.shape[0], self.row_stride), dtype=torch.int8)
        out[:, :self.N] = torch.from_numpy(np.asarray(occ).astype(np.int8))
        return out

    def occupancy_to_int32(self, occ_dev, rows, stride):
        return occ_dev[:, :self.N].to(dtype=__import__("torch").int32)

    def full_features(self, occ_dev):
        import torch
        ens = self.ens_factory()
        f = np.array([ens.compute_feature_vector(o.numpy()[:self.N].astype(np.int32)) for o in occ_dev])
        return torch.from_numpy(f), torch.from_numpy(f @ np.asarray(ens.natural_parameters))

    def distance_tables(self, processor):
        import torch
        z = torch.zeros(1, dtype=torch.float64)
        return dict(target=z, tol=0.0, goff=z, gidx=z, gdiam=z, ngrp=0)

    def distance_init(self, processor, feat, enth=None):
        """the oracle ensemble of a distance processor already returns distance vectors from full_features"""
        import torch
        return torch.zeros_like(feat)

    def run(self, cfg):
        import ctypes as C
        import math
        O = _oracle()
        W = cfg.num_walkers

        def arr(ptr, ctype, n):
            return np.ctypeslib.as_array((ctype * n).from_address(ptr))
        mask = arr(cfg.walker_mask_dev, C.c_uint8, W)
        off = arr(cfg.accept_offset_dev, C.c_double, W) if cfg.accept_offset_dev else np.zeros(W)
        occ = arr(cfg.occ_dev, C.c_int8, W * self.row_stride).reshape(W, self.row_stride)
        feat = arr(cfg.features_dev, C.c_double, W * self.F).reshape(W, self.F)
        enth = arr(cfg.enthalpy_dev, C.c_double, W)
        seeds = arr(cfg.seeds_dev, C.c_uint64, W)
        beta = arr(cfg.beta_dev, C.c_double, W)
        acc_o = arr(cfg.trace_accepted_dev, C.c_uint8, W)
        nacc_o = arr(cfg.trace_naccepted_dev, C.c_int32, W)
        assert cfg.num_samples == 1 and cfg.spec_mode == 1
        for w in range(W):
            if not mask[w]:
                continue
            ens = self.ens_factory()
            ush = self.usher_cls(ens.sublattices)
            o = occ[w, :self.N].astype(np.int32)
            nacc, accepted = 0, True
            for i in range(cfg.thin_by):
                rnd = O.StepRandom(int(seeds[w]), cfg.walker_id_base + w, cfg.step_begin + i)
                step = ush.propose_step(o, rnd)
                dfeat = np.array(ens.compute_feature_vector_change(o, step), dtype=np.float64)
                dH = float(np.dot(ens.natural_parameters, dfeat))
                exponent = -beta[w] * (dH + off[w]) + ush.compute_log_priori_factor(o, step)
                accepted = True if exponent >= 0 else exponent > math.log(O.u01(rnd.word(3)))
                if accepted:
                    for site, sp in step:
                        o[site] = sp
                    feat[w] += dfeat
                    enth[w] += dH
                    nacc += 1
            occ[w, :self.N] = o.astype(np.int8)
            acc_o[w], nacc_o[w] = int(accepted), nacc


@pytest.mark.parametrize("share", [True, False], ids=["shared-live-row", "own-rows"])
@pytest.mark.parametrize("step", ["swap", "flip"])
def test_multicell_host_logic_with_oracle_engine(monkeypatch, step, share):
    """schedule, masks, offsets, current-shape bookkeeping and traces of MulticellSampler (host side) against the
    oracle's multicell chain, with the CUDA library replaced by an oracle-backed stand-in"""
    import smol_b200.engine as E
    from smol_b200 import lattice as L
    from smol_b200.multicell import MulticellSampler
    O = _oracle()
    sub = M.fcc_subspace()
    coefs = M.fcc_coefs(sub)
    it = L.cluster_interaction_tensors(sub, coefs)
    pens = _product_ensembles(sub, coefs, SHAPES)
    ush = O.Swap if step == "swap" else O.Flip
    _OracleEngine.queue = [
        ((lambda scm=scm, k=k: O.Ensemble(O.ClusterDecompositionProcessor(sub, scm, it),
                                          M.oracle_sublattices(O, pens[k].sublattices))), ush, None)
        for k, scm in enumerate(SHAPES)]
    monkeypatch.setattr(E, "LmcEngine", _OracleEngine)
    W, K = 3, len(SHAPES)
    occ0 = np.stack([np.stack([M.random_occupancies(sub, scm, 1, seed=100 * w + k, balanced=True)[0]
                               for k, scm in enumerate(SHAPES)]) for w in range(W)])
    seeds = [11 + w for w in range(W)]
    kseeds = np.array([[1000 * (k + 1) + w for w in range(W)] for k in range(K)], dtype=np.uint64)
    kw = dict(kernel_hop_periods=[3, 5], kernel_hop_probabilities=[0.5, 0.5], kernel_probabilities=[0.25, 0.25, 0.5])
    T = 3000.0
    smp = MulticellSampler(pens, T, step_type=step, nwalkers=W, seeds=seeds, kernel_seeds=kseeds, share_visited=share, **kw)
    smp.run(120, occ0, thin_by=6)
    chains = [_oracle_chain(O, sub, coefs, SHAPES, pens, T, step, seeds[w], kseeds[:, w], w, share_visited=share, **kw)
              for w in range(W)]
    ref = O.run_multicell(chains, occ0, 120, 6)
    s = smp.samples
    np.testing.assert_array_equal(s.get_trace_value("kernel_index", flat=False), ref["kernel_index"])
    np.testing.assert_array_equal(s.get_occupancies(flat=False), ref["occupancy"])
    np.testing.assert_array_equal(s.get_trace_value("accepted", flat=False), ref["accepted"])
    np.testing.assert_array_equal(s.get_trace_value("n_accepted", flat=False), ref["n_accepted"])
    np.testing.assert_allclose(s.get_enthalpies(flat=False), ref["enthalpy"], rtol=RTOL,
                               atol=RTOL * np.abs(ref["enthalpy"]).max())
    np.testing.assert_allclose(s.get_feature_vectors(flat=False), ref["features"], rtol=RTOL,
                               atol=RTOL * np.abs(ref["features"]).max())
    assert len(np.unique(ref["kernel_index"])) > 1
    np.testing.assert_array_equal(smp.current_occupancies()[np.arange(W), smp.current_kernel_indices()],
                                  ref["occupancy"][-1])


def _sqs_ensembles():
    """product + oracle ensembles of the SQS kind: CorrelationDistanceProcessors of the three shapes, random-alloy target"""
    import smol_b200 as S
    O = _oracle()
    sub = M.fcc_subspace()
    target = np.zeros(sub.num_corr_functions)
    target[0] = 1.0
    kw = dict(target_vector=target, match_weight=0.05, match_tol=1e-5)
    pens = [S.Ensemble(S.CorrelationDistanceProcessor(sub, scm, **kw)) for scm in SHAPES]
    oens = [(lambda scm=scm, k=k: O.Ensemble(O.CorrelationDistanceProcessor(sub, scm, **kw),
                                             M.oracle_sublattices(O, pens[k].sublattices))) for k, scm in enumerate(SHAPES)]
    return sub, pens, oens


def _sqs_reference(O, oens, T, seeds, kseeds, W, kw, occ0, nsteps, thin):
    chains = []
    for w in range(W):
        kernels = []
        for k, mk in enumerate(oens):
            ens = mk()
            kernels.append(O.Metropolis(ens, O.Swap(ens.sublattices), T, seed=int(kseeds[k, w]), walker=w))
        chains.append(O.MulticellMetropolis(kernels, T, seed=seeds[w], **kw))
    return O.run_multicell(chains, occ0, nsteps, thin)


def test_multicell_over_distance_processors_host_logic(monkeypatch):
    """the SQS combination (multicell hops between distance-processor ensembles) through MulticellSampler's host logic,
    CUDA library replaced by the oracle-backed stand-in, vs the oracle chain (itself pinned against the reference's
    classes for this very combination, tests/test_oracle_golden.py)"""
    import smol_b200.engine as E
    from smol_b200.multicell import MulticellSampler
    O = _oracle()
    sub, pens, oens = _sqs_ensembles()
    _OracleEngine.queue = [(mk, O.Swap, None) for mk in oens]
    monkeypatch.setattr(E, "LmcEngine", _OracleEngine)
    W, K, T = 3, len(SHAPES), 3000.0
    occ0 = np.stack([np.stack([M.random_occupancies(sub, scm, 1, seed=100 * w + k, balanced=True)[0]
                               for k, scm in enumerate(SHAPES)]) for w in range(W)])
    seeds = [21 + w for w in range(W)]
    kseeds = np.array([[4000 * (k + 1) + w for w in range(W)] for k in range(K)], dtype=np.uint64)
    kw = dict(kernel_hop_periods=[2, 4], kernel_hop_probabilities=[0.5, 0.5])
    smp = MulticellSampler(pens, T, step_type="swap", nwalkers=W, seeds=seeds, kernel_seeds=kseeds, **kw)
    smp.run(160, occ0, thin_by=4)
    ref = _sqs_reference(O, oens, T, seeds, kseeds, W, kw, occ0, 160, 4)
    s = smp.samples
    np.testing.assert_array_equal(s.get_trace_value("kernel_index", flat=False), ref["kernel_index"])
    np.testing.assert_array_equal(s.get_occupancies(flat=False), ref["occupancy"])
    np.testing.assert_array_equal(s.get_trace_value("accepted", flat=False), ref["accepted"])
    np.testing.assert_allclose(s.get_feature_vectors(flat=False), ref["features"], rtol=RTOL, atol=RTOL)
    np.testing.assert_allclose(s.get_enthalpies(flat=False), ref["enthalpy"], rtol=RTOL, atol=RTOL)
    assert len(np.unique(ref["kernel_index"])) > 1 and ref["n_accepted"].sum() > 10


@pytest.mark.gpu
def test_multicell_over_distance_processors_vs_oracle(cuda_device):
    """the same on the CUDA path (DIST kernel variants with walker masks and hop offsets)"""
    from smol_b200.multicell import MulticellSampler
    O = _oracle()
    sub, pens, oens = _sqs_ensembles()
    W, K, T = 4, len(SHAPES), 3000.0
    occ0 = np.stack([np.stack([M.random_occupancies(sub, scm, 1, seed=100 * w + k, balanced=True)[0]
                               for k, scm in enumerate(SHAPES)]) for w in range(W)])
    seeds = [21 + w for w in range(W)]
    kseeds = np.array([[4000 * (k + 1) + w for w in range(W)] for k in range(K)], dtype=np.uint64)
    kw = dict(kernel_hop_periods=[2, 4], kernel_hop_probabilities=[0.5, 0.5])
    smp = MulticellSampler(pens, T, step_type="swap", nwalkers=W, seeds=seeds, kernel_seeds=kseeds, **kw)
    smp.run(160, occ0, thin_by=4)
    ref = _sqs_reference(O, oens, T, seeds, kseeds, W, kw, occ0, 160, 4)
    s = smp.samples
    np.testing.assert_array_equal(s.get_trace_value("kernel_index", flat=False), ref["kernel_index"])
    np.testing.assert_array_equal(s.get_occupancies(flat=False), ref["occupancy"])
    np.testing.assert_array_equal(s.get_trace_value("accepted", flat=False), ref["accepted"])
    np.testing.assert_allclose(s.get_feature_vectors(flat=False), ref["features"], rtol=RTOL, atol=RTOL)
    np.testing.assert_allclose(s.get_enthalpies(flat=False), ref["enthalpy"], rtol=RTOL, atol=RTOL)
