"""Host-side check of the tables of the speculative-batch kernel (csrc/lmc_api.cu:build_spec_tables):
the merged three-gather records and the pre-differenced table must reproduce the oracle's energy change
of every single flip.  Runs without a GPU (``lmc_spec_tables_host`` makes no CUDA call)."""
import ctypes as C

import numpy as np
import pytest

from smol_b200 import _capi as capi
from smol_b200 import lattice as L
from tests import models as M


def _oracle():
    from oracle import lmc_oracle as O
    return O


def _tables(packed):
    lib = capi.load()
    info = (C.c_int32 * 8)()
    capi.check(lib.lmc_spec_tables_host(C.byref(packed.desc), info, None, 0, None, 0))
    info = list(info)
    if not info[0]:
        return info, None, None
    dtab = np.zeros(info[6] // 8)
    rec = np.zeros(info[7], dtype=np.uint8)
    capi.check(lib.lmc_spec_tables_host(C.byref(packed.desc), (C.c_int32 * 8)(), dtab.ctypes.data_as(C.c_void_p),
                                        dtab.size, rec.ctypes.data_as(C.c_void_p), rec.size))
    NC, Lp, NQ = info[1], info[2], info[3]
    return info, dtab.reshape(NC, Lp), rec.view(np.uint32).reshape(packed.desc.num_sites, NQ, 2)


def _spec_delta(dtab, rec, NC, occ, site, new, patch=None):
    """energy change of flipping `site` to `new` as the kernel computes it (spec_rec2)"""
    N = len(occ)
    row = np.append(occ, 0)                 # zero pad byte gathered by unused slots
    if patch is not None:
        row = row.copy(); row[patch[0]] = patch[1]
    old = occ[site]
    x, y = rec[site, :, 0], rec[site, :, 1]
    s0, s1, s2, tb = x & 0xffff, x >> 16, y & 0xffff, y >> 16
    assert max(s0.max(), s1.max(), s2.max()) <= N
    idx = tb + old + NC * (row[s0] + NC * (row[s1] + NC * row[s2]))
    return dtab[new][idx].sum()


CASES = [
    ("fcc2", lambda: (M.fcc_subspace(), 2, "decomposition")),
    ("fcc4_corr", lambda: (M.fcc_subspace(), 4, "expansion")),
    ("rocksalt3", lambda: (M.rocksalt_subspace(), 3, "decomposition")),
    ("rocksalt_two_sublattices", lambda: (M.rocksalt_subspace(anions=("O2-", "F-")), 3, "expansion")),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("merge", ["2", "1", "0"])
def test_spec_tables_reproduce_oracle_flip_energies(name, make, merge, monkeypatch):
    import smol_b200 as S
    monkeypatch.setenv("LMC_SPEC_MERGE", merge)
    O = _oracle()
    sub, n, kind = make()
    scm = np.eye(3, dtype=int) * n
    rng = np.random.default_rng(3)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    if kind == "decomposition":
        it = L.cluster_interaction_tensors(sub, coefs)
        proc = S.ClusterDecompositionProcessor(sub, scm, it)
        ora = O.ClusterDecompositionProcessor(sub, scm, it)
    else:
        proc = S.ClusterExpansionProcessor(sub, scm, coefs)
        ora = O.ClusterExpansionProcessor(sub, scm, coefs)
    ens = S.Ensemble(proc)
    packed = ens.packed_model()
    info, dtab, rec = _tables(packed)
    assert info[0] == 1, info
    assert info[5] <= int(merge)
    NC = info[1]
    occ = M.random_occupancies(sub, scm, 1, seed=5)[0]
    spaces = sub.allowed_species(scm)
    nat = np.asarray(ens.natural_parameters)
    scale = 0.0
    checked = 0
    for site in rng.permutation(len(occ))[:40]:
        ns = len(spaces[site])
        if ns < 2:
            continue
        for new in range(ns):
            if new == occ[site]:
                continue
            ref = float(np.dot(nat, ens_change(ora, occ, [(int(site), int(new))])))
            got = _spec_delta(dtab, rec, NC, occ, int(site), new)
            scale = max(scale, abs(ref))
            assert abs(got - ref) <= 1e-10 * max(1.0, abs(ref)), (site, new, got, ref)
            checked += 1
    assert checked >= 8 and scale > 0
    # a swap: the second flip sees the first applied (PATCH of the gathers)
    act = [i for i in range(len(occ)) if len(spaces[i]) > 1]
    for _ in range(20):
        a, b = rng.choice(act, 2, replace=False)
        if occ[a] == occ[b] or len(spaces[a]) != len(spaces[b]):
            continue
        ref = float(np.dot(nat, ens_change(ora, occ, [(int(a), int(occ[b])), (int(b), int(occ[a]))])))
        got = _spec_delta(dtab, rec, NC, occ, int(a), int(occ[b])) + \
            _spec_delta(dtab, rec, NC, occ, int(b), int(occ[a]), patch=(int(a), int(occ[b])))
        assert abs(got - ref) <= 1e-10 * max(1.0, abs(ref))


def ens_change(ora, occ, flips):
    return ora.compute_feature_vector_change(occ, flips)


def test_cover_merge_quarters_the_lookups():
    import smol_b200 as S
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 8
    it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub))
    packed = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it)).packed_model()
    info, dtab, rec = _tables(packed)
    assert info[0] == 1 and info[5] == 2
    # 87 local clusters of config 2 -> 22 records: 8 nearest-neighbour triangles of sites (each carries a
    # tetrahedron, its three triangles and nearest-neighbour pairs) + 14 records of three farther pairs
    assert info[3] == 24
    assert info[6] <= 8 * 1024    # the difference table stays a few KB of shared memory
