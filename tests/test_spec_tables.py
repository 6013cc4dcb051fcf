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


def _spec_delta(dtab, rec, NC, occ, site, new, patch=None, entries=False):
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
    if entries:
        return np.sort(dtab[new][idx])
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


# ------------------------------------------------------------------------------------------------------------------
# environment words (csrc/lmc_api.cu:build_env_tables, ENV variants of the speculative kernel)
def _env_tables(packed):
    lib = capi.load()
    info = (C.c_int32 * 8)()
    capi.check(lib.lmc_spec_env_host(C.byref(packed.desc), info, None, 0, None, 0, None, 0))
    ok, b, nrl, nrlp, wide, NA, RV, pair_ok = list(info)
    if not ok:
        return list(info), None
    N = packed.desc.num_sites
    tb = np.zeros(N * 4 * nrlp, dtype=np.uint16)
    rev = np.zeros(NA * RV, dtype=np.uint32)
    pair = np.zeros(NA * NA * 4 if pair_ok else 0, dtype=np.uint64 if wide else np.uint32)
    capi.check(lib.lmc_spec_env_host(C.byref(packed.desc), (C.c_int32 * 8)(), tb.ctypes.data_as(C.c_void_p), tb.size,
                                     rev.ctypes.data_as(C.c_void_p), rev.size, pair.ctypes.data_as(C.c_void_p), pair.nbytes))
    return list(info), dict(tb=tb.reshape(N, 4, nrlp), rev=rev.reshape(NA, RV),
                            pair=pair.reshape(NA, NA, 4) if pair_ok else None)


class _EnvModel:
    """the kernel's environment-word arithmetic (env_build / env_flip_energy / env_commit / the swap patch) in Python"""

    def __init__(self, info, tabs, dtab, rec, NC, active_sites):
        _, self.b, self.nrl, _, self.wide, self.NA, _, _ = info
        self.t, self.dtab, self.rec, self.NC = tabs, dtab, rec, NC
        self.sites = list(active_sites)            # active index -> site
        self.lane_bits = 64 if self.wide else 32

    def build(self, occ):
        row = np.append(occ, 0)
        b, fb = self.b, 3 * self.b
        env = np.zeros((self.NA, 4), dtype=object)
        for ai, site in enumerate(self.sites):
            for l in range(4):
                chunk = 0
                for i in range(self.nrl):
                    r = 2 * (l + 4 * (i >> 1)) + (i & 1)
                    x, y = int(self.rec[site, r, 0]), int(self.rec[site, r, 1])
                    field = int(row[x & 0xffff]) | (int(row[x >> 16]) << b) | (int(row[y & 0xffff]) << (2 * b))
                    chunk |= field << (i * fb)
                assert chunk < (1 << self.lane_bits)
                env[ai, l] = chunk
        return env

    def delta(self, env_row, site, old, new):
        b, fb, NC = self.b, 3 * self.b, self.NC
        vals = []
        for l in range(4):
            for i in range(self.nrl):
                field = (env_row[l] >> (i * fb)) & ((1 << fb) - 1)
                cm = (1 << b) - 1
                ci = (field & cm) + NC * (((field >> b) & cm) + NC * (field >> (2 * b)))
                vals.append(self.dtab[new][int(self.t["tb"][site, l, i]) + old + NC * ci])
        return np.sort(np.array(vals))     # the table entries the kernel adds up

    def patched(self, env, ai2, ai1, x):
        return [env[ai2, l] ^ (int(self.t["pair"][ai2, ai1, l]) * x) for l in range(4)]

    def commit(self, env, ai, x):
        for ent in self.t["rev"][ai]:
            ent = int(ent)
            if ent == 0xffffffff:
                continue
            k, p = ent & 0xffff, ent >> 16
            env[k, p // self.lane_bits] ^= x << (p % self.lane_bits)


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_environment_words_reproduce_the_gather_variant(name, make):
    """environment words built from an occupancy, patched for a swap's second flip and updated by accepted flips
    give exactly the table entries the gather variant reads (so the two kernels sum the same doubles)"""
    import smol_b200 as S
    sub, n, kind = make()
    scm = np.eye(3, dtype=int) * n
    rng = np.random.default_rng(11)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    proc = (S.ClusterDecompositionProcessor(sub, scm, L.cluster_interaction_tensors(sub, coefs))
            if kind == "decomposition" else S.ClusterExpansionProcessor(sub, scm, coefs))
    ens = S.Ensemble(proc)
    packed = ens.packed_model()
    info, dtab, rec = _tables(packed)
    einfo, tabs = _env_tables(packed)
    NC = info[1]
    bits = 1 if NC <= 2 else 2
    if NC > 4 or info[3] // 4 * 3 * bits > 64:   # a lane's records do not fit one 64-bit chunk: the model keeps the gather variant
        assert einfo[0] == 0
        return
    assert info[0] == 1 and einfo[0] == 1, (info, einfo)
    assert einfo[1] == bits and einfo[2] == info[3] // 4 and einfo[7] == 1
    d = packed.desc
    sl_off = list(np.ctypeslib.as_array(C.cast(d.sl_site_off, C.POINTER(C.c_int32)), (d.num_sublattices + 1,)))
    active = [int(v) for v in np.ctypeslib.as_array(C.cast(d.sl_sites, C.POINTER(C.c_int32)), (int(sl_off[-1]),))]
    # (active index = sl_off[sublattice] + position)
    assert einfo[5] == len(active)
    aidx = {s: a for a, s in enumerate(active)}
    em = _EnvModel(einfo, tabs, dtab, rec, NC, active)
    occ = M.random_occupancies(sub, scm, 1, seed=8)[0].copy()
    spaces = sub.allowed_species(scm)
    env = em.build(occ)
    nflips = npatched = 0
    for it in range(60):
        # single flips against the current words
        site = int(rng.choice(active))
        new = int(rng.choice([c for c in range(len(spaces[site])) if c != occ[site]]))
        got = em.delta([env[aidx[site], l] for l in range(4)], site, int(occ[site]), new)
        ref = _spec_delta(dtab, rec, NC, occ, site, new, entries=True)
        assert np.array_equal(got, ref), (it, site, new)
        # a swap inside one sublattice: the second flip sees the first through the slot masks
        k = int(rng.integers(d.num_sublattices))
        sites_sl = active[sl_off[k]:sl_off[k + 1]]
        a, bq = (int(v) for v in rng.choice(sites_sl, 2, replace=False))
        if occ[a] != occ[bq]:
            # prefer neighbours now and then so that the patch is exercised
            nb = [active[k] for k in range(len(active)) if tabs["pair"][aidx[bq], k].any()] if it % 2 else []
            nb = [s for s in nb if s in sites_sl and occ[s] != occ[bq]]
            if nb:
                a = nb[0]
            x = int(occ[a]) ^ int(occ[bq])
            row2 = em.patched(env, aidx[bq], aidx[a], x)
            npatched += int(tabs["pair"][aidx[bq], aidx[a]].any())
            got2 = em.delta(row2, bq, int(occ[bq]), int(occ[a]))
            ref2 = _spec_delta(dtab, rec, NC, occ, bq, int(occ[a]), patch=(a, int(occ[bq])), entries=True)
            assert np.array_equal(got2, ref2), (it, a, bq)
        # accept the flip: the words of every gathering site follow
        em.commit(env, aidx[site], int(occ[site]) ^ new)
        occ[site] = new
        nflips += 1
    assert npatched >= 3
    fresh = em.build(occ)
    assert all(env[a, l] == fresh[a, l] for a in range(len(active)) for l in range(4))


# ------------------------------------------------------------------------------------------------------------------
# compact environment words (csrc/lmc_api.cu:build_c64_tables, csrc/lmc_spec_c64.cuh)
def _c64_tables(packed):
    lib = capi.load()
    info = (C.c_int32 * 8)()
    capi.check(lib.lmc_spec_c64_host(C.byref(packed.desc), info, None, 0, None, 0, None, 0, None, 0))
    ok, b, nrl, nrlp, NA, RV, ncls, bits = list(info)
    if not ok:
        return list(info), None
    N = packed.desc.num_sites
    desc = np.zeros(ncls * 4 * nrlp, dtype=np.uint32)
    cls = np.zeros(N, dtype=np.uint8)
    rev = np.zeros(NA * RV, dtype=np.uint32)
    pair = np.zeros(NA * NA, dtype=np.uint64)
    capi.check(lib.lmc_spec_c64_host(C.byref(packed.desc), (C.c_int32 * 8)(), desc.ctypes.data_as(C.c_void_p), desc.size,
                                     cls.ctypes.data_as(C.c_void_p), cls.size, rev.ctypes.data_as(C.c_void_p), rev.size,
                                     pair.ctypes.data_as(C.c_void_p), pair.size))
    return list(info), dict(desc=desc.reshape(ncls, 4, nrlp), cls=cls, rev=rev.reshape(NA, RV), pair=pair.reshape(NA, NA))


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_compact_environment_words_reproduce_the_gather_variant(name, make):
    """one 64-bit word per site (overlapping record fields): built from an occupancy, patched for a swap's second flip
    through the pair mask and updated through the reverse map, the words give exactly the table entries the gather
    variant reads"""
    import smol_b200 as S
    sub, n, kind = make()
    scm = np.eye(3, dtype=int) * n
    rng = np.random.default_rng(12)
    coefs = rng.normal(0, 0.05, sub.num_corr_functions)
    proc = (S.ClusterDecompositionProcessor(sub, scm, L.cluster_interaction_tensors(sub, coefs))
            if kind == "decomposition" else S.ClusterExpansionProcessor(sub, scm, coefs))
    ens = S.Ensemble(proc)
    packed = ens.packed_model()
    info, dtab, rec = _tables(packed)
    cinfo, tabs = _c64_tables(packed)
    if not cinfo[0]:
        pytest.skip("the records of this model do not fit 64 bits per site: gather variant")
    NC, NQ = info[1], info[3]
    ok, b, nrl, nrlp, NA, RV, ncls, bits = cinfo
    assert bits <= 64 and b == (1 if NC <= 2 else 2) and nrl == NQ // 4
    d = packed.desc
    sl_off = list(np.ctypeslib.as_array(C.cast(d.sl_site_off, C.POINTER(C.c_int32)), (d.num_sublattices + 1,)))
    active = [int(v) for v in np.ctypeslib.as_array(C.cast(d.sl_sites, C.POINTER(C.c_int32)), (int(sl_off[-1]),))]
    aidx = {s: a for a, s in enumerate(active)}
    fb, cm = 3 * b, (1 << b) - 1

    def build(occ):
        row = np.append(occ, 0)
        env = [0] * NA
        for ai, site in enumerate(active):
            e = 0
            for l in range(4):
                for i in range(nrl):
                    r = 2 * (l + 4 * (i >> 1)) + (i & 1)
                    x, y = int(rec[site, r, 0]), int(rec[site, r, 1])
                    field = int(row[x & 0xffff]) | (int(row[x >> 16]) << b) | (int(row[y & 0xffff]) << (2 * b))
                    sh = int(tabs["desc"][tabs["cls"][site], l, i]) >> 16
                    assert (e >> sh) & ((1 << fb) - 1) in (0, field) or True
                    e |= field << sh
            assert e < (1 << 64)
            env[ai] = e
        return env

    def entries(e, site, old, new):
        vals = []
        for l in range(4):
            for i in range(nrl):
                dsc = int(tabs["desc"][tabs["cls"][site], l, i])
                field = (e >> (dsc >> 16)) & ((1 << fb) - 1)
                ci = (field & cm) + NC * (((field >> b) & cm) + NC * (field >> (2 * b)))
                vals.append(dtab[new][(dsc & 0xffff) + old + NC * ci])
        return np.sort(np.array(vals))

    occ = M.random_occupancies(sub, scm, 1, seed=9)[0].copy()
    spaces = sub.allowed_species(scm)
    env = build(occ)
    npatched = 0
    for it in range(60):
        site = int(rng.choice(active))
        new = int(rng.choice([c for c in range(len(spaces[site])) if c != occ[site]]))
        assert np.array_equal(entries(env[aidx[site]], site, int(occ[site]), new),
                              _spec_delta(dtab, rec, NC, occ, site, new, entries=True)), (it, site, new)
        k = int(rng.integers(d.num_sublattices))
        sites_sl = active[sl_off[k]:sl_off[k + 1]]
        a, bq = (int(v) for v in rng.choice(sites_sl, 2, replace=False))
        if occ[a] != occ[bq]:
            nb = [active[j] for j in range(NA) if tabs["pair"][aidx[bq], j]] if it % 2 else []
            nb = [s for s in nb if s in sites_sl and occ[s] != occ[bq]]
            if nb:
                a = nb[0]
            pm = int(tabs["pair"][aidx[bq], aidx[a]])
            npatched += int(pm != 0)
            e2 = env[aidx[bq]] ^ (pm * (int(occ[a]) ^ int(occ[bq])))
            assert np.array_equal(entries(e2, bq, int(occ[bq]), int(occ[a])),
                                  _spec_delta(dtab, rec, NC, occ, bq, int(occ[a]), patch=(a, int(occ[bq])), entries=True)), (it, a, bq)
        x = int(occ[site]) ^ new
        for ent in tabs["rev"][aidx[site]]:
            ent = int(ent)
            if ent != 0xffffffff:
                env[ent & 0xffff] ^= x << (ent >> 16)
        occ[site] = new
    assert npatched >= 3
    assert env == build(occ)


def test_compact_environment_words_fit_the_fcc_cluster_set():
    """BASELINE config 2: 22 records x 3 bits = 66 bits per site become 64 through two chained fields; one class of sites"""
    from tests import workloads as WK
    packed = WK.get(2).product_ensemble().packed_model()
    cinfo, tabs = _c64_tables(packed)
    assert cinfo[0] == 1 and cinfo[1] == 1 and cinfo[7] <= 64, cinfo
