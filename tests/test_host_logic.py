"""Host-side logic that needs no GPU: C-ABI library exports, table packing, sampler arguments."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import lmc_oracle as O
from smol_b200 import _capi as capi
from smol_b200 import lattice as L
from smol_b200.container import SampleContainer
from smol_b200.dist import shard_walkers
from smol_b200.model import ExpansionTables
from smol_b200.sampler import table_flip_tables
from smol_b200.sublattice import Sublattice
from tests import models as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    """include/lmc.h is the boundary: every declared entry point must be exported (no compute calls)."""
    hdr = open(os.path.join(ROOT, "include", "lmc.h")).read()
    declared = set(re.findall(r"\b(lmc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    path = capi.lib_path()
    if not os.path.exists(path):
        from smol_b200 import build
        build.build()
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), name
    lib.lmc_version.restype = ctypes.c_int
    assert lib.lmc_version() == capi.LMC_ABI_VERSION
    assert lib.lmc_row_stride(512) == 528 and lib.lmc_row_stride(27) == 32   # >= 1 zero pad byte


def test_ctypes_struct_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "lmc.h")).read()
    body = hdr[hdr.index("typedef struct LmcModelDesc {"):hdr.index("} LmcModelDesc;")]
    fields = re.findall(r"\b([a-z_0-9]+);", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert fields == [f[0] for f in capi.LmcModelDesc._fields_]
    body = hdr[hdr.index("typedef struct LmcRunConfig {"):hdr.index("} LmcRunConfig;")]
    fields = re.findall(r"\b([a-z_0-9]+);", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert fields == [f[0] for f in capi.LmcRunConfig._fields_]


@pytest.mark.parametrize("n", [2, 4])
def test_packed_records_reproduce_reference_local_tables(n):
    """The record/class/self-stride packing must encode exactly the reference's per-site reduced
    index arrays (processor/expansion.py:124-138), including aliased rows on small cells."""
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * n
    coefs = M.fcc_coefs(sub)
    it = L.cluster_interaction_tensors(sub, coefs)
    e = ExpansionTables(sub, scm, "interaction", interaction_tensors=it)
    ora = O.ClusterDecompositionProcessor(sub, scm, it)
    rng = np.random.default_rng(0)
    occ = M.random_occupancies(sub, scm, 1, seed=1)[0]
    for site in rng.choice(len(occ), size=min(len(occ), 6), replace=False):
        new = 1 - occ[site]
        want = ora.compute_feature_vector_change(occ, [(site, new)])
        got = np.zeros(e.num_features)
        for r in range(e.site_rec_off[site], e.site_rec_off[site + 1]):
            i0, i1, i2, c = (int(x) for x in e.site_rec[r])
            st = e.cls_stride[c]
            orb = e.cls_orbit[c]
            base = st[0] * occ[i0] + st[1] * occ[i1] + st[2] * occ[i2]
            t = e.ftab[e.orb_tab_off[orb]:e.orb_tab_off[orb] + e.orb_tab_len[orb]]
            got[e.orb_fidx[orb]] += (t[base + new * st[3]] - t[base + occ[site] * st[3]]) * e.orb_weight[orb]
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-13)
        # segments partition the records by orbit, in order
        segs = e.site_seg[e.site_seg_off[site]:e.site_seg_off[site + 1]]
        assert segs[:, 1].sum() == e.site_rec_off[site + 1] - e.site_rec_off[site]
        assert np.all(np.diff(segs[:, 2]) > 0)


def test_table_flip_tables_and_errors():
    sl = [Sublattice(("Li+", "Mn3+", "Ti4+"), np.arange(4)), Sublattice(("O2-", "F-"), np.arange(4, 8))]
    t = table_flip_tables(sl, [[-1, 1, 0, 2, -2], [0, -1, 1, 1, -1]])
    assert t["num_dims"] == 5 and t["dim_sl"] == [0, 0, 0, 1, 1] and t["dim_code"] == [0, 1, 2, 0, 1]
    assert t["max_n"] == [4, 4, 4, 4, 4] and len(t["weights"]) == 4
    with pytest.raises(ValueError):
        table_flip_tables(sl, [[1, -1, 0]])
    with pytest.raises(ValueError):
        table_flip_tables(sl, [[-1, 1, 0, 2, -2]], flip_weights=[1, 2, 3])
    fixed = [Sublattice(("Li+", "Mn3+"), np.arange(4)), Sublattice(("O2-",), np.arange(4, 8))]
    t = table_flip_tables(fixed, [[-1, 1, 0]])
    assert t["dim_sl"] == [0, 0, -1] and t["max_n"] == [4, 4, 0]


def test_sublattice_semantics():
    """smol/moca/sublattice.py:23-110."""
    s = Sublattice(("A", "B"), [3, 1, 2, 2])
    assert s.sites.tolist() == [1, 2, 3] and s.is_active and s.encoding.tolist() == [0, 1]
    s.restrict_sites([2])
    assert s.active_sites.tolist() == [1, 3] and s.restricted_sites.tolist() == [2]
    s.reset_restricted_sites()
    assert s.active_sites.tolist() == [1, 2, 3]
    assert not Sublattice(("O2-",), [0, 1]).is_active
    parts = Sublattice(("A", "B", "C"), np.arange(6)).split_by_species(np.array([0, 1, 2, 0, 1, 2]), [[0, 1], [2]])
    assert parts[0].sites.tolist() == [0, 1, 3, 4] and parts[0].encoding.tolist() == [0, 1]
    assert parts[1].encoding.tolist() == [2] and not parts[1].is_active


def test_shard_walkers_partitions_exactly():
    for W, G in ((4096, 8), (10, 4), (3, 8), (32768, 8)):
        blocks = [shard_walkers(W, G, r) for r in range(G)]
        assert sum(c for _, c in blocks) == W
        assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))


def test_sample_container_accessors():
    class Ens:
        sublattices = [Sublattice(("A", "B"), np.arange(4))]
        natural_parameters = np.array([1.0, 2.0])
        num_energy_coefs = 2
    shapes = {"occupancy": ((4,), np.int32), "features": ((2,), np.float64), "enthalpy": ((1,), np.float64),
              "accepted": ((1,), bool), "n_accepted": ((), np.int32)}
    c = SampleContainer(Ens(), 3, shapes)
    tr = dict(occupancy=np.zeros((5, 3, 4), dtype=np.int8), features=np.ones((5, 3, 2)),
              enthalpy=np.arange(15.0).reshape(5, 3, 1), accepted=np.ones((5, 3, 1), dtype=bool),
              n_accepted=np.full((5, 3), 2, dtype=np.int32))
    c.append(tr, thinned_by=4)
    assert c.num_samples == 5 and c.total_mc_steps == 20 and c.shape == (3, 4)
    assert c.get_occupancies().dtype == np.int32 and c.get_occupancies().shape == (15, 4)
    assert c.get_enthalpies(flat=False).shape == (5, 3, 1)
    assert c.sampling_efficiency() == 1.0 and c.step_efficiency() == 0.5
    assert np.allclose(c.get_energies(), 3.0) and c.get_minimum_enthalpy()[0] == 0.0
    c.clear()
    assert c.num_samples == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "lib_path", lambda: str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load()
