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
    fields = re.findall(r"\b([a-z_0-9]+)(?:\[[A-Za-z_0-9]+\])*;", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert fields == [f[0] for f in capi.LmcRunConfig._fields_]
    # array members keep their extents
    assert ctypes.sizeof(capi.LmcRunConfig().comp_sl_cum) == 8 * capi.LMC_MAX_COMPOSITE * capi.LMC_MAX_SUBLATTICES


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
    # no extra terms: the energies ARE the enthalpies (container.py:210-211)
    # (flat values are squeezed like the reference's, container.py:514-519)
    assert c.get_enthalpies().shape == (15,) and c.get_enthalpies(flat=False).shape == (5, 3, 1)
    assert np.array_equal(c.get_energies(), c.get_enthalpies()) and c.get_minimum_enthalpy() == 0.0
    assert c.get_minimum_energy() == 0.0 and c.get_minimum_energy_occupancy().shape == (4,)
    c.clear()
    assert c.num_samples == 0


def test_sample_container_compositions_and_energies():
    """accessors of smol/moca/sampler/container.py:208-382 on a hand-made chain"""
    sl_a, sl_b = Sublattice(("A", "B"), np.array([0, 1, 2, 3])), Sublattice(("C", "A"), np.array([4, 5]))

    class Ens:
        sublattices = [sl_a, sl_b]
        natural_parameters = np.array([1.0, 2.0, -1.0])      # two energy coefficients + chemical work
        num_energy_coefs = 2
    shapes = {"occupancy": ((6,), np.int32), "features": ((3,), np.float64), "enthalpy": ((1,), np.float64),
              "accepted": ((1,), bool), "n_accepted": ((), np.int32)}
    c = SampleContainer(Ens(), 2, shapes)
    occ = np.array([[[0, 0, 1, 1, 0, 1], [1, 1, 1, 1, 0, 0]],
                    [[0, 1, 1, 1, 1, 1], [0, 0, 0, 0, 1, 0]]], dtype=np.int8)     # [S=2][W=2][N=6]
    feats = np.arange(12.0).reshape(2, 2, 3)
    tr = dict(occupancy=occ, features=feats, enthalpy=(feats @ Ens.natural_parameters)[:, :, None],
              accepted=np.ones((2, 2, 1), dtype=bool), n_accepted=np.ones((2, 2), dtype=np.int32))
    c.append(tr, thinned_by=1)
    # energies drop the chemical-work term and keep the trailing axis of the enthalpy trace
    e = c.get_energies(flat=False)
    assert e.shape == (2, 2, 1) and np.allclose(e[..., 0], feats[..., 0] + 2 * feats[..., 1])
    assert c.get_energies().shape == (4,) and np.isclose(c.get_minimum_energy(), 2.0)
    assert c.get_minimum_energy_occupancy().tolist() == occ[0, 0].tolist()
    sub = c.get_sublattice_species_counts(sl_a, flat=False)
    assert sub.shape == (2, 2, 2) and sub[0, 0].tolist() == [2, 2] and sub[1, 1].tolist() == [4, 0]
    assert c.get_sublattice_species_counts(sl_b).tolist() == [[1, 1], [2, 0], [0, 2], [1, 1]]
    counts = c.get_species_counts(flat=False)
    assert set(counts) == {"A", "B", "C"}                      # species A lives on both sublattices
    # (chain form [walkers, samples], as the reference's subcounts.T gives it)
    assert counts["A"].tolist() == [[3, 3], [0, 5]] and counts["B"].tolist() == [[2, 3], [4, 0]]
    comps = c.get_compositions()
    assert np.allclose(sum(comps.values()), 1.0) and np.allclose(comps["C"], np.array([1, 2, 0, 1]) / 6)
    assert np.allclose(c.mean_composition()["A"], np.mean([3, 0, 3, 5]) / 6)
    assert np.allclose(c.composition_variance()["B"], np.var(np.array([2, 4, 3, 0]) / 6))
    assert np.allclose(c.mean_sublattice_composition(sl_a), [(2 + 0 + 1 + 4) / 16, (2 + 4 + 3 + 0) / 16])
    assert np.allclose(c.sublattice_composition_variance(sl_b), np.var(np.array([[1, 1], [2, 0], [0, 2], [1, 1]]) / 2, axis=0))
    with pytest.raises(ValueError, match="not recognized"):
        c.get_sublattice_species_counts(Sublattice(("X", "Y"), np.array([0, 1])))
    of = c.get_orbit_factors(np.array([0, 1, 1]))
    vals = Ens.natural_parameters * feats.reshape(4, 3)
    assert np.isclose(of[0], vals[:, 0].sum()) and np.isclose(of[1], vals[:, 1:].sum()) and of[2] == 0.0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "lib_path", lambda: str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load()


def test_bias_tables_follow_the_reference_layout():
    """bias.py:216-233, 259-272: per-site tables indexed [site, code]; oracle restatement of the bias change
    equals the difference of full bias values"""
    from smol_b200 import bias as B
    from oracle import lmc_oracle as O
    sub = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 2
    import smol_b200 as S
    proc = S.ClusterExpansionProcessor(sub, scm, np.zeros(sub.num_corr_functions))
    sls = proc.get_sublattices()
    assert [B.get_oxi_state(x) for x in ("Li+", "Mn3+", "Ti4+", "O2-", "F-", "A")] == [1, 3, 4, -2, -1, 0]
    sq = B.SquareChargeBias(sls, penalty=0.5)
    N = proc.num_sites
    assert sq.table.shape == (N, 3) and sq.mode == capi.LMC_BIAS_SQUARE_SUM
    cat = next(s for s in sls if "Li+" in s.species)
    ani = next(s for s in sls if "F-" in s.species)
    np.testing.assert_array_equal(sq.table[cat.sites[0]], [1, 3, 4])
    np.testing.assert_array_equal(sq.table[ani.sites[0]], [-2, -1, 0])
    fr = [{"Li+": 0.5, "Mn3+": 0.25, "Ti4+": 0.25}, {"O2-": 0.75, "F-": 0.25}]
    if "Li+" not in sls[0].species:
        fr = fr[::-1]
    fu = B.FugacityBias(sls, fugacity_fractions=fr)
    np.testing.assert_allclose(fu.table[cat.sites[3]], np.log([0.5, 0.25, 0.25]))
    np.testing.assert_allclose(fu.table[ani.sites[3]], [np.log(0.75), np.log(0.25), 0.0])
    assert B.FugacityBias(sls).fugacity_fractions[0] is not None          # default: equal fractions
    with pytest.raises(ValueError):
        B.SquareChargeBias(sls, penalty=-1.0)
    with pytest.raises(ValueError):
        B.FugacityBias(sls, fugacity_fractions=[{"Li+": 1.0}, {"O2-": 0.5, "F-": 0.5}])
    with pytest.raises(ValueError):
        B.mcbias_factory("no-such-bias", sls)
    # hyperplanes: row r of the table = A[r][dim(site, code)], dims counted sublattice by sublattice
    A = [[1, 3, 4, -2, -1], [1, -2, 0, 0, 0]] if "Li+" in sls[0].species else [[-2, -1, 1, 3, 4], [0, 0, 1, -2, 0]]
    hp = B.SquareHyperplaneBias(sls, A, [0, 3], penalty=0.1)
    assert hp.table.shape == (N, 3, 2) and hp.rows == 2 and list(hp.intercepts) == [0, 3]
    np.testing.assert_array_equal(hp.table[cat.sites[0], :, 0], [1, 3, 4])
    np.testing.assert_array_equal(hp.table[cat.sites[0], :, 1], [1, -2, 0])
    np.testing.assert_array_equal(hp.table[ani.sites[0], :2, 0], [-2, -1])
    with pytest.raises(ValueError):
        B.SquareHyperplaneBias(sls, [[1, 2, 3]], [0])
    # oracle: change == difference of totals (the reference's generic compute_bias_change)
    osl = M.oracle_sublattices(O, sls)
    rng = np.random.default_rng(0)
    occ = M.random_occupancies(sub, scm, 1, seed=3)[0]
    # with A = the charges and b = 0 the hyperplane bias is the charge bias
    occ_t = M.random_occupancies(sub, scm, 1, seed=8)[0]
    assert O.SquareHyperplaneBias(osl, [A[0]], [0], 0.3).compute_bias(occ_t) == O.SquareChargeBias(osl, 0.3).compute_bias(occ_t)
    for ob in (O.SquareChargeBias(osl, 0.3), O.FugacityBias(osl, fr), O.SquareHyperplaneBias(osl, A, [0, 3], 0.2)):
        for _ in range(10):
            s1, s2 = int(rng.choice(cat.sites)), int(rng.choice(ani.sites))
            step = [(s1, int((occ[s1] + 1) % 3)), (s2, int((occ[s2] + 1) % 2))]
            nxt = occ.copy()
            for site, code in step:
                nxt[site] = code
            np.testing.assert_allclose(ob.compute_bias_change(occ, step), ob.compute_bias(nxt) - ob.compute_bias(occ),
                                       rtol=1e-12, atol=1e-12)


def test_interop_extracts_a_smol_like_ensemble():
    """smol_b200.interop reads only attribute names of the reference interface (processor/base.py:59-107,
    expansion.py:324, ewald.py:78-101, composite.py:56-59, ensemble.py:219-321).  smol itself cannot be imported
    here, so stand-ins with exactly those names are used; the packed device tables must equal the ones of the
    natively built ensemble."""
    from types import SimpleNamespace
    import smol_b200 as S
    from smol_b200 import interop

    sub = M.rocksalt_subspace()
    scm = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(5)
    it = L.cluster_interaction_tensors(sub, rng.normal(0, 0.05, sub.num_corr_functions))
    ewm, ewi = L.ewald_matrix(sub, scm)
    mus = {"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2}
    native_p = S.CompositeProcessor(sub, scm)
    native_p.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    native_p.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.1, ewald_matrix=ewm, ewald_inds=ewi))
    native = S.Ensemble(native_p, chemical_potentials=mus)

    # what a live smol ensemble exposes (no allowed_species() on the subspace: the processor carries the list)
    ref_names = ("orbits", "num_orbits", "num_corr_functions", "orbit_multiplicities", "get_orbit_indices",
                 "function_total_multiplicities")
    smol_subspace = SimpleNamespace(**{n: getattr(sub, n) for n in ref_names})
    allowed = [tuple(sp) for sp in sub.allowed_species(scm)]

    class ClusterDecompositionProcessor:      # names as in smol.moca.processor
        cluster_subspace, supercell_matrix, allowed_species = smol_subspace, scm, allowed
        coefs, _interaction_tensors = np.array(sub.orbit_multiplicities, dtype=float), it

    class EwaldProcessor:
        cluster_subspace, supercell_matrix, allowed_species = smol_subspace, scm, allowed
        coefs, ewald_matrix, _ewald_inds, _ewald_term = np.array(0.1), ewm, ewi, None

    class CompositeProcessor:
        cluster_subspace, supercell_matrix, allowed_species = smol_subspace, scm, allowed
        processors = [ClusterDecompositionProcessor(), EwaldProcessor()]

    subl = [SimpleNamespace(site_space={sp: 1.0 / len(s.species) for sp in s.species}, sites=s.sites,
                            active_sites=s.active_sites, encoding=s.encoding) for s in native.sublattices]
    smol_ens = SimpleNamespace(processor=CompositeProcessor(), sublattices=subl, chemical_potentials=mus)
    got = interop.from_smol_ensemble(smol_ens)
    np.testing.assert_array_equal(got.natural_parameters, native.natural_parameters)
    a, b = got.packed_model(), native.packed_model()
    assert len(a.keep) == len(b.keep)
    for x, y in zip(a.keep, b.keep):
        np.testing.assert_array_equal(x, y)
    for f, ctype in capi.LmcModelDesc._fields_:
        if ctype is not ctypes.c_void_p:          # scalars; the arrays behind the pointers were compared above
            assert getattr(a.desc, f) == getattr(b.desc, f), f
    with pytest.raises(NotImplementedError):
        interop.from_smol_processor(SimpleNamespace(cluster_subspace=smol_subspace, supercell_matrix=scm,
                                                    allowed_species=allowed))


def test_ctypes_structs_match_the_c_header(tmp_path):
    """the ctypes mirrors in smol_b200/_capi.py have the size and field offsets gcc gives the structs of
    include/lmc.h (an ABI drift here would silently shift every pointer of a run configuration)"""
    import ctypes
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {"LmcRunConfig": capi.LmcRunConfig, "LmcWangLandau": capi.LmcWangLandau, "LmcModelDesc": capi.LmcModelDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lmc.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in out:
        if not line:
            continue
        name, field, value = line.split()
        cls = structs[name]
        expect = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == expect, f"{name}.{field}: C {value} vs ctypes {expect}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())


def test_bias_tables_match_reference_python_bias_objects():
    """host-side tables of smol_b200/bias.py (what the device sums) against the tables of the reference's own
    FugacityBias / SquareChargeBias / SquareHyperplaneBias objects (tests/golden/ref_python_steps.npz)"""
    import importlib.util
    from smol_b200.bias import mcbias_factory
    path = os.path.join(os.path.dirname(__file__), "golden", "make_reference_python_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_python_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_steps.npz"))
    factory, occ0 = mod.table_flip_model()
    subl = [Sublattice(s.species, s.sites) for s in factory().sublattices]
    n = np.arange(occ0.shape[1])
    fug = mcbias_factory(*mod.BIAS_CASES["fugacity"][:1], subl, **mod.BIAS_CASES["fugacity"][1])
    np.testing.assert_allclose(fug.table, np.log(gold["bias_fugacity_table"]), rtol=1e-15, atol=0)
    np.testing.assert_allclose([fug.table[n, o].sum() for o in occ0], gold["bias_fugacity_values"], rtol=1e-13)
    chg = mcbias_factory(*mod.BIAS_CASES["charge"][:1], subl, **mod.BIAS_CASES["charge"][1])
    np.testing.assert_array_equal(chg.table, gold["bias_charge_table"])
    np.testing.assert_allclose([-chg.penalty * chg.table[n, o].sum() ** 2 for o in occ0], gold["bias_charge_values"], rtol=1e-13)
    hyp = mcbias_factory(*mod.BIAS_CASES["hyperplane"][:1], subl, **mod.BIAS_CASES["hyperplane"][1])
    np.testing.assert_array_equal(hyp._dim_ids_table, gold["bias_hyperplane_dim_ids"])
    resid = [hyp.table[n, o].sum(axis=0) - hyp.intercepts for o in occ0]
    np.testing.assert_allclose([-hyp.penalty * (r ** 2).sum() for r in resid], gold["bias_hyperplane_values"], rtol=1e-13)


def test_table_flip_tables_match_reference_python_usher():
    """dimension layout, per-dimension site counts, weights of smol_b200.sampler.table_flip_tables against the
    attributes of the reference's own TableFlip usher (tests/golden/ref_python_steps.npz)"""
    import importlib.util
    from smol_b200.sampler import table_flip_tables
    path = os.path.join(os.path.dirname(__file__), "golden", "make_reference_python_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_python_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_steps.npz"))
    factory, _ = mod.table_flip_model()
    subl = [Sublattice(s.species, s.sites) for s in factory().sublattices]
    t = table_flip_tables(subl, mod.TF_TABLE, swap_weight=0.2)
    assert t["num_dims"] == int(gold["tf_d"][0]) and list(t["max_n"]) == gold["tf_max_n"].tolist()
    np.testing.assert_array_equal(t["table"], gold["tf_table"])
    np.testing.assert_array_equal(t["weights"], gold["tf_weights"])
    assert gold["tf_dim_ids"].tolist() == list(range(t["num_dims"]))            # dims run sublattice by sublattice
    # (site, code) -> dimension, the reference's active-sites table
    active = [s for s in subl if len(s.active_sites) > 0]
    ours = np.full(gold["tf_dim_ids_active"].shape, -1, dtype=int)
    for d, (a, code) in enumerate(zip(t["dim_sl"], t["dim_code"])):
        if a >= 0:
            ours[active[a].active_sites, code] = d
    np.testing.assert_array_equal(ours, gold["tf_dim_ids_active"])


def test_ensemble_mu_table_matches_reference_python_ensemble():
    """smol_b200.Ensemble (host side): chemical-potential table, natural parameters and the number of energy
    coefficients against the reference's own Ensemble object (tests/golden/ref_python_steps.npz)"""
    import importlib.util
    import smol_b200 as S
    path = os.path.join(os.path.dirname(__file__), "golden", "make_reference_python_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_python_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python_steps.npz"))
    sub, scm, _ = mod.processor_cases()["rs2of"]
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, mod.TF_INTERACTIONS()), chemical_potentials=dict(mod.TF_MUS))
    np.testing.assert_array_equal(ens.mu_table, gold["ens_mu_table"])
    np.testing.assert_allclose(ens.natural_parameters, gold["ens_natural_parameters"], rtol=1e-15, atol=0)
    assert ens.num_energy_coefs == int(gold["ens_num_energy_coefs"][0]) and ens.natural_parameters[-1] == -1.0
    ens.chemical_potentials = None                     # ChemicalPotentialManager.__delete__ (ensemble.py:72-84)
    assert len(ens.natural_parameters) == ens.num_energy_coefs and ens.mu_table is None


def test_engine_limits_fail_loudly_without_a_device():
    """Limits the reference does not have.  The ones the Python table builder or the argument
    checks of ``lmc_model_create`` see are raised before any CUDA call, with a message naming the limit; the ones
    that need the device tables (strides > 255, codes >= 8 after splitting) are in tests/test_gpu_api.py."""
    import ctypes as C
    import smol_b200 as S
    from smol_b200.model import PackedModel
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 2
    proc = S.ClusterDecompositionProcessor(sub, scm, L.cluster_interaction_tensors(sub, M.fcc_coefs(sub)))
    lib = capi.load()

    def create(mutate):
        pm = PackedModel(proc.num_sites, proc.coefs, proc.get_sublattices(), **proc._tables())
        mutate(pm.desc)
        handle = C.c_void_p()
        rc = lib.lmc_model_create(C.byref(pm.desc), C.byref(handle))
        return rc, lib.lmc_last_error().decode()

    rc, msg = create(lambda d: setattr(d, "num_sites", 70000))
    assert rc < 0 and "65535" in msg                                       # u16 site indices
    rc, msg = create(lambda d: setattr(d, "num_sublattices", 9))
    assert rc < 0 and "sublattices" in msg
    rc, msg = create(lambda d: setattr(d, "tf_num_dims", 17))
    assert rc < 0 and "flip table" in msg
    rc, msg = create(lambda d: setattr(d, "abi_version", 1))
    assert rc < 0 and "ABI" in msg
    # Python side: clusters of more than four sites, more than eight species codes
    big = L.ClusterSubspace.from_cutoffs(L.fcc_prim(), {2: 3.0, 5: 3.0})
    if any(o.num_sites > 4 for o in big.orbits):
        with pytest.raises(ValueError, match="more than 4 sites"):
            p5 = S.ClusterExpansionProcessor(big, scm, np.zeros(big.num_corr_functions))
            PackedModel(p5.num_sites, p5.coefs, p5.get_sublattices(), **p5._tables())
    nine = L.ClusterSubspace.from_cutoffs(L.fcc_prim(species=tuple("A%d" % i for i in range(9))), {2: 3.0})
    p9 = S.ClusterExpansionProcessor(nine, scm, np.zeros(nine.num_corr_functions))
    with pytest.raises(ValueError, match="species codes"):
        PackedModel(p9.num_sites, p9.coefs, p9.get_sublattices(), **p9._tables())
    # more than four changed sites per step
    from smol_b200.sampler import table_flip_tables
    with pytest.raises(ValueError, match="more than 4 sites"):
        table_flip_tables(proc.get_sublattices(), [[-5, 5]])


@pytest.mark.skipif(not (os.path.isdir("/root/reference/smol") and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "smol"))),
                    reason="needs the reference sources and its compiled evaluators (build container only)")
def test_interop_extracts_live_reference_objects():
    """processor/expansion.py:142-156, ensemble.py:89-99, processor/ewald.py:76-101: the reference's REAL Ensemble /
    ClusterDecompositionProcessor / EwaldProcessor / CompositeProcessor instances go through
    smol_b200.interop.from_smol_ensemble (fresh process: the import shells rewire sys.modules)"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "extractor_reference_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res == {"semigrand_decomposition": True, "canonical_composite_ewald": True}


def test_sublattice_order_and_use_concentration():
    """processor/base.py:75-82: sublattices in the order of the SORTED site spaces (species compared as pymatgen does:
    electronegativity, symbol, oxidation state), whatever the order of the sites; use_concentration is refused loudly"""
    import smol_b200 as S
    from smol_b200.processor import species_sort_key
    assert sorted(["O2-", "F-", "Li+", "Mn3+", "Ti4+", "Mn2+"], key=species_sort_key) == \
        ["Li+", "Ti4+", "Mn2+", "Mn3+", "O2-", "F-"]
    # anion site first in the primitive cell: the sublattices still come cations first (Li < O by electronegativity)
    lat = 0.5 * 4.2 * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float)
    prim = L.PrimCell(lat, [[0.5, 0.5, 0.5], [0, 0, 0]], [("O2-", "F-"), ("Li+", "Mn3+")],
                      {"Li+": 1, "Mn3+": 3, "O2-": -2, "F-": -1})
    sub = L.ClusterSubspace.from_cutoffs(prim, {2: 3.1})
    proc = S.ClusterExpansionProcessor(sub, np.eye(3, dtype=int) * 2, np.zeros(sub.num_corr_functions))
    subl = proc.get_sublattices()
    assert [s.species for s in subl] == [("Li+", "Mn3+"), ("O2-", "F-")]
    assert subl[0].sites.tolist() == list(range(8, 16)) and subl[1].sites.tolist() == list(range(8))
    with pytest.raises(NotImplementedError, match="use_concentration"):
        S.ClusterExpansionProcessor(sub, np.eye(3, dtype=int), np.zeros(sub.num_corr_functions), use_concentration=True)
    with pytest.raises(NotImplementedError, match="use_concentration"):
        S.CompositeProcessor(sub, np.eye(3, dtype=int), use_concentration=True)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference --config 2`: the CPU arm alone, same JSON shape (no GPU needed)"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "2", "--steps", "1",
                        "--warmup", "0", "--ref-seconds", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["value"] > 0 and line["higher_is_better"]
    assert line["cpu_baseline"]["kind"] in ("reference", "oracle") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["baseline_config"] == 2 and line["gpu_launches"] == 0
