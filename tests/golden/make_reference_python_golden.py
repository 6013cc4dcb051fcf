"""Golden step records from the reference's OWN Python kernels (run in the build container only).

``smol.moca`` cannot be imported as a package here (pymatgen / monty / h5py are not installed), but the modules of
the step loop itself -- ``smol/moca/kernel/{base,metropolis,wanglandau,mcusher}.py``, ``smol/moca/trace.py``,
``smol/utils/math.py`` ... -- only need ``monty.json`` / ``monty.dev`` names and two helper modules at import time.
This script registers empty package shells for ``smol`` (so that no ``__init__`` pulls pymatgen in), four stub modules,
imports the UNMODIFIED reference files from ``/root/reference`` and drives their ``Metropolis`` / ``WangLandau``
kernels with ``Flip`` / ``Swap`` ushers step by step.

The reference draws from ``numpy.random.Generator`` (a data-dependent number of draws per step); the oracle and the
CUDA engine use counter-based Philox words at fixed positions.  To compare STEP SEMANTICS the kernels' generator is
replaced by ``ScriptedRng``: the n-th ``choice`` / ``random`` call of a step returns what the oracle derives from the
same Philox word (``choice(seq, p)`` = first index whose cumulative probability exceeds u01(word 0), ``choice(seq)`` =
``seq[mulhi32(word, len(seq))]`` for words 1 and 2, ``random()`` = u01(word 3)).  Everything else -- proposal
logic, feature / enthalpy deltas, acceptance rule, occupancy update, Wang-Landau bookkeeping -- is the reference's code.

Feature vectors come from the oracle's processors (pure-Python evaluator restatements, themselves pinned against the
reference's compiled evaluators in ``ref_vectors.npz``) through a duck-typed ensemble.

Output: ``tests/golden/ref_python_steps.npz`` (per-step proposals, acceptance flags, enthalpy changes, occupancy
snapshots, final Wang-Landau arrays).  ``tests/test_oracle_golden.py`` replays the oracle against it anywhere.

    python tests/golden/make_reference_python_golden.py
"""
import importlib
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/smol"


def import_reference_kernels():
    def pkg(name, paths):
        m = types.ModuleType(name)
        m.__path__ = paths
        m.__package__ = name
        sys.modules[name] = m
    for name, sub in [("smol", ""), ("smol.moca", "/moca"), ("smol.moca.kernel", "/moca/kernel"),
                      ("smol.utils", "/utils"), ("smol.cofe", "/cofe"), ("smol.cofe.space", "/cofe/space"),
                      ("smol.moca.composition", "/moca/composition")]:
        pkg(name, [REF + sub])
    mj = types.ModuleType("monty.json")
    mj.MSONable = type("MSONable", (), {})
    mj.jsanitize = lambda x, **k: x
    mj.MontyDecoder = object
    md = types.ModuleType("monty.dev")
    md.requires = lambda cond, msg: (lambda f: f)
    sys.modules.update({"monty": types.ModuleType("monty"), "monty.json": mj, "monty.dev": md})
    sp = types.ModuleType("smol.moca.composition.space")       # only TableFlip / the charge bias use these
    class CompositionSpace:                 # TableFlip with a given flip table only reads the constraint system
        def __init__(self, bits, sublattice_sizes, **kwargs):
            d = sum(len(b) for b in bits)
            self._A, self._b, self.flip_table = np.zeros((0, d)), np.zeros(0), None
    sp.CompositionSpace = CompositionSpace
    def get_oxi_state(label):               # 'Mn3+' -> 3, 'O2-' -> -2, 'F-' -> -1, 'A' -> 0
        import re
        m = re.search(r"(\d*)([+-])$", str(label))
        return 0 if not m else (int(m.group(1) or 1) * (1 if m.group(2) == "+" else -1))
    sp.get_oxi_state = get_oxi_state
    sys.modules["smol.moca.composition.space"] = sp
    dm = types.ModuleType("smol.cofe.space.domain")
    dm.get_species = lambda s: s
    dm.Vacancy = type("Vacancy", (), {})
    sys.modules["smol.cofe.space.domain"] = dm
    met = importlib.import_module("smol.moca.kernel.metropolis")
    wl = importlib.import_module("smol.moca.kernel.wanglandau")
    assert met.__file__.startswith(REF) and wl.__file__.startswith(REF)
    return met.Metropolis, wl.WangLandau


class ScriptedRng:
    """Stands in for the kernels' numpy Generator: values follow the oracle's Philox word positions."""

    def __init__(self, O, seed, walker):
        self.O, self.seed, self.walker = O, seed, walker
        self.rnd, self.plain = None, 0

    def begin_step(self, t):
        self.rnd, self.plain = self.O.StepRandom(self.seed, self.walker, t), 0

    def choice(self, a, p=None):
        if p is not None:                      # MCUsher.get_random_sublattice
            u = self.O.u01(self.rnd.word(0))
            cdf = np.cumsum(np.asarray(p, dtype=np.float64))
            cdf[-1] = 1.0
            return a[int(np.searchsorted(cdf, u, side="right"))] if len(a) > 1 else a[0]
        self.plain += 1
        assert self.plain <= 2
        return a[self.O.mulhi32(self.rnd.word(self.plain), len(a))]

    def random(self):
        return self.O.u01(self.rnd.word(3))


class TableFlipRng(ScriptedRng):
    """TableFlip.propose_step (mcusher.py:553-639): random() #1 = swap decision (word 0), choose_section_from_partition
    = rng.choice(n, p) (word 1), choice(list, size=m, replace=False) = m sequential bounded draws from the shrinking
    list (words 4, 5, ...: the oracle's restatement of numpy's sampling without replacement), last random() = accept
    (word 3).  The fallback Swap usher of the reference owns a separate generator (mcusher.py:539: Swap(sublattices)
    without rng); it gets ``swapper()`` below (words 4, 5, 6)."""

    def begin_step(self, t):
        super().begin_step(t)
        self.nrandom, self.pick = 0, 4

    def random(self):
        self.nrandom += 1
        return self.O.u01(self.rnd.word(0 if self.nrandom == 1 else 3))

    def choice(self, a, p=None, size=None, replace=True):
        if p is not None:
            n = a if isinstance(a, (int, np.integer)) else len(a)
            u = self.O.u01(self.rnd.word(1))
            cdf = np.cumsum(np.asarray(p, dtype=np.float64))
            i = int(np.searchsorted(cdf, u, side="right"))
            i = min(i, n - 1)
            return i if isinstance(a, (int, np.integer)) else a[i]
        assert size is not None and replace is False
        pool, out = list(a), []
        for _ in range(int(size)):
            out.append(pool.pop(self.O.mulhi32(self.rnd.word(self.pick), len(pool))))
            self.pick += 1
        return np.array(out, dtype=int)

    def swapper(self):
        parent = self

        class _Sw:
            def choice(self, a, p=None):
                if p is not None:
                    u = parent.O.u01(parent.rnd.word(4))
                    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
                    cdf[-1] = 1.0
                    return a[int(np.searchsorted(cdf, u, side="right"))] if len(a) > 1 else a[0]
                parent.plain += 1
                return a[parent.O.mulhi32(parent.rnd.word(4 + parent.plain), len(a))]
        return _Sw()


class AutoRng(ScriptedRng):
    """for runs driven by the reference's own Sampler loop (no hook between steps): a step of a Flip / Swap usher
    always begins with the sublattice draw, which advances the step counter here"""

    def __init__(self, O, seed, walker):
        super().__init__(O, seed, walker)
        self.t = -1

    def choice(self, a, p=None):
        if p is not None:
            self.t += 1
            self.begin_step(self.t)
        return super().choice(a, p)


def import_reference_sampler():
    """smol/moca/sampler/{sampler,container}.py unmodified (after import_reference_kernels): h5py, the MSON encoder,
    smol.moca.Ensemble / Sublattice are only named at import time"""
    from oracle import lmc_oracle as O
    sys.modules["h5py"] = types.ModuleType("h5py")
    sys.modules["monty.json"].MontyEncoder = object
    sl = types.ModuleType("smol.moca.sublattice")
    sl.Sublattice = O.Sublattice
    sys.modules["smol.moca.sublattice"] = sl
    sys.modules["smol.moca"].Ensemble = O.Ensemble
    m = types.ModuleType("smol.moca.sampler")
    m.__path__ = [REF + "/moca/sampler"]
    sys.modules["smol.moca.sampler"] = m
    spec = importlib.util.spec_from_file_location("smol.moca.kernel._init", REF + "/moca/kernel/__init__.py")
    ki = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ki)
    sys.modules["smol.moca.kernel"].mckernel_factory = ki.mckernel_factory
    return importlib.import_module("smol.moca.sampler.sampler").Sampler


def import_reference_processors():
    """smol/moca/processor/{base,expansion}.py unmodified, on top of the reference's COMPILED evaluators in
    oracle/_ref (after import_reference_kernels / import_reference_sampler).  pymatgen's Structure is only copied,
    enlarged and measured by the constructors, so a four-method stand-in is enough; the cluster subspace is this
    build's geometry-made one behind an adapter exposing the attribute names the reference reads."""
    cref = os.path.join(ROOT, "oracle", "_ref", "smol")
    assert os.path.isdir(cref + "/utils/cluster"), "build oracle/_ref first (python oracle/build_ref.py)"
    sys.modules["smol.utils"].__path__ = [REF + "/utils", cref + "/utils"]
    spec = importlib.util.spec_from_file_location(
        "smol.utils.cluster", REF + "/utils/cluster/__init__.py",
        submodule_search_locations=[cref + "/utils/cluster", REF + "/utils/cluster"])
    uc = importlib.util.module_from_spec(spec)
    sys.modules["smol.utils.cluster"] = uc
    spec.loader.exec_module(uc)
    container = importlib.import_module("smol.utils.cluster.container")
    pm = types.ModuleType("pymatgen.core")
    pm.PeriodicSite = pm.Structure = object
    sys.modules.update({"pymatgen": types.ModuleType("pymatgen"), "pymatgen.core": pm})
    space = sys.modules["smol.cofe.space"]
    space.Vacancy = sys.modules["smol.cofe.space.domain"].Vacancy
    space.get_allowed_species = lambda structure: structure.allowed
    space.get_site_spaces = lambda structure, include_measure=False: list(structure.allowed)
    cs = types.ModuleType("smol.cofe.space.clusterspace")

    class OrbitIndices:
        def __init__(self, arrays, container):
            self.arrays, self.container = arrays, container
    cs.OrbitIndices = OrbitIndices

    class FakeStructure:
        def __init__(self, prim_spaces, size=1):
            self.prim_spaces, self.size = prim_spaces, size
            self.allowed = [sp for sp in prim_spaces for _ in range(size)]      # basis major, like the tables

        def copy(self):
            return FakeStructure(self.prim_spaces, self.size)

        def make_supercell(self, scm):
            self.__init__(self.prim_spaces, int(round(abs(np.linalg.det(np.asarray(scm, dtype=float))))))

        def __len__(self):
            return len(self.allowed)

    class RefSubspace:
        """adapter: attribute names of smol.cofe.ClusterSubspace over this build's lattice.ClusterSubspace"""
        def __init__(self, sub):
            self._sub = sub
            self.structure = FakeStructure([tuple(s) for s in sub.prim.site_spaces])
            self.orbits, self.num_threads = sub.orbits, 1
            self.num_orbits, self.num_corr_functions = sub.num_orbits, sub.num_corr_functions
            self.orbit_multiplicities = sub.orbit_multiplicities

        external_terms = ()

        def __len__(self):                       # clusterspace.py: number of correlation functions (+ external terms)
            return self.num_corr_functions

        @property
        def orbits_by_diameter(self):
            """clusterspace.py:367-381: {diameter rounded to 6 decimals: orbits}, ascending"""
            from itertools import groupby
            key = lambda orb: float(np.round(orb.diameter, 6))                      # noqa: E731
            return {d: tuple(orbs) for d, orbs in groupby(sorted(self.orbits, key=key), key=key)}

        def num_prims_from_matrix(self, scm):
            return self._sub.supercell_size(scm)

        def get_orbit_indices(self, scm):
            arrays = self._sub.get_orbit_indices(scm).arrays
            return OrbitIndices(arrays, container.IntArray2DContainer(arrays))
    cs.ClusterSubspace = RefSubspace
    sys.modules["smol.cofe.space.clusterspace"] = cs
    pkg = types.ModuleType("smol.moca.processor")
    pkg.__path__ = [REF + "/moca/processor"]
    sys.modules["smol.moca.processor"] = pkg
    ex = importlib.import_module("smol.moca.processor.expansion")
    assert ex.__file__.startswith(REF)
    pkg.ClusterExpansionProcessor, pkg.ClusterDecompositionProcessor = (ex.ClusterExpansionProcessor,
                                                                        ex.ClusterDecompositionProcessor)
    return ex.ClusterExpansionProcessor, ex.ClusterDecompositionProcessor, RefSubspace


def processor_cases():
    """name -> (subspace, supercell matrix, correlation coefficients); flips are drawn in processor_flips"""
    from tests import models as M
    out = {}
    for name, sub, scm, seed in [("fcc3", M.fcc_subspace(), np.eye(3, dtype=int) * 3, 5),
                                 ("fcc421", M.fcc_subspace(), np.diag([4, 2, 1]), 6),
                                 ("rs2of", M.rocksalt_subspace(anions=("O2-", "F-")), np.eye(3, dtype=int) * 2, 7)]:
        out[name] = (sub, scm, np.random.default_rng(seed).normal(0, 0.05, sub.num_corr_functions))
    return out


def processor_flips(sub, scm, seed, n=12):
    """seeded occupancies and 1..3-site flip lists (sequential, a site may repeat)"""
    from tests import models as M
    rng = np.random.default_rng(seed)
    spaces = sub.allowed_species(scm)
    active = [i for i, s in enumerate(spaces) if len(s) > 1]
    occs = M.random_occupancies(sub, scm, n, seed=seed)
    flips = []
    for w in range(n):
        cur = occs[w].copy()
        fl = []
        for _ in range(1 + w % 3):
            site = int(rng.choice(active))
            code = int(rng.choice([c for c in range(len(spaces[site])) if c != cur[site]]))
            fl.append((site, code))
            cur[site] = code
        flips.append(fl)
    return occs, flips


SQS_T, SQS_TOL, SQS_W = 3000.0, 1e-5, 0.05   # light weight on the matched diameter: the chain keeps moving
DIST_TOL = 0.12      # loose on purpose: the largest exactly-matched diameter L takes intermediate values
DIST_TOL_INT = 0.004  # (cluster interactions are ~coefficient sized)
CONTAINER_QUERIES = [dict(discard=0, thin_by=1), dict(discard=7, thin_by=3)]


def container_answers(c, sublattices, flat):
    """every accessor of the reference's SampleContainer that this build mirrors, as a flat name -> array dict"""
    out = {}
    for qi, q in enumerate(CONTAINER_QUERIES):
        tag = f"q{qi}_{'flat' if flat else 'chain'}_"
        kw = dict(q, flat=flat)
        for name in ("get_energies", "get_enthalpies", "get_feature_vectors", "mean_energy", "energy_variance",
                     "mean_enthalpy", "enthalpy_variance", "mean_feature_vector", "feature_vector_variance",
                     "get_minimum_energy", "get_minimum_enthalpy", "get_minimum_energy_occupancy",
                     "get_minimum_enthalpy_occupancy"):
            out[tag + name] = np.asarray(getattr(c, name)(**kw))
        out[tag + "sampling_efficiency"] = np.asarray(c.sampling_efficiency(discard=q["discard"], flat=flat))
        for name in ("get_compositions", "mean_composition", "composition_variance"):
            for sp, v in getattr(c, name)(**kw).items():
                out[tag + name + ":" + str(sp)] = np.asarray(v)
        for i, sl in enumerate(sublattices):
            for name in ("get_sublattice_species_counts", "get_sublattice_compositions",
                         "mean_sublattice_composition", "sublattice_composition_variance"):
                out[tag + name + f":{i}"] = np.asarray(getattr(c, name)(sl, **kw))
    return out


class PickRng:
    """generator of a Composite / MultiStep usher itself: its one draw per step, rng.choice(seq, p=p), is random word 4"""

    def __init__(self, parent):
        self.parent = parent

    def choice(self, a, p=None):
        O = self.parent.O
        u = O.u01(self.parent.rnd.word(4))
        cdf = np.cumsum(np.asarray(p, dtype=np.float64))
        cdf[-1] = 1.0
        return a[int(np.searchsorted(cdf, u, side="right"))]


class ChainRng:
    """generator of MultiStep's inner usher: proposal j of a step (each starts with the sublattice draw) reads the
    words of Philox block 2 + j"""

    def __init__(self, parent):
        self.parent, self.step_seen, self.j, self.plain = parent, None, -1, 0

    def choice(self, a, p=None):
        par, O = self.parent, self.parent.O
        if self.step_seen is not par.rnd:                    # first draw of a new step
            self.step_seen, self.j = par.rnd, -1
        if p is not None:
            self.j += 1
            self.plain = 0
            u = O.u01(par.rnd.word(8 + 4 * self.j))
            cdf = np.cumsum(np.asarray(p, dtype=np.float64))
            cdf[-1] = 1.0
            return a[int(np.searchsorted(cdf, u, side="right"))] if len(a) > 1 else a[0]
        self.plain += 1
        return a[O.mulhi32(par.rnd.word(8 + 4 * self.j + self.plain), len(a))]


class _SiteSpace(dict):
    """species -> concentration in site-space order, with the one pymatgen-flavoured method MCBias.__init__ calls"""

    def as_dict(self):
        return {"composition": dict(self)}


BIAS_CASES = {
    "fugacity": ("fugacity-bias", dict(fugacity_fractions=[{"Li+": 0.5, "Mn3+": 0.25, "Ti4+": 0.25},
                                                           {"O2-": 0.75, "F-": 0.25}])),
    "charge": ("square-charge-bias", dict(penalty=0.3)),
    "hyperplane": ("square-hyperplane-bias", dict(hyperplane_normals=[[1, 3, 4, -2, -1], [1, 1, 1, 0, 0]],
                                                  hyperplane_intercepts=[0, 8], penalty=0.2)),
}
USHER_CASES = ("composite", "multistep-flip", "multistep-swap")
TF_TABLE = [[-1, 1, 0, 2, -2], [0, -1, 1, 1, -1]]     # charge-neutral, site-conserving flips


TF_MUS = {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0}


def TF_INTERACTIONS():
    """interaction tensors of table_flip_model()"""
    from smol_b200 import lattice as L
    from tests import models as M
    rs = M.rocksalt_subspace(anions=("O2-", "F-"))
    return L.cluster_interaction_tensors(rs, np.random.default_rng(21).normal(0, 0.03, rs.num_corr_functions))


def table_flip_model():
    """5-species rocksalt 2x2x2 (8 cation + 8 anion sites), charge-neutral starts: (ensemble factory, occupancies)"""
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    from tests import models as M
    rs = M.rocksalt_subspace(anions=("O2-", "F-"))
    scm = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(21)
    it = L.cluster_interaction_tensors(rs, rng.normal(0, 0.03, rs.num_corr_functions))
    spaces = rs.allowed_species(scm)
    cat = np.array([i for i, s in enumerate(spaces) if len(s) == 3])
    ani = np.array([i for i, s in enumerate(spaces) if len(s) == 2])
    subl = [O.Sublattice(("Li+", "Mn3+", "Ti4+"), cat), O.Sublattice(("O2-", "F-"), ani)]
    mus = {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0}
    occ0 = np.zeros((2, len(spaces)), dtype=np.int32)
    for w in range(2):      # 6 Li+, 1 Mn3+, 1 Ti4+ (charge 13) against 5 O2- and 3 F-
        occ0[w, cat] = rng.permutation([0] * 6 + [1, 2])
        occ0[w, ani] = rng.permutation([0] * 5 + [1] * 3)
    return (lambda: O.Ensemble(O.ClusterDecompositionProcessor(rs, scm, it), subl, chemical_potentials=mus)), occ0


def models():
    """name -> (oracle ensemble factory, initial occupancies [W][N]); rebuilt identically by the test"""
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    from tests import models as M
    out = {}
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * 3
    it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub, seed=7))
    subl = [O.Sublattice(("A", "B"), np.arange(27))]
    out["fcc3"] = (lambda: O.Ensemble(O.ClusterDecompositionProcessor(sub, scm, it), subl),
                   M.random_occupancies(sub, scm, 2, seed=4, balanced=False))
    rs = M.rocksalt_subspace()
    scm2 = np.eye(3, dtype=int) * 2
    rng = np.random.default_rng(11)
    it2 = L.cluster_interaction_tensors(rs, rng.normal(0, 0.05, rs.num_corr_functions))
    spaces = rs.allowed_species(scm2)
    cat = np.array([i for i, s in enumerate(spaces) if len(s) > 1])
    ani = np.array([i for i, s in enumerate(spaces) if len(s) == 1])
    subl2 = [O.Sublattice(("Li+", "Mn3+", "Ti4+"), cat), O.Sublattice(("O2-",), ani)]
    mus = {"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2}
    out["rs2"] = (lambda: O.Ensemble(O.ClusterDecompositionProcessor(rs, scm2, it2), subl2, chemical_potentials=mus),
                  M.random_occupancies(rs, scm2, 2, seed=2))
    return out


MC_SHAPES = [np.diag([2, 2, 2]), np.diag([4, 2, 1]), np.array([[2, 1, 0], [0, 2, 0], [0, 0, 2]])]


def multicell_model():
    """(shapes, oracle ensemble factory per shape, initial occupancies [W][K][N]) of the multicell case"""
    from oracle import lmc_oracle as O
    from smol_b200 import lattice as L
    from tests import models as M
    sub = M.fcc_subspace()
    it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub))
    subl = [O.Sublattice(("A", "B"), np.arange(8))]
    occ0 = np.stack([np.stack([M.random_occupancies(sub, scm, 1, seed=100 * w + k, balanced=True)[0]
                               for k, scm in enumerate(MC_SHAPES)]) for w in range(2)])
    return MC_SHAPES, (lambda k: O.Ensemble(O.ClusterDecompositionProcessor(sub, MC_SHAPES[k], it), subl)), occ0


def record(kernel, rngs, occ, nsteps, snap_every):
    """drive ONE reference kernel: per-step (accepted, proposal, enthalpy change), occupancy snapshots"""
    occ = np.array(occ, dtype=np.int32)
    acc = np.zeros(nsteps, dtype=bool)
    prop = np.full((nsteps, 4, 2), -1, dtype=np.int64)
    dh = np.zeros(nsteps)
    db = np.zeros(nsteps)
    snaps = []
    kernel.set_aux_state(occ)
    for t in range(nsteps):
        rngs.begin_step(t)
        # (the proposal is re-derived here only to RECORD it: same scripted words, no state change)
        step = kernel.mcusher.propose_step(occ)
        for j, (s, c) in enumerate(step):
            prop[t, j] = (s, c)
        rngs.begin_step(t)
        trace = kernel.single_step(occ)
        acc[t] = bool(trace.accepted)
        dh[t] = float(trace.delta_trace.enthalpy)
        if kernel.bias is not None:
            db[t] = float(trace.delta_trace.bias)
        if (t + 1) % snap_every == 0:
            snaps.append(occ.copy())
    record.last_dbias = db
    return acc, prop, dh, np.array(snaps)


def reference_ensemble_objects():
    """LIVE objects of the reference's own classes (smol/moca/ensemble.py, processor/{expansion,ewald,composite}.py,
    unmodified, imported behind the package shells) for the extractor test of smol_b200.interop:
    name -> (reference Ensemble, (subspace, supercell matrix, interaction tensors, Ewald (matrix, inds, coef) | None,
    chemical potentials)).  Fresh process only (it rewires sys.modules)."""
    from smol_b200 import lattice as L
    import_reference_kernels()
    import_reference_sampler()
    CE, CD, RefSubspace = import_reference_processors()
    cm = types.ModuleType("pymatgen.core.composition")
    cm.ChemicalPotential = dict
    sys.modules["pymatgen.core.composition"] = cm
    pp = sys.modules["smol.moca.processor"]
    pa = types.ModuleType("pymatgen.analysis")
    pe = types.ModuleType("pymatgen.analysis.ewald")
    pe.EwaldSummation = lambda *a, **k: None
    sys.modules.update({"pymatgen.analysis": pa, "pymatgen.analysis.ewald": pe})
    ext = types.ModuleType("smol.cofe.extern")
    ext.__path__ = [REF + "/cofe/extern"]
    sys.modules["smol.cofe.extern"] = ext
    sys.modules["smol.cofe.space.domain"].get_allowed_species = sys.modules["smol.cofe.space"].get_allowed_species
    RealEwaldTerm = importlib.import_module("smol.cofe.extern.ewald").EwaldTerm
    RefEwaldProcessor = importlib.import_module("smol.moca.processor.ewald").EwaldProcessor
    RefComposite = importlib.import_module("smol.moca.processor.composite").CompositeProcessor
    pp.CompositeProcessor, pp.EwaldProcessor = RefComposite, RefEwaldProcessor
    RefEnsemble = importlib.import_module("smol.moca.ensemble").Ensemble
    factory, _ = table_flip_model()
    sub, scm, coefs = processor_cases()["rs2of"]
    it = TF_INTERACTIONS()
    out = {}

    def sublattices():
        o_ens = factory()
        for sl in o_ens.sublattices:
            sl.site_space = _SiteSpace({spc: 1.0 / len(sl.species) for spc in sl.species})
        return o_ens.sublattices
    out["semigrand_decomposition"] = (RefEnsemble(CD(RefSubspace(sub), scm, it), sublattices=sublattices(),
                                                  chemical_potentials=dict(TF_MUS)), (sub, scm, it, None, dict(TF_MUS)))
    ewm, ewi = L.ewald_matrix(sub, scm)

    class GivenMatrixTerm(RealEwaldTerm):
        def get_ewald_structure(self, structure):
            return None, ewi

        def get_ewald_matrix(self, ewald_summation):
            return ewm
    term = GivenMatrixTerm()
    rsub = RefSubspace(sub)
    rsub.external_terms = (term,)
    rsub.add_external_term = lambda t: None
    comp = RefComposite(rsub, scm)
    comp.add_processor(CD(rsub, scm, it))
    comp.add_processor(RefEwaldProcessor(rsub, scm, term, coefficient=0.1))
    out["canonical_composite_ewald"] = (RefEnsemble(comp, sublattices=sublattices()), (sub, scm, it, (ewm, ewi, 0.1), None))
    return out


def main():
    from oracle import lmc_oracle as O
    Metropolis, WangLandau = import_reference_kernels()
    mods = models()
    out = {}
    nsteps, snap = 300, 25
    for name, step_type, T in [("fcc3", "swap", 2000.0), ("fcc3", "flip", 2500.0), ("rs2", "flip", 3000.0)]:
        factory, occ0 = mods[name]
        for w in range(len(occ0)):
            ens = factory()
            seed = 1000 + 17 * w
            k = Metropolis(ens, step_type, T, seed=seed)
            rngs = ScriptedRng(O, seed, w)
            k._rng = rngs
            k.mcusher._rng = rngs
            acc, prop, dh, snaps = record(k, rngs, occ0[w], nsteps, snap)
            key = f"met_{name}_{step_type}_w{w}"
            out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps})
            out[key + "_meta"] = np.array([seed, T])
    # Metropolis + TableFlip (swap_weight 0.2) on the 5-species rocksalt cell
    factory, occ0 = table_flip_model()
    for w in range(len(occ0)):
        seed = 700 + w
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")      # the constructor's trace-initialising step runs on an all-zero occupancy
            k = Metropolis(factory(), "table-flip", 4000.0, seed=seed, flip_table=TF_TABLE, swap_weight=0.2)
        assert type(k.mcusher).__name__ == "TableFlip"
        if w == 0:      # tables of the reference's usher (host-side mirror: smol_b200.sampler.table_flip_tables)
            u = k.mcusher
            out.update(tf_max_n=np.array(u.max_n), tf_d=np.array([u.d]), tf_weights=np.array(u.flip_weights),
                       tf_dim_ids_active=np.array(u._dim_ids_table), tf_table=np.array(u.flip_table),
                       tf_dim_ids=np.array([d for ids in u.dim_ids for d in ids]))
        rngs = TableFlipRng(O, seed, w)
        k._rng = rngs
        k.mcusher._rng = rngs
        k.mcusher._swapper._rng = rngs.swapper()
        acc, prop, dh, snaps = record(k, rngs, occ0[w], 400, 25)
        key = f"met_rs2of_tableflip_w{w}"
        out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps})
        out[key + "_meta"] = np.array([seed, 4000.0, 0.2])
    # biased Metropolis flips (kernel/bias.py) on the 5-species cell: the bias term of the exponent (metropolis.py:43-44)
    for tag, (bname, bkw) in BIAS_CASES.items():
        for w in range(len(occ0)):
            seed = 900 + w
            ens = factory()
            for sl in ens.sublattices:               # the reference's biases read Sublattice.site_space
                sl.site_space = _SiteSpace({spc: 1.0 / len(sl.species) for spc in sl.species})
            k = Metropolis(ens, "flip", 4000.0, seed=seed, bias_type=bname, bias_kwargs=dict(bkw))
            assert type(k.bias).__name__.lower() == bname.replace("-", "")
            if w == 0:      # the tables the reference's bias objects evaluate (host-side mirror: smol_b200/bias.py)
                b = k.bias
                if tag == "fugacity":
                    out["bias_fugacity_table"] = np.array(b._fu_table)
                elif tag == "charge":
                    out["bias_charge_table"] = np.array(b._c_table)
                else:
                    out["bias_hyperplane_dim_ids"] = np.array(b._dim_ids_table)
                out[f"bias_{tag}_values"] = np.array([b.compute_bias(o) for o in occ0])
            rngs = ScriptedRng(O, seed, w)
            k._rng = rngs
            k.mcusher._rng = rngs
            acc, prop, dh, snaps = record(k, rngs, occ0[w], 300, 25)
            key = f"met_rs2of_flip+{tag}_w{w}"
            out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps,
                        key + "_dbias": record.last_dbias.copy(), key + "_meta": np.array([seed, 4000.0])})
    # Composite (Flip with its own sublattice probabilities + Swap on the cation sublattice, weights 2:1) and MultiStep
    # (chains of Flip / Swap proposals) ushers, mcusher.py:203-394
    mcu = importlib.import_module("smol.moca.kernel.mcusher")
    for tag in USHER_CASES:
        for w in range(len(occ0)):
            seed = 1300 + w
            ens = factory()
            subl = ens.sublattices
            rngs = ScriptedRng(O, seed, w)
            if tag == "composite":
                subs = [mcu.Flip(subl, sublattice_probabilities=[0.7, 0.3]), mcu.Swap([subl[0]])]
                k = Metropolis(ens, "composite", 4000.0, seed=seed, mcushers=subs, mcusher_weights=[2, 1])
                for u in subs:
                    u._rng = rngs
            else:
                inner = (mcu.Flip if tag == "multistep-flip" else mcu.Swap)(subl)
                # (uniform length probabilities: passing step_probabilities to the reference raises AttributeError,
                #  mcusher.py:247 sets `step_p`, :274 reads `_step_p`)
                lens = [1, 2, 3] if tag == "multistep-flip" else [1, 2]
                k = Metropolis(ens, "multi-step", 4000.0, seed=seed, mcusher=inner, step_lengths=lens)
                inner._rng = ChainRng(rngs)
            assert type(k.mcusher).__name__ == ("Composite" if tag == "composite" else "MultiStep")
            k._rng = rngs
            k.mcusher._rng = PickRng(rngs)
            acc, prop, dh, snaps = record(k, rngs, occ0[w], 300, 25)
            key = f"met_rs2of_{tag}_w{w}"
            out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps,
                        key + "_meta": np.array([seed, 4000.0])})
    # Wang-Landau (flip) on the binary FCC cell
    factory, occ0 = mods["fcc3"]
    ens = factory()
    e0 = np.array([float(np.dot(ens.natural_parameters, ens.compute_feature_vector(o))) for o in occ0])
    lo, hi = float(e0.min() - 2.0), float(e0.max() + 2.0)
    bin_size = (hi - lo) / 23.7
    for w in range(len(occ0)):
        ens = factory()
        seed = 500 + w
        k = WangLandau(ens, "flip", lo, hi, bin_size, flatness=0.3, check_period=40, seed=seed)
        rngs = ScriptedRng(O, seed, w)
        k._rng = rngs
        k.mcusher._rng = rngs
        acc, prop, dh, snaps = record(k, rngs, occ0[w], 600, 50)
        key = f"wl_fcc3_flip_w{w}"
        out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps,
                    key + "_entropy": np.array(k._entropy), key + "_histogram": np.array(k._histogram),
                    key + "_occurrences": np.array(k._occurrences), key + "_mean_features": np.array(k._mean_features),
                    key + "_mod_factor": np.array([k._m])})
        out[key + "_meta"] = np.array([seed, lo, hi, bin_size, 0.3, 40])
    # Wang-Landau with canonical swaps, update_period 2 (entropy / histogram every second valid step, running mean
    # of the features instead of sums) and a modification factor divided by 3
    for w in range(len(occ0)):
        ens = factory()
        seed = 600 + w
        k = WangLandau(ens, "swap", lo, hi, bin_size, flatness=0.25, check_period=30, update_period=2, mod_factor=0.5,
                       mod_update=3.0, seed=seed)
        rngs = ScriptedRng(O, seed, w)
        k._rng = rngs
        k.mcusher._rng = rngs
        acc, prop, dh, snaps = record(k, rngs, occ0[w], 600, 50)
        key = f"wl2_fcc3_swap_w{w}"
        out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps,
                    key + "_entropy": np.array(k._entropy), key + "_histogram": np.array(k._histogram),
                    key + "_occurrences": np.array(k._occurrences), key + "_mean_features": np.array(k._mean_features),
                    key + "_mod_factor": np.array([k._m])})
        out[key + "_meta"] = np.array([seed, lo, hi, bin_size, 0.25, 30])
    # UniformlyRandom kernel (kernel/random.py): every proposal with log_priori >= 0 is accepted
    UniformlyRandom = importlib.import_module("smol.moca.kernel.random").UniformlyRandom
    for w in range(len(occ0)):
        seed = 650 + w
        k = UniformlyRandom(factory(), "swap", seed=seed)
        rngs = ScriptedRng(O, seed, w)
        k._rng = rngs
        k.mcusher._rng = rngs
        acc, prop, dh, snaps = record(k, rngs, occ0[w], 100, 25)
        key = f"uni_fcc3_swap_w{w}"
        out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps,
                    key + "_meta": np.array([seed, 0.0])})
    # MulticellMetropolis over three shapes of the 8-site FCC cell (uniform kernel probabilities: passing
    # kernel_probabilities to the reference raises AttributeError, base.py:505 sets `kernel_p`, :543 reads `_kernel_p`)
    MulticellMetropolis = importlib.import_module("smol.moca.kernel.metropolis").MulticellMetropolis
    shapes, ens_of, mc_occ0 = multicell_model()
    T, hop_periods, hop_p = 3000.0, [3, 5], [0.5, 0.5]
    for w in range(len(mc_occ0)):
        seed, kseeds = 11 + w, [1000 * (k + 1) + w for k in range(len(shapes))]
        kernels, rngs = [], []
        for k in range(len(shapes)):
            kk = Metropolis(ens_of(k), "swap", T, seed=kseeds[k])
            r = ScriptedRng(O, kseeds[k], w)
            kk._rng = r
            kk.mcusher._rng = r
            kernels.append(kk)
            rngs.append(r)
        mc = MulticellMetropolis(kernels, T, seed=seed, kernel_hop_periods=hop_periods, kernel_hop_probabilities=hop_p)

        class McRng:
            """shape / period choices from numpy as in the reference; a hop's uniform = Philox word 3 of the target"""
            def __init__(self):
                self.np = np.random.default_rng(seed)
                self.np.choice(np.array(hop_periods), p=hop_p)      # the draw of the constructor (base.py:532)
                self.t = 0

            def choice(self, a, p=None):
                return self.np.choice(a, p=p)

            def random(self):
                return O.u01(O.StepRandom(kseeds[int(mc.trace.kernel_index)], w, self.t).word(3))
        mrng = McRng()
        mc._rng = mrng
        mc.set_aux_state(np.array(mc_occ0[w], dtype=np.int32))
        occ = np.array(mc_occ0[w][0], dtype=np.int32)       # the sampler's live row (sampler.py:411-418)
        nst = 240
        idx, acc, occs = np.zeros(nst, dtype=np.int64), np.zeros(nst, dtype=bool), np.zeros((nst, len(occ)), dtype=np.int32)
        for t in range(nst):
            for r in rngs:
                r.begin_step(t)
            mrng.t = t
            tr = mc.single_step(occ)
            idx[t], acc[t], occs[t] = int(mc._current_kernel_index), bool(tr.accepted), occ
        key = f"mc_fcc8_swap_w{w}"
        out.update({key + "_idx": idx, key + "_acc": acc, key + "_occ": occs,
                    key + "_meta": np.array([seed, T, *kseeds])})
    # the reference's Sampler.run loop and SampleContainer accessors (sampler.py:164-297, container.py:131-382):
    # semigrand flips on the 5-species cell, two walkers, 360 steps thinned by 6
    Sampler = import_reference_sampler()
    factory, occ0 = table_flip_model()
    ens = factory()
    ens.thermo_boundaries = {}
    ens.num_energy_coefs = len(ens.natural_parameters) - 1          # the last natural parameter is the chemical work
    for sl in ens.sublattices:
        sl.site_space = _SiteSpace({spc: 1.0 / len(sl.species) for spc in sl.species})
    seeds = [41, 42]
    smp = Sampler.from_ensemble(ens, temperature=4000.0, step_type="flip", kernel_type="Metropolis", seeds=seeds,
                                nwalkers=2)
    for i, k in enumerate(smp.mckernels):
        r = AutoRng(O, seeds[i], i)
        k._rng = r
        k.mcusher._rng = r
    smp.run(360, occ0, thin_by=6, progress=False)
    c = smp.samples
    out["smp_occupancy"] = c.get_occupancies(flat=False)
    out["smp_features"] = c.get_feature_vectors(flat=False)
    out["smp_enthalpy"] = c.get_enthalpies(flat=False)
    out["smp_accepted"] = c.get_trace_value("accepted", flat=False)
    out["smp_temperature"] = c.get_trace_value("temperature", flat=False)
    out["smp_meta"] = np.array([*seeds, 4000.0, c.num_samples, c.total_mc_steps])
    for flat in (True, False):
        for name, v in container_answers(c, ens.sublattices, flat).items():
            out["smpq_" + name] = v
    # the reference's processors: ClusterExpansionProcessor / ClusterDecompositionProcessor (expansion.py, unmodified
    # Python on the compiled evaluators): full vectors, flip changes (1..3 sequential flips), properties
    from smol_b200 import lattice as L
    CE, CD, RefSubspace = import_reference_processors()
    for name, (sub, scm, coefs) in processor_cases().items():
        occs, flips = processor_flips(sub, scm, seed=3)
        rsub = RefSubspace(sub)
        it = L.cluster_interaction_tensors(sub, coefs)
        for tag, proc in (("ce", CE(rsub, scm, coefs)), ("cd", CD(rsub, scm, it))):
            key = f"proc_{name}_{tag}"
            out[key + "_full"] = np.array([proc.compute_feature_vector(o) for o in occs])
            out[key + "_delta"] = np.array([proc.compute_feature_vector_change(o, f) for o, f in zip(occs, flips)])
            out[key + "_prop"] = np.array([proc.compute_property(o) for o in occs])
            out[key + "_dprop"] = np.array([proc.compute_property_change(o, f) for o, f in zip(occs, flips)])
            out[key + "_meta"] = np.array([proc.size, proc.num_sites])
    # the reference's Ensemble (ensemble.py, unmodified: chemical-potential table, chemical work, natural parameters)
    # over its ClusterDecompositionProcessor inside its Metropolis kernel: semigrand flips on the 5-species cell -- every
    # layer between the lattice tables and the random words is the reference's code
    cm = types.ModuleType("pymatgen.core.composition")
    cm.ChemicalPotential = dict
    sys.modules["pymatgen.core.composition"] = cm
    pp = sys.modules["smol.moca.processor"]
    pp.CompositeProcessor = importlib.import_module("smol.moca.processor.composite").CompositeProcessor
    pp.EwaldProcessor = type("EwaldProcessor", (), {})
    RefEnsemble = importlib.import_module("smol.moca.ensemble").Ensemble
    factory, occ0 = table_flip_model()
    sub, scm, _ = processor_cases()["rs2of"]
    for w in range(len(occ0)):
        o_ens = factory()
        for sl in o_ens.sublattices:
            sl.site_space = _SiteSpace({spc: 1.0 / len(sl.species) for spc in sl.species})
        proc = CD(RefSubspace(sub), scm, TF_INTERACTIONS())
        ens = RefEnsemble(proc, sublattices=o_ens.sublattices, chemical_potentials=dict(TF_MUS))
        assert len(ens.natural_parameters) == len(o_ens.natural_parameters)
        if w == 0:      # the reference ensemble's chemical-potential table (host-side mirror: smol_b200.Ensemble.mu_table)
            out["ens_mu_table"] = np.array(ens._chemical_potentials["table"])
            out["ens_natural_parameters"] = np.array(ens.natural_parameters)
            out["ens_num_energy_coefs"] = np.array([ens.num_energy_coefs])
        seed = 1500 + w
        k = Metropolis(ens, "flip", 4000.0, seed=seed)
        rngs = ScriptedRng(O, seed, w)
        k._rng = rngs
        k.mcusher._rng = rngs
        acc, prop, dh, snaps = record(k, rngs, occ0[w], 300, 25)
        key = f"fullref_rs2of_flip_w{w}"
        out.update({key + "_acc": acc, key + "_prop": prop, key + "_dh": dh, key + "_snaps": snaps,
                    key + "_meta": np.array([seed, 4000.0]),
                    key + "_feat0": np.array(ens.compute_feature_vector(occ0[w])),
                    key + "_natural": np.array(ens.natural_parameters)})
    # the reference's EwaldProcessor and CompositeProcessor (processor/ewald.py, composite.py, unmodified; the Ewald
    # MATRIX is an input here -- pymatgen computes it in the reference -- so EwaldTerm's two structure-dependent
    # methods are overridden to hand over this build's matrix / index layout; get_ewald_occu, the masked matrix sum,
    # the sequential flip loop over the compiled delta_ewald_single_flip and the feature concatenation are the
    # reference's)
    pa = types.ModuleType("pymatgen.analysis")
    pe = types.ModuleType("pymatgen.analysis.ewald")
    pe.EwaldSummation = lambda *a, **k: None
    sys.modules.update({"pymatgen.analysis": pa, "pymatgen.analysis.ewald": pe})
    ext = types.ModuleType("smol.cofe.extern")
    ext.__path__ = [REF + "/cofe/extern"]
    sys.modules["smol.cofe.extern"] = ext
    sys.modules["smol.cofe.space.domain"].get_allowed_species = sys.modules["smol.cofe.space"].get_allowed_species
    RealEwaldTerm = importlib.import_module("smol.cofe.extern.ewald").EwaldTerm
    RefEwaldProcessor = importlib.import_module("smol.moca.processor.ewald").EwaldProcessor
    RefComposite = importlib.import_module("smol.moca.processor.composite").CompositeProcessor
    sub, scm, coefs = processor_cases()["rs2of"]
    occs, flips = processor_flips(sub, scm, seed=3)
    ewm, ewi = L.ewald_matrix(sub, scm)

    class GivenMatrixTerm(RealEwaldTerm):
        def get_ewald_structure(self, structure):
            return None, ewi

        def get_ewald_matrix(self, ewald_summation):
            return ewm
    term = GivenMatrixTerm()
    rsub = RefSubspace(sub)
    rsub.external_terms = (term,)
    rsub.add_external_term = lambda t: None
    ew = RefEwaldProcessor(rsub, scm, term, coefficient=0.1)
    comp = RefComposite(rsub, scm)
    comp.add_processor(CD(rsub, scm, L.cluster_interaction_tensors(sub, coefs)))
    comp.add_processor(ew)
    for tag, proc in (("ewald", ew), ("composite", comp)):
        key = f"proc_rs2of_{tag}"
        out[key + "_full"] = np.array([np.atleast_1d(proc.compute_feature_vector(o)) for o in occs])
        out[key + "_delta"] = np.array([np.atleast_1d(proc.compute_feature_vector_change(o, f)) for o, f in zip(occs, flips)])
        out[key + "_prop"] = np.array([float(np.sum(proc.compute_property(o))) for o in occs])
        out[key + "_dprop"] = np.array([float(np.sum(proc.compute_property_change(o, f))) for o, f in zip(occs, flips)])
        out[key + "_coefs"] = np.atleast_1d(np.array(proc.coefs, dtype=np.float64))
    # the distance processors behind smol's SQS generation (processor/distance.py, unmodified): features
    # [L, |f_i - target_i| ...]; the target is the vector of the first occupancy, so part of the features match exactly
    dist = importlib.import_module("smol.moca.processor.distance")
    for name in ("fcc3", "rs2of"):
        sub, scm, coefs = processor_cases()[name]
        occs, flips = processor_flips(sub, scm, seed=3)
        rsub = RefSubspace(sub)
        it = L.cluster_interaction_tensors(sub, coefs)
        size = sub.supercell_size(scm)
        targets = {"corr": out[f"proc_{name}_ce_full"][0] / size, "int": out[f"proc_{name}_cd_full"][0] / size}
        procs = {"corr": dist.CorrelationDistanceProcessor(rsub, scm, target_vector=targets["corr"], match_weight=0.7,
                                                          match_tol=DIST_TOL),
                 "int": dist.ClusterInteractionDistanceProcessor(rsub, scm, interaction_tensors=it,
                                                                 target_vector=targets["int"], match_weight=0.7,
                                                                 match_tol=DIST_TOL_INT)}
        for tag, proc in procs.items():
            key = f"dist_{name}_{tag}"
            out[key + "_target"] = targets[tag]
            out[key + "_full"] = np.array([proc.compute_feature_vector(o) for o in occs])
            out[key + "_delta"] = np.array([proc.compute_feature_vector_change(o, f) for o, f in zip(occs, flips)])
            out[key + "_coefs"] = np.array(proc.coefs)
    # the SQS sampling kernel as smol runs it (capp/generate/special/sqs.py:523-540): MulticellMetropolis over
    # Metropolis kernels whose ensembles wrap CorrelationDistanceProcessors of different supercell shapes -- all of it
    # the reference's classes (ensemble, distance processors on the compiled evaluators, kernels, ushers)
    from tests import models as M
    fsub = M.fcc_subspace()
    target = np.zeros(fsub.num_corr_functions)
    target[0] = 1.0                                   # random 50/50 alloy in the sinusoid basis
    for w in range(len(mc_occ0)):
        seed, kseeds = 21 + w, [4000 * (k + 1) + w for k in range(len(MC_SHAPES))]
        kernels, rngs = [], []
        for k, shape in enumerate(MC_SHAPES):
            proc = dist.CorrelationDistanceProcessor(RefSubspace(fsub), shape, target_vector=target, match_weight=SQS_W,
                                                     match_tol=SQS_TOL)
            subl = [O.Sublattice(("A", "B"), np.arange(8))]
            for sl in subl:
                sl.site_space = _SiteSpace({spc: 0.5 for spc in sl.species})
            kk = Metropolis(RefEnsemble(proc, sublattices=subl), "swap", SQS_T, seed=kseeds[k])
            r = ScriptedRng(O, kseeds[k], w)
            kk._rng = r
            kk.mcusher._rng = r
            kernels.append(kk)
            rngs.append(r)
        mc = MulticellMetropolis(kernels, SQS_T, seed=seed, kernel_hop_periods=[2, 4], kernel_hop_probabilities=[0.5, 0.5])

        class McRng2:
            def __init__(self):
                self.np = np.random.default_rng(seed)
                self.np.choice(np.array([2, 4]), p=[0.5, 0.5])
                self.t = 0

            def choice(self, a, p=None):
                return self.np.choice(a, p=p)

            def random(self):
                return O.u01(O.StepRandom(kseeds[int(mc.trace.kernel_index)], w, self.t).word(3))
        mrng = McRng2()
        mc._rng = mrng
        mc.set_aux_state(np.array(mc_occ0[w], dtype=np.int32))
        occ = np.array(mc_occ0[w][0], dtype=np.int32)
        nst = 200
        idx, acc, occs_ = np.zeros(nst, dtype=np.int64), np.zeros(nst, dtype=bool), np.zeros((nst, 8), dtype=np.int32)
        for t in range(nst):
            for r in rngs:
                r.begin_step(t)
            mrng.t = t
            tr = mc.single_step(occ)
            idx[t], acc[t], occs_[t] = int(mc._current_kernel_index), bool(tr.accepted), occ
        key = f"mcdist_fcc8_swap_w{w}"
        out.update({key + "_idx": idx, key + "_acc": acc, key + "_occ": occs_, key + "_meta": np.array([seed, SQS_T, *kseeds]),
                    key + "_features": np.array(mc._features[mc._current_kernel_index])})
    # site bases (cofe/space/basis.py, unmodified): function arrays of the sinusoid and indicator bases, raw and
    # orthonormalised, uniform and concentration measures -- what every correlation tensor is built from
    from collections import OrderedDict
    sys.modules["smol.cofe.space.domain"].SiteSpace = type("SiteSpace", (), {})
    basis = importlib.import_module("smol.cofe.space.basis")
    for name in ("sinusoid", "indicator"):
        for n in (2, 3, 4, 5):
            for mtag, measure in (("uniform", np.full(n, 1.0 / n)), ("conc", np.arange(1, n + 1) / (n * (n + 1) / 2))):
                space = OrderedDict((f"S{i}", float(measure[i])) for i in range(n))
                b = basis.basis_factory(name, space)
                out[f"basis_{name}_{n}_{mtag}_raw"] = np.array(b.function_array)
                b.orthonormalize()
                assert b.is_orthonormal
                out[f"basis_{name}_{n}_{mtag}_orth"] = np.array(b.function_array)
    path = os.path.join(HERE, "ref_python_steps.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in list(out.items())[:6]})
    for k in out:
        if k.endswith("_acc"):
            print(k, "accepted", int(out[k].sum()), "of", len(out[k]))


if __name__ == "__main__":
    main()
