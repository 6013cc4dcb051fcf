"""The BASELINE.json configurations as ONE definition each, shared by
``bench.py --config``, the full-size parity tests and the profiling scripts: the SAME tables feed
the product (``smol_b200``) and the checker (``oracle/``).

Nothing here is product code; the oracle side is only touched by ``oracle_kernels`` (tests and the
CPU legs of ``bench.py``).
"""
from __future__ import annotations

import functools
import os

import numpy as np

from smol_b200 import lattice as L
from tests import models as M


class Workload:
    """One BASELINE config: model tables, sampler arguments, initial occupancies, byte models."""

    id = 0
    name = ""
    usher = "swap"
    kernel = "Metropolis"
    walkers_per_gpu = 4096
    temperature = None
    thin_by = 1
    samples_per_bench_step = 1      # sampling intervals advanced by one bench "step" (one lmc_run launch)
    record_occupancy = True
    # bytes per attempted step of the REFERENCE algorithm (int8 storage)
    algorithmic_bytes = 0.0

    # -- tables ---------------------------------------------------------------------------------
    def subspace(self):
        raise NotImplementedError

    def supercell(self):
        raise NotImplementedError

    @property
    def num_sites(self):
        return len(self.subspace().allowed_species(self.supercell()))

    def chemical_potentials(self):
        return None

    def interaction_tensors(self):
        raise NotImplementedError

    def ewald(self):
        """(matrix, inds, coefficient) or None"""
        return None

    # -- product side ---------------------------------------------------------------------------
    def product_ensemble(self):
        import smol_b200 as S
        sub, scm = self.subspace(), self.supercell()
        ce = S.ClusterDecompositionProcessor(sub, scm, self.interaction_tensors())
        ew = self.ewald()
        if ew is None:
            proc = ce
        else:
            proc = S.CompositeProcessor(sub, scm)
            proc.add_processor(ce)
            proc.add_processor(S.EwaldProcessor(sub, scm, coefficient=ew[2], ewald_matrix=ew[0], ewald_inds=ew[1]))
        return S.Ensemble(proc, chemical_potentials=self.chemical_potentials())

    def sampler_args(self, ens=None):
        """positional args of Sampler.from_ensemble after the ensemble"""
        return (self.temperature,)

    def sampler_kwargs(self):
        return {}

    def sampler(self, ens, W, seeds, walker_id_base=0, **kw):
        import smol_b200 as S
        kwargs = dict(step_type=self.usher, kernel_type=self.kernel, nwalkers=W, seeds=list(seeds),
                      walker_id_base=walker_id_base, record_occupancy=self.record_occupancy)
        kwargs.update(self.sampler_kwargs())
        kwargs.update(kw)
        return S.Sampler.from_ensemble(ens, *self.sampler_args(ens), **kwargs)

    def initial_occupancies(self, W, seed=0):
        raise NotImplementedError

    # -- checker side ---------------------------------------------------------------------------
    def oracle_processor(self, use_ref=False):
        from oracle import lmc_oracle as O
        sub, scm = self.subspace(), self.supercell()
        ce = O.ClusterDecompositionProcessor(sub, scm, self.interaction_tensors(), use_ref=use_ref)
        ew = self.ewald()
        if ew is None:
            return ce
        return O.CompositeProcessor([ce, O.EwaldProcessor(ew[0], ew[1], ew[2], use_ref=use_ref)])

    def oracle_sublattices(self, ens=None):
        """oracle Sublattice objects equal to the product ensemble's (built without a device)"""
        from oracle import lmc_oracle as O
        import smol_b200 as S
        if ens is None:    # host-only: no engine is created for the sublattice list
            ens = S.Ensemble(S.ClusterDecompositionProcessor(self.subspace(), self.supercell(),
                                                             self.interaction_tensors()))
        return M.oracle_sublattices(O, ens.sublattices)

    def oracle_usher(self, subl):
        from oracle import lmc_oracle as O
        return {"swap": O.Swap, "flip": O.Flip}[self.usher](subl)

    def oracle_kernels(self, seeds, walker0=0, use_ref=False):
        from oracle import lmc_oracle as O
        subl = self.oracle_sublattices()
        ens = O.Ensemble(self.oracle_processor(use_ref), subl, chemical_potentials=self.chemical_potentials())
        return [O.Metropolis(ens, self.oracle_usher(subl), self.temperature, seed=int(s), walker=walker0 + w)
                for w, s in enumerate(seeds)]

    # -- measurement ----------------------------------------------------------------------------
    # True: the kernel moves the bytes of SURVEY 8(d)'s model (roofline.frac is quoted on them);
    # False: the built algorithm replaces the reference's (frac is quoted on built_bytes)
    roofline_uses_survey_bytes = True

    def built_bytes(self, acceptance, ewald_cache=False):
        """(bytes per attempted step the BUILT algorithm moves per walker, description) -- same accounting as
        SURVEY 8(d): per-walker state and per-flip gathers count, tables shared by all walkers do not"""
        return self.algorithmic_bytes, "the reference algorithm's (SURVEY 8d)"


# ------------------------------------------------------------------------------------------------
class Config2(Workload):
    """binary FCC 8x8x8, canonical Metropolis swap, T = 1000 K, 4096 walkers, one sample per sweep"""
    id = 2
    n_cell = 8
    usher = "swap"
    temperature = 1000.0
    thin_by = 512
    samples_per_bench_step = 8
    algorithmic_bytes = 430.0   # 426 int8 gathers + 2 writes + trace / thin_by
    name = "binary FCC 8x8x8 (512 sites), S_fcc clusters, canonical Metropolis swap, T=1000K, %d walkers/GPU, thin_by=512"

    def subspace(self):
        return M.fcc_subspace()

    def supercell(self):
        return np.eye(3, dtype=int) * self.n_cell

    @functools.lru_cache(maxsize=None)
    def interaction_tensors(self):
        sub = self.subspace()
        return L.cluster_interaction_tensors(sub, M.fcc_coefs(sub))

    def initial_occupancies(self, W, seed=0):
        return M.random_occupancies(self.subspace(), self.supercell(), W, seed=seed, balanced=True)

    def built_bytes(self, acceptance, ewald_cache=False):
        # speculative kernel over compact environment words: per attempted swap the two sites' 64-bit words + the pair
        # mask + the 2 occupancy bytes of the proposal; an accepted swap is re-evaluated with the classic records (213
        # gathers per flip), writes 2 bytes and xors one bit into the words of the ~54 sites around each changed site
        b = 2 * 8 + 8 + 2 + acceptance * (2 * 213 + 2 + 2 * 64 * 8) + (512 + 64 + 9) / 512
        return b, ("compact environment words: 2 x 8 B site words + 8 B pair mask + 2 occupancy bytes per attempted swap + "
                   "(2 x 213 gathers + 2 writes + 2 x 64 word updates) per ACCEPTED swap + trace / thin_by")


class Config3(Workload):
    """ternary rocksalt 8x8x8 + Ewald, semigrand flips, T = 1500 K, 4096 walkers"""
    id = 3
    usher = "flip"
    temperature = 1500.0
    thin_by = 512
    samples_per_bench_step = 8
    algorithmic_bytes = 214 + 2 * 1024 * 8 + 1024.0
    name = ("ternary rocksalt 8x8x8 (1024 sites, 512 active) + Ewald (E=2048), cluster decomposition, semigrand flip, "
            "T=1500K, %d walkers/GPU, thin_by=512")

    def subspace(self):
        return M.rocksalt_subspace()

    def supercell(self):
        return np.eye(3, dtype=int) * 8

    def chemical_potentials(self):
        return {"Li+": 0.0, "Mn3+": 0.3, "Ti4+": -0.2}

    @functools.lru_cache(maxsize=None)
    def interaction_tensors(self):
        sub = self.subspace()
        return L.cluster_interaction_tensors(sub, np.random.default_rng(3).normal(0, 0.02, sub.num_corr_functions))

    @functools.lru_cache(maxsize=None)
    def ewald(self):
        ewm, ewi = L.ewald_matrix(self.subspace(), self.supercell())
        return ewm, ewi, 0.1          # coefficient 1 / eps, eps = 10

    def initial_occupancies(self, W, seed=5):
        return M.random_occupancies(self.subspace(), self.supercell(), W, seed=seed)

    roofline_uses_survey_bytes = False

    def built_bytes(self, acceptance, ewald_cache=True):
        # speculative kernel + Ewald potential cache: 66 gathers + 8 B cached potential per attempted flip; an
        # accepted flip re-evaluates (213 gathers), writes 1 byte, reads one row of the site kernel K (N f64) and
        # read-modify-writes the walker's potential row (2 x N f64)
        N = 1024
        b = 66 + 8 + acceptance * (213 + 1 + 2 * 8 * N) + (N + 80 + 9) / 512
        return b, ("66 int8 gathers + 1 cached potential (8 B) per attempted flip + (213 gathers + 1 write + potential "
                   "row RMW 16N) per ACCEPTED flip + trace / thin_by; the rows of the site kernel K are a table shared by "
                   "all walkers (L2 resident, not counted: SURVEY 8d accounting); the reference's two E-long matrix rows "
                   "per flip are never read")


class Config4(Workload):
    """binary FCC 8x8x8 Wang-Landau, AFM-Ising-like coefficients of the wang-landau notebook, 1024 walkers"""
    id = 4
    usher = "flip"
    kernel = "WangLandau"
    walkers_per_gpu = 1024
    thin_by = 512
    samples_per_bench_step = 40
    record_occupancy = False
    bin_size = 4.0
    flatness = 0.8
    check_period = 1000
    algorithmic_bytes = 214 + 16 + 48 + 8 * 16.0
    name = ("binary FCC 8x8x8 Wang-Landau flip (h=2, J=1 AFM Ising coefficients, bin 4.0, flatness 0.8, "
            "check_period 1000), %d independent walkers/GPU, thin_by=512")

    def subspace(self):
        return M.fcc_subspace()

    def supercell(self):
        return np.eye(3, dtype=int) * 8

    @functools.lru_cache(maxsize=None)
    def interaction_tensors(self):
        sub = self.subspace()
        coefs = np.zeros(sub.num_corr_functions)
        mult = sub.function_total_multiplicities
        coefs[1], coefs[2] = 2.0 * mult[1], 1.0 * mult[2]     # wang-landau-ising.ipynb:76-80, kB := 1
        return L.cluster_interaction_tensors(sub, coefs)

    def initial_occupancies(self, W, seed=2):
        return M.random_occupancies(self.subspace(), self.supercell(), W, seed=seed)

    def window(self):
        """(min_enthalpy, max_enthalpy) = mean -+ (5 sigma + 200) of the enthalpies of the 64 random occupancies
        ``initial_occupancies(64)``, levels centred on the lattice energies (multiples of bin_size): 298 bins.
        Constants (so that both bench arms use the same window without evaluating anything);
        ``tests/test_gpu_scale.py::test_config4_*`` re-derives them on the GPU."""
        return -586.0, 604.9333891803888

    @staticmethod
    def window_from_enthalpies(e0, bin_size=4.0):
        lo = float(np.floor((e0.mean() - 5 * e0.std() - 200) / bin_size) * bin_size - 2.0)
        return lo, float(e0.mean() + 5 * e0.std() + 200)

    def sampler_args(self, ens=None):
        lo, hi = self.window()
        return (lo, hi, self.bin_size)

    def sampler_kwargs(self):
        return dict(flatness=self.flatness, check_period=self.check_period, wl_trace="none")

    def oracle_kernels(self, seeds, walker0=0, use_ref=False):
        from oracle import lmc_oracle as O
        subl = self.oracle_sublattices()
        ens = O.Ensemble(self.oracle_processor(use_ref), subl)
        lo, hi = self.window()
        return [O.WangLandau(ens, O.Flip(subl), lo, hi, self.bin_size, flatness=self.flatness,
                             check_period=self.check_period, seed=int(s), walker=walker0 + w)
                for w, s in enumerate(seeds)]


class Config5(Workload):
    """5-species rocksalt 12x12x12 + Ewald, charge-neutral semigrand table flips, T = 1500 K"""
    id = 5
    n_cell = 12
    usher = "table_flip"
    temperature = 1500.0
    samples_per_bench_step = 1
    record_occupancy = True
    flip_table = ((-1, 1, 0, 2, -2), (0, -1, 1, 1, -1))
    swap_weight = 0.1
    name = ("5-species rocksalt %dx%dx%d (N=%d, Ewald E=%d) + Ewald, charge-neutral table flips (swap_weight 0.1), "
            "T=1500K, %%d walkers/GPU, thin_by=%d")

    def __init__(self, n_cell=12):
        self.n_cell = n_cell
        nc = n_cell ** 3
        self.thin_by = 2 * nc
        self.algorithmic_bytes = 2.5 * (250 + 2 * 2 * nc * 8 + 2 * nc)
        self.name = self.name % (n_cell, n_cell, n_cell, 2 * nc, 5 * nc, self.thin_by)

    def subspace(self):
        return M.rocksalt_subspace(anions=("O2-", "F-"))

    def supercell(self):
        return np.eye(3, dtype=int) * self.n_cell

    def chemical_potentials(self):
        return {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0}

    @functools.lru_cache(maxsize=None)
    def interaction_tensors(self):
        sub = self.subspace()
        return L.cluster_interaction_tensors(sub, np.random.default_rng(21).normal(0, 0.01, sub.num_corr_functions))

    @functools.lru_cache(maxsize=None)
    def ewald(self):
        """The E x E matrix (597 MB at 12x12x12) is built once per process; a cache file lets the CPU-arm worker
        processes map it instead of rebuilding it (LMC_EWALD_CACHE_DIR, default /tmp)."""
        path = os.path.join(os.environ.get("LMC_EWALD_CACHE_DIR", "/tmp"), "lmc_ewald_cfg5_n%d.npz" % self.n_cell)
        ipath = path.replace(".npz", "_m.npy")
        if os.path.exists(path) and os.path.exists(ipath):
            ewi = np.load(path)["ewi"]
            return np.load(ipath, mmap_mode="r"), ewi, 0.1
        ewm, ewi = L.ewald_matrix(self.subspace(), self.supercell())
        return ewm, ewi, 0.1

    def cache_ewald(self):
        """write the matrix where ``ewald`` of other processes finds it"""
        path = os.path.join(os.environ.get("LMC_EWALD_CACHE_DIR", "/tmp"), "lmc_ewald_cfg5_n%d.npz" % self.n_cell)
        ipath = path.replace(".npz", "_m.npy")
        if not (os.path.exists(path) and os.path.exists(ipath)):
            ewm, ewi, _ = self.ewald()
            np.save(ipath + ".tmp.npy", ewm)
            os.replace(ipath + ".tmp.npy", ipath)
            np.savez(path, ewi=ewi)

    def sampler_kwargs(self):
        return dict(flip_table=[list(r) for r in self.flip_table], swap_weight=self.swap_weight)

    def initial_occupancies(self, W, seed=21):
        """charge neutral: nLi + 3 nMn + 4 nTi = 2 nO + nF (capp/generate/random.py:88-144 semantics)"""
        rng = np.random.default_rng(seed)
        nc = self.n_cell ** 3
        nMn, nTi = nc // 6, nc // 8
        nLi = nc - nMn - nTi
        nO = nLi + 3 * nMn + 4 * nTi - nc
        cat = np.array([0] * nLi + [1] * nMn + [2] * nTi)
        ani = np.array([0] * nO + [1] * (nc - nO))
        occ0 = np.zeros((W, 2 * nc), dtype=np.int32)
        for w in range(W):
            occ0[w, :nc], occ0[w, nc:] = rng.permutation(cat), rng.permutation(ani)
        return occ0

    roofline_uses_survey_bytes = False

    def built_bytes(self, acceptance, ewald_cache=False):
        N = 2 * self.n_cell ** 3
        if ewald_cache:
            # speculative table-flip kernel + potential cache (while fewer than ~1/3 of the steps are accepted): per
            # changed site 64 merged records x 3 int8 gathers + the cached potential (8 B) + one element of K per earlier
            # flip of the step; an ACCEPTED step is re-evaluated with the classic records (~241 gathers per changed
            # site), reads one row of K per changed site and read-modify-writes the walker's potential row
            b = 2.5 * (192 + 8 + 8 + 1) + acceptance * (2.5 * 241 + 16 * N) + (N + 88 + 9) / self.thin_by
            return b, ("speculative table-flip kernel, potential cache: per changed site 192 int8 gathers (64 merged "
                       "records) + 16 B of cached potential / K elements; per ACCEPTED step 2.5 x 241 gathers + the "
                       "walker's potential row read-modify-write (16N B); the rows of the site kernel K (2.5 x 8N B per "
                       "accepted step) are a table shared by all walkers, L2 resident, not counted (SURVEY 8d accounting)")
        # Ewald through the factorised site kernel: per changed site ~241 gathers + ONE row of K (N f64) + the
        # walker's charge indices (N bytes); k ~ 2.5 changed sites per step
        b = 2.5 * (241 + 8 * N + N + 1) + (N + 88 + 9) / self.thin_by
        return b, ("per changed site: ~241 int8 gathers + one row of the site kernel K (8N B, instead of the "
                   "reference's two E-matrix rows) + the walker's charge indices (N B); 2.5 changed sites per step")

    def oracle_usher(self, subl):
        from oracle import lmc_oracle as O
        return O.TableFlip(subl, [list(r) for r in self.flip_table], swap_weight=self.swap_weight)


CONFIGS = {2: Config2, 3: Config3, 4: Config4, 5: Config5}


@functools.lru_cache(maxsize=None)
def get(config_id: int, **kw):
    return CONFIGS[int(config_id)](**kw)
