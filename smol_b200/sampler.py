"""Sampler: drop-in for ``smol.moca.Sampler`` whose step loop runs in one CUDA kernel.

Mirrors ``smol/moca/sampler/sampler.py``: ``Sampler.from_ensemble(ensemble, *args,
step_type, kernel_type, seeds, nwalkers, **kwargs)``, ``run``, ``anneal``, ``samples``,
``efficiency``, ``clear_samples``, ``mckernels[i].temperature``.  The whole body of
``Sampler.sample`` (sampler.py:195-210: steps x walkers x single_step) is ``lmc_run``.

Randomness: counter-based Philox4x32-10 keyed by the walker's seed, counter = (global step,
block, global walker id) -- independent of how walkers are sharded over GPUs.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from . import _capi as capi
from .container import SampleContainer

kB = 8.617333262145e-5  # smol/constants.py:4

_POOL = None


def _fast_copy(src: np.ndarray) -> np.ndarray:
    """Copy out of the page-locked staging buffer; large arrays are split over a few threads
    (numpy releases the GIL in copyto, and the page faults of the fresh destination parallelise)."""
    global _POOL
    dst = np.empty_like(src)
    if src.nbytes < (4 << 20) or src.shape[0] < 2:
        np.copyto(dst, src)
        return dst
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=4)
    parts = min(4, src.shape[0])
    bounds = np.linspace(0, src.shape[0], parts + 1).astype(int)
    list(_POOL.map(lambda ab: np.copyto(dst[ab[0]:ab[1]], src[ab[0]:ab[1]]), zip(bounds[:-1], bounds[1:])))
    return dst



class _PinnedPool:
    """Recycles page-locked host blocks that receive the per-chunk trace copies.

    The trace arrays handed to the SampleContainer are views of these blocks (no staging copy on
    the way out); ``SampleContainer.clear`` returns a block once no outside view refers to it.
    ``LMC_PINNED_POOL_MB`` caps the page-locked memory in flight (default 4096); beyond the cap the
    sampler falls back to a fixed staging slot plus a host copy."""

    def __init__(self):
        self.free = {}
        self.outstanding = 0
        self.cap = int(os.environ.get("LMC_PINNED_POOL_MB", "4096")) << 20

    def acquire(self, shape, dtype):
        import torch
        key = (tuple(shape), dtype)
        lst = self.free.get(key)
        if lst:
            t = lst.pop()
        else:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            if self.outstanding + nbytes > self.cap:
                return None
            t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        self.outstanding += t.numel() * t.element_size()
        return t

    def release(self, t, recycle=True):
        """``recycle=False``: the block is still viewed by a caller (it lives on until that view dies), it only
        stops counting against the budget"""
        self.outstanding -= t.numel() * t.element_size()
        if not recycle:
            return
        lst = self.free.setdefault((tuple(t.shape), t.dtype), [])
        if len(lst) < 8:
            lst.append(t)


_PINNED = _PinnedPool()

_USHERS = {"flip": capi.LMC_USHER_FLIP, "swap": capi.LMC_USHER_SWAP,
           "tableflip": capi.LMC_USHER_TABLEFLIP, "table_flip": capi.LMC_USHER_TABLEFLIP,
           "composite": capi.LMC_USHER_COMPOSITE, "multistep": capi.LMC_USHER_MULTISTEP}
_KERNELS = {"metropolis": capi.LMC_KERNEL_METROPOLIS, "uniformlyrandom": capi.LMC_KERNEL_METROPOLIS,
            "wanglandau": capi.LMC_KERNEL_WANGLANDAU, "wang-landau": capi.LMC_KERNEL_WANGLANDAU}


class _KernelView:
    """Per-walker handle standing in for the reference's MCKernel objects
    (``sampler.mckernels[i].temperature = T`` is how ``anneal`` drives them, sampler.py:357-371)."""

    def __init__(self, sampler, index):
        self._s, self._i = sampler, index

    @property
    def temperature(self):
        return float(self._s._temperature[self._i])

    @temperature.setter
    def temperature(self, value):
        self._s._temperature[self._i] = float(value)

    @property
    def beta(self):
        return 1.0 / (self._s.kB * self.temperature)

    @property
    def seed(self):
        return int(self._s.seeds[self._i])

    @property
    def spec(self):
        return {"kernel": self._s.kernel_type, "step": self._s.step_type, "seed": self.seed}


def table_flip_tables(sublattices, flip_table, flip_weights=None, swap_weight=0.1):
    """TableFlip.__init__ tables (mcusher.py:486-550): dims over ALL sublattices in order."""
    flip_table = np.array(flip_table, dtype=np.int32)
    active = [s for s in sublattices if len(s.active_sites) > 0]
    dim_sl, dim_code, max_n = [], [], []
    for s in sublattices:
        a = next((i for i, t in enumerate(active) if t is s), -1)
        for code in s.encoding:
            dim_sl.append(a)
            dim_code.append(int(code))
            max_n.append(len(s.active_sites))
    d = len(dim_sl)
    if flip_table.ndim != 2 or flip_table.shape[1] != d:
        raise ValueError(f"flip_table must have shape [n_flips, {d}]")
    if flip_weights is None:
        weights = np.ones(2 * len(flip_table))
    elif len(flip_weights) == len(flip_table):
        weights = np.repeat(np.asarray(flip_weights, dtype=np.float64), 2)
    elif len(flip_weights) == 2 * len(flip_table):
        weights = np.asarray(flip_weights, dtype=np.float64)
    else:
        raise ValueError(f"{len(flip_weights)} weights provided. You must provide either 1* or "
                         f"2* weights given {len(flip_table)} flip vectors!")
    changed = np.abs(flip_table).clip(min=0)
    if (np.maximum(flip_table, 0).sum(axis=1) > capi.LMC_MAX_FLIPS).any() or \
            (np.maximum(-flip_table, 0).sum(axis=1) > capi.LMC_MAX_FLIPS).any():
        raise ValueError("flip vectors changing more than 4 sites are not supported")
    del changed
    return dict(num_dims=d, table=flip_table, weights=weights, max_n=max_n, dim_sl=dim_sl,
                dim_code=dim_code, swap_weight=float(swap_weight))


class Sampler:
    """GPU Monte-Carlo sampler with the interface of ``smol.moca.Sampler``."""

    def __init__(self, ensemble, kernel_type="Metropolis", step_type="swap", nwalkers=1, seeds=None,
                 temperature=None, wl_params=None, usher_kwargs=None, walker_id_base=0,
                 group_size=0, block_threads=0, record_occupancy=True, device=None, kB_=kB,
                 spec_mode=0, ewald_field="auto", bias_type=None, bias_kwargs=None, wl_trace="full", spec_env=False):
        from .engine import LmcEngine
        self.ensemble = ensemble
        self.kernel_type = kernel_type
        self.step_type = step_type
        key = kernel_type.lower().replace("_", "")
        if key not in _KERNELS:
            raise ValueError(f"{kernel_type} is not a supported MCKernel "
                             f"(available: Metropolis, UniformlyRandom, WangLandau)")
        self._kernel = _KERNELS[key]
        self._uniform = key == "uniformlyrandom"
        ukey = step_type.lower().replace("-", "").replace("_", "") if step_type else "swap"
        ukey = "tableflip" if ukey == "tableflip" else ukey
        if ukey not in _USHERS:
            raise ValueError(f"{step_type} is not a supported MCUsher (available: flip, swap, "
                             f"table_flip, composite, multi_step)")
        self._usher = _USHERS[ukey]
        usher_kwargs = dict(usher_kwargs or {})
        self._composite = None
        if self._usher == capi.LMC_USHER_COMPOSITE:
            # Composite(sublattices, mcushers, mcusher_weights), mcusher.py:314-350
            from .usher import Composite
            comp = Composite(ensemble.sublattices, usher_kwargs.pop("mcushers", None),
                             usher_kwargs.pop("mcusher_weights", None))
            if not comp.mcushers:
                raise ValueError("a composite usher needs at least one mcusher")
            self._composite = comp.device_tables(ensemble.sublattices)
            self.mcusher = comp
        self._multistep = None
        if self._usher == capi.LMC_USHER_MULTISTEP:
            # MultiStep(sublattices, mcusher, step_lengths, step_probabilities), mcusher.py:206-272
            from .usher import MultiStep
            if "mcusher" not in usher_kwargs or "step_lengths" not in usher_kwargs:
                raise TypeError("MultiStep needs mcusher and step_lengths")
            ms = MultiStep(ensemble.sublattices, usher_kwargs.pop("mcusher"), usher_kwargs.pop("step_lengths"),
                           usher_kwargs.pop("step_probabilities", None))
            self._multistep = ms.device_tables()
            usher_kwargs.setdefault("sublattice_probabilities", ms.sublattice_probabilities
                                    if len(ms.sublattice_probabilities) == len(ensemble.active_sublattices) else None)
            self.mcusher = ms
        self.nwalkers = int(nwalkers)
        if seeds is None:
            ss = np.random.SeedSequence()
            seeds = [int(x) for x in ss.generate_state(self.nwalkers, dtype=np.uint64)]
        if len(seeds) != self.nwalkers:
            raise ValueError("Number of seeds does not match number of kernels!")
        self.seeds = np.array([int(s) & 0xFFFFFFFFFFFFFFFF for s in seeds], dtype=np.uint64)
        self.kB = kB_
        self._temperature = np.zeros(self.nwalkers)
        if self._kernel == capi.LMC_KERNEL_METROPOLIS:
            if self._uniform:
                self._temperature[:] = np.inf
            else:
                if temperature is None:
                    raise TypeError("Metropolis kernel needs a temperature")
                self._temperature[:] = float(temperature)
        self._wl = dict(wl_params or {})
        table_flip = None
        if self._usher == capi.LMC_USHER_TABLEFLIP:
            if usher_kwargs.get("flip_table") is None:
                raise NotImplementedError(
                    "TableFlip needs an explicit flip_table (CompositionSpace generation of the "
                    "table, smol/moca/composition/space.py, is outside the hot path)")
            table_flip = table_flip_tables(ensemble.sublattices, usher_kwargs["flip_table"],
                                           usher_kwargs.get("flip_weights"),
                                           usher_kwargs.get("swap_weight", 0.1))
        self._packed = ensemble.packed_model(
            table_flip=table_flip,
            sublattice_probabilities=usher_kwargs.get("sublattice_probabilities"))
        self.engine = LmcEngine(self._packed, device=device)
        self.walker_id_base = int(walker_id_base)
        self.group_size, self.block_threads = int(group_size), int(block_threads)
        # Metropolis flip/swap kernel: 0 auto (by measured acceptance), 1 classic, 2 speculative batch
        self.spec_mode = int(spec_mode)
        # speculative kernel over per-site environment words (one packed word per flip and lane instead of three
        # occupancy gathers per record) where the model has the tables.  Opt-in: the words live in L2 (16-32 bytes
        # per site and walker do not fit shared memory at 28 walkers per SM) and the exposed L2 latency costs more than
        # the gathers save on the measured configurations
        self.spec_env = bool(spec_env) or os.environ.get("LMC_SPEC_ENV_DEFAULT") == "1"
        # Ewald term through the per-walker potential cache (O(1) per flip, one row of the site kernel per
        # accepted flip) instead of gathering matrix rows at every flip: "auto" = while fewer than a quarter
        # of the steps are accepted (needs an Ewald matrix of the form q_i q_j K[site_i, site_j])
        if ewald_field == "auto" and os.environ.get("LMC_EWALD_FIELD") in ("0", "1"):   # A/B switch
            ewald_field = os.environ["LMC_EWALD_FIELD"] == "1"
        if ewald_field not in ("auto", True, False):
            raise ValueError("ewald_field must be 'auto', True or False")
        self.ewald_field = ewald_field
        self._ew_field = None
        self._spec_env_ws = None
        self._acc_est = None
        self.record_occupancy = record_occupancy
        # bias term of the Metropolis exponent (kernel/base.py:229-235, metropolis.py:43-44)
        self.bias = None
        if bias_type is not None:
            if self._kernel != capi.LMC_KERNEL_METROPOLIS:
                raise ValueError(f"{bias_type} is not a valid MCBias for this kernel.")   # wanglandau.py:24
            from .bias import MCBias, mcbias_factory
            self.bias = bias_type if isinstance(bias_type, MCBias) else \
                mcbias_factory(bias_type, ensemble.sublattices, **(bias_kwargs or {}))
            if self.bias.table.shape[0] != ensemble.num_sites:
                raise ValueError("the bias table does not cover all sites of the ensemble")
        self._bias_dev = None
        self.mckernels = [_KernelView(self, i) for i in range(self.nwalkers)]
        self._step_counter = 0
        self._occ_dev = None
        self._wl_state = None
        N, F = ensemble.num_sites, len(ensemble.natural_parameters)
        shapes = {"occupancy": ((N,), np.int32), "features": ((F,), np.float64),
                  "enthalpy": ((1,), np.float64), "accepted": ((1,), bool),
                  "n_accepted": ((), np.int32)}
        if self._kernel == capi.LMC_KERNEL_METROPOLIS:
            shapes["temperature"] = ((1,), np.float64)
        if self.bias is not None:
            shapes["bias"] = ((1,), np.float64)
        # Wang-Landau: the kernel's arrays are part of every sample (wanglandau.py:247-251).  "full" as the
        # reference, "no_means" without cumulative_mean_features ([bins, F] per walker and sample), "none"
        if wl_trace not in ("full", "no_means", "none"):
            raise ValueError("wl_trace must be 'full', 'no_means' or 'none'")
        self._wl_trace = []
        if self._kernel == capi.LMC_KERNEL_WANGLANDAU and wl_trace != "none":
            p = self._wl
            nb = len(np.arange(p["min_enthalpy"], p["max_enthalpy"], p["bin_size"]))
            self._wl_trace = [("entropy", (nb,), np.float64), ("histogram", (nb,), np.int64),
                              ("occurrences", (nb,), np.int64), ("mod_factor", (1,), np.float64)]
            if wl_trace == "full":
                self._wl_trace.append(("cumulative_mean_features", (nb, F), np.float64))
            for name, shape, dtype in self._wl_trace:
                shapes[name] = (shape, dtype)
        self._container = SampleContainer(ensemble, self.nwalkers, shapes,
                                          dict(ensemble.thermo_boundaries))

    # ---- construction (sampler.py:52-139) ------------------------------------------------------
    @classmethod
    def from_ensemble(cls, ensemble, *args, step_type=None, kernel_type=None, seeds=None, nwalkers=1,
                      **kwargs):
        if step_type is None:
            step_type = "flip" if getattr(ensemble, "chemical_potentials", None) is not None else "swap"
        if kernel_type is None:
            kernel_type = "Metropolis"
        engine_kw = {k: kwargs.pop(k) for k in ("walker_id_base", "group_size", "block_threads", "spec_mode",
                                                "record_occupancy", "device", "ewald_field", "bias_type", "bias_kwargs", "wl_trace", "spec_env") if k in kwargs}
        key = kernel_type.lower().replace("_", "").replace("-", "")
        temperature, wl = None, None
        if key == "wanglandau":
            names = ("min_enthalpy", "max_enthalpy", "bin_size")
            vals = list(args)
            wl = {}
            for n in names:
                wl[n] = kwargs.pop(n) if n in kwargs else vals.pop(0)
            for n, dflt in (("flatness", 0.8), ("mod_factor", 1.0), ("check_period", 1000),
                            ("update_period", 1), ("mod_update", None)):
                wl[n] = kwargs.pop(n, dflt)
        elif key == "metropolis":
            temperature = kwargs.pop("temperature") if "temperature" in kwargs else args[0]
        usher_kwargs = kwargs
        return cls(ensemble, kernel_type=kernel_type, step_type=step_type, nwalkers=nwalkers,
                   seeds=seeds, temperature=temperature, wl_params=wl, usher_kwargs=usher_kwargs,
                   **engine_kw)

    @property
    def samples(self):
        return self._container

    def _slot_key(self, index, nmax, W, N, F):
        return (index, nmax, W, N, F, self.record_occupancy, self.bias is not None, len(self._wl_trace))

    def _trace_slot(self, index, nmax, W, N, F, dev):
        """Cached trace slot ``index``: device buffers + page-locked host staging for ``nmax`` samples."""
        import torch
        cache = self.__dict__.setdefault("_slots", {})
        key = self._slot_key(index, nmax, W, N, F)
        if cache.get(index, {}).get("key") != key:
            shapes = {"features": ((nmax, W, F), torch.float64), "enthalpy": ((nmax, W), torch.float64),
                      "accepted": ((nmax, W), torch.uint8), "n_accepted": ((nmax, W), torch.int32),
                      "occupancy": ((nmax, W, N if self.record_occupancy else 0), torch.int8)}
            if self.bias is not None:
                shapes["bias"] = ((nmax, W), torch.float64)
            for name, shape, dtype in self._wl_trace:
                tail = () if name == "mod_factor" else shape
                shapes[name] = ((nmax, W, *tail), torch.float64 if dtype == np.float64 else torch.int64)
            cache[index] = {"key": key, "shapes": shapes,
                            "dev": {k: torch.empty(sh, dtype=dt, device=dev) for k, (sh, dt) in shapes.items()},
                            "host": None}
        return cache[index]

    @staticmethod
    def _slot_staging(slot):
        """fixed page-locked staging of a slot (fallback when the pinned pool is exhausted)"""
        import torch
        if slot["host"] is None:
            slot["host"] = {k: torch.empty(sh, dtype=dt, pin_memory=True) for k, (sh, dt) in slot["shapes"].items()}
        return slot["host"]

    def detach_samples(self):
        """Hand over the samples collected so far and start an empty container: with ``run(block=False)`` the
        returned container resolves its last chunk when it is first read, while the sampler already runs on."""
        old = self._container
        self._container = SampleContainer(self.ensemble, self.nwalkers, dict(old._shapes), dict(old.metadata))
        return old

    def efficiency(self, discard=0, flat=True):
        return self.samples.sampling_efficiency(discard=discard, flat=flat)

    def clear_samples(self):
        self.samples.clear()

    # ---- Wang-Landau state ------------------------------------------------------------------------
    def _init_wl(self):
        import torch
        p = self._wl
        if p["min_enthalpy"] > p["max_enthalpy"]:
            raise ValueError("min_enthalpy can not be larger than max_enthalpy.")
        if (p["max_enthalpy"] - p["min_enthalpy"]) / p["bin_size"] <= 1:
            raise ValueError("The values provided for min and max enthalpy and bin sizer result in "
                             "a single bin!")
        if p["mod_factor"] <= 0:
            raise ValueError("mod_factor must be greater than 0.")
        # a callable mod_update (wanglandau.py:100-105) cannot run on the device: its orbit from the initial factor is
        # tabulated here (every walker's factor is always a member of it) and a flatness event steps along the table
        self._wl_mod_table = None
        if callable(p.get("mod_update")):
            tab = [float(p["mod_factor"])]
            for _ in range(255):
                nxt = float(p["mod_update"](tab[-1]))
                if not np.isfinite(nxt) or nxt == tab[-1]:
                    break
                if nxt in tab:
                    raise ValueError("mod_update revisits a modification factor: the sequence must not cycle")
                tab.append(nxt)
            self._wl_mod_table = np.array(tab, dtype=np.float64)
        levels = np.arange(p["min_enthalpy"], p["max_enthalpy"], p["bin_size"])  # wanglandau.py:107
        nb, W, F = len(levels), self.nwalkers, self.engine.F
        dev = self.engine.device
        self._wl_state = dict(
            levels=levels,
            entropy=torch.zeros((W, nb), dtype=torch.float64, device=dev),
            histogram=torch.zeros((W, nb), dtype=torch.int64, device=dev),
            occurrences=torch.zeros((W, nb), dtype=torch.int64, device=dev),
            mean_features=torch.zeros((W, nb, F), dtype=torch.float64, device=dev),
            mod_factor=torch.full((W,), float(p["mod_factor"]), dtype=torch.float64, device=dev),
            mod_table=None if self._wl_mod_table is None else torch.from_numpy(self._wl_mod_table).to(dev),
            steps_counter=torch.zeros((W,), dtype=torch.int64, device=dev))

    @property
    def wang_landau_state(self):
        """Per-walker arrays (entropy, histogram, occurrences, mean_features, mod_factor) on host."""
        if self._wl_state is None:
            return None
        out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in self._wl_state.items()}
        if int(self._wl["update_period"]) == 1:
            # the device accumulates per-bin feature SUMS; with update_period == 1 the reference's running
            # mean (wanglandau.py:234-238) equals sum / occurrences
            occ = out["occurrences"].astype(np.float64)[:, :, None]
            out["mean_features"] = np.divide(out["mean_features"], occ, out=np.zeros_like(out["mean_features"]),
                                             where=occ > 0)
        return out

    # ---- run (sampler.py:164-297, 386-434) ------------------------------------------------------------
    def _begin_run(self, initial_occupancies, thin_by):
        """Everything ``Sampler.run`` does before the step loop (sampler.py:386-434): copy of the initial
        occupancies onto the device, initial trace (full evaluation), auxiliary states of the kernels."""
        import torch
        from types import SimpleNamespace
        eng = self.engine
        if initial_occupancies is None:
            if self._occ_dev is None:
                raise RuntimeError("There are no saved samples to obtain the initial occupancies."
                                   "These must be provided.")
        else:
            if self.samples.num_samples > 0:
                warnings.warn("Initial occupancies where provided with a pre-existing set of samples."
                              "\n Make real sure that is what you want. If not, reset the samples in "
                              "the sampler.", RuntimeWarning)
            occ = initial_occupancies if isinstance(initial_occupancies, torch.Tensor) else np.asarray(initial_occupancies)
            if occ.ndim == 1 and self.nwalkers == 1:
                occ = occ[None, :]
            if tuple(occ.shape) != (self.nwalkers, eng.N):
                raise AttributeError("The given initial occcupancies have incompompatible dimensions. "
                                     f"Shape should be {(self.nwalkers, eng.N)}.")
            # copied (sampler.py:401) and converted to int32 (sampler.py:406) on the way into the
            # page-locked staging buffer; the caller's array is never modified.  A CUDA tensor never
            # touches the host (device-resident entry).
            self._occ_dev = eng.upload_occupancy(occ)
        W, N, F = self.nwalkers, eng.N, eng.F
        dev = eng.device
        ctx = SimpleNamespace(thin_by=thin_by)
        # initial trace: full evaluation of the starting occupancies (base.py:345-365,
        # wanglandau.py:290-300)
        # Ewald potential cache: rebuilt from the occupancies at every run (bounds its rounding drift to one
        # run), kept current by the kernels in between.  With a factorising Ewald matrix the full evaluation
        # computes it anyway (Ewald energy = sum_k q_k field[k] + diagonal terms).
        ctx.use_field = False
        if self.ewald_field is not False and eng.model_info()[0]:
            ctx.use_field = self.ewald_field is True or self._acc_est is None or self._acc_est < 0.25
        elif self.ewald_field is True:
            raise RuntimeError("ewald_field=True needs an Ewald term whose matrix factorises as q_i q_j K[site_i, site_j]")
        self.ewald_cache_in_use = ctx.use_field     # (the field buffer doubles as scratch of the full evaluation)
        if eng.model_info()[0] and (self._ew_field is None or self._ew_field.shape[0] != self.nwalkers):
            self._ew_field = torch.empty((self.nwalkers, eng.N), dtype=torch.float64, device=dev)
        ctx.feat, ctx.enth = eng.full_features(self._occ_dev, field=self._ew_field if eng.model_info()[0] else None)
        from .processor import DistanceProcessor
        ctx.dist_proc = self.ensemble.processor if isinstance(self.ensemble.processor, DistanceProcessor) else None
        ctx.dist_vec = None
        if ctx.dist_proc is not None:
            # distance processors: features = distance vector, running correlation vector kept beside it
            ctx.dist_vec = eng.distance_init(ctx.dist_proc, ctx.feat, ctx.enth)
        if self.bias is not None:
            # initial bias of the starting occupancies (base.py:362-363), then kept current by the kernel
            if self._bias_dev is None:
                self._bias_dev = dict(
                    table=torch.from_numpy(np.ascontiguousarray(self.bias.table, dtype=np.float64)).to(dev),
                    value=torch.empty((W,), dtype=torch.float64, device=dev),
                    tsum=torch.empty((W, self.bias.rows), dtype=torch.float64, device=dev))
            b = self._bias_dev
            icpt = self.bias._intercept_array()
            capi.check(eng.lib.lmc_bias_init(self._occ_dev.data_ptr(), W, N, self.bias.mode, self.bias.table.shape[1],
                                             self.bias.rows, float(self.bias.penalty), icpt.ctypes.data,
                                             b["table"].data_ptr(), b["value"].data_ptr(), b["tsum"].data_ptr(),
                                             eng._stream()))
        if self._kernel == capi.LMC_KERNEL_WANGLANDAU and self._wl_state is None:
            self._init_wl()
        if getattr(self, "_seeds_dev", None) is None:
            self._seeds_dev = torch.from_numpy(self.seeds.view(np.int64)).to(dev)
        ctx.seeds = self._seeds_dev
        with np.errstate(divide="ignore"):
            beta_h = np.where(np.isinf(self._temperature), 0.0, 1.0 / (self.kB * self._temperature))
        bkey = beta_h.tobytes()
        if getattr(self, "_beta_key", None) != bkey:     # re-uploaded only when a temperature changed
            self._beta_dev = torch.from_numpy(np.ascontiguousarray(beta_h)).to(dev)
            self._beta_key = bkey
        ctx.beta = self._beta_dev
        # Metropolis flip / swap kernel (classic or speculative batches): decided HERE from the acceptance of the
        # runs this sampler has completed, so identical call sequences take identical kernels
        ctx.spec_mode = self.spec_mode
        if ctx.spec_mode == 0:
            ctx.spec_mode = 3 if (self._acc_est is None or self._acc_est < 0.35) else 1
        # workspace of the speculative kernel's environment words (lmc.h: LmcRunConfig.spec_env_dev), rebuilt by every
        # launch from the occupancies: allocated once per sampler, only where a speculative launch can happen
        env_bytes = eng.model_info()[4]
        if (self.spec_env and env_bytes > 0 and ctx.spec_mode != 1 and self._kernel == capi.LMC_KERNEL_METROPOLIS
                and self._usher in (capi.LMC_USHER_FLIP, capi.LMC_USHER_SWAP)):
            if self._spec_env_ws is None or self._spec_env_ws.numel() != self.nwalkers * env_bytes:
                self._spec_env_ws = torch.empty((self.nwalkers * env_bytes,), dtype=torch.uint8, device=dev)
            ctx.spec_env = self._spec_env_ws
        else:
            ctx.spec_env = None
        return ctx

    def _run_config(self, ctx, d, n):
        """``LmcRunConfig`` of one launch: ``n`` sampling intervals into the device trace buffers ``d``."""
        eng = self.engine
        thin_by = ctx.thin_by
        cfg = capi.LmcRunConfig()
        cfg.num_walkers, cfg.walker_id_base = self.nwalkers, self.walker_id_base
        cfg.usher, cfg.kernel = self._usher, self._kernel
        cfg.num_samples, cfg.thin_by = n, thin_by
        cfg.group_size, cfg.block_threads = self.group_size, self.block_threads
        cfg.spec_mode = ctx.spec_mode
        cfg.step_begin = self._step_counter
        cfg.seeds_dev, cfg.beta_dev = ctx.seeds.data_ptr(), ctx.beta.data_ptr()
        cfg.occ_dev, cfg.features_dev, cfg.enthalpy_dev = \
            self._occ_dev.data_ptr(), ctx.feat.data_ptr(), ctx.enth.data_ptr()
        cfg.trace_occ_dev = d["occupancy"].data_ptr() if self.record_occupancy else None
        cfg.trace_features_dev, cfg.trace_enthalpy_dev = d["features"].data_ptr(), d["enthalpy"].data_ptr()
        cfg.trace_accepted_dev, cfg.trace_naccepted_dev = d["accepted"].data_ptr(), d["n_accepted"].data_ptr()
        cfg.ewald_field_dev = self._ew_field.data_ptr() if ctx.use_field else None
        cfg.spec_env_dev = ctx.spec_env.data_ptr() if ctx.spec_env is not None else None
        if ctx.dist_proc is not None:
            dt = eng.distance_tables(ctx.dist_proc)
            cfg.dist_mode, cfg.dist_num_groups, cfg.dist_tol = 1, dt["ngrp"], dt["tol"]
            cfg.dist_target_dev, cfg.dist_group_off_dev = dt["target"].data_ptr(), dt["goff"].data_ptr()
            cfg.dist_group_idx_dev, cfg.dist_group_diam_dev = dt["gidx"].data_ptr(), dt["gdiam"].data_ptr()
            cfg.dist_vector_dev = ctx.dist_vec.data_ptr()
        if self._multistep is not None:
            code, lens, cum = self._multistep
            cfg.ms_usher, cfg.ms_num = code, len(lens)
            for i, ln in enumerate(lens):
                cfg.ms_len[i] = ln
                cfg.ms_cum[i] = float(cum[i])
        if self._composite is not None:
            codes, cum, sl_cum = self._composite
            cfg.comp_num = len(codes)
            for i, code in enumerate(codes):
                cfg.comp_usher[i] = code
                cfg.comp_cum[i] = float(cum[i])
                for k in range(capi.LMC_MAX_SUBLATTICES):
                    cfg.comp_sl_cum[i][k] = float(sl_cum[i, k])
        if self.bias is not None:
            b = self._bias_dev
            cfg.bias_mode, cfg.bias_width, cfg.bias_rows = self.bias.mode, self.bias.table.shape[1], self.bias.rows
            cfg.bias_penalty = float(self.bias.penalty)
            cfg.bias_table_dev, cfg.bias_dev, cfg.bias_sum_dev = b["table"].data_ptr(), b["value"].data_ptr(), b["tsum"].data_ptr()
            cfg.trace_bias_dev = d["bias"].data_ptr()
        if self._kernel == capi.LMC_KERNEL_WANGLANDAU:
            p, st = self._wl, self._wl_state
            wl = cfg.wl
            wl.min_enthalpy, wl.max_enthalpy, wl.bin_size = p["min_enthalpy"], p["max_enthalpy"], p["bin_size"]
            wl.flatness = p["flatness"]
            wl.mod_update = float(p["mod_update"]) if (p.get("mod_update") is not None and not callable(p["mod_update"])) else 2.0
            if st.get("mod_table") is not None:
                wl.mod_table_dev, wl.mod_table_len = st["mod_table"].data_ptr(), int(st["mod_table"].numel())
            wl.num_bins, wl.check_period, wl.update_period = len(st["levels"]), p["check_period"], p["update_period"]
            wl.reserved = 1 if int(p["update_period"]) == 1 else 0   # mean_features buffer holds sums
            wl.entropy_dev, wl.histogram_dev = st["entropy"].data_ptr(), st["histogram"].data_ptr()
            wl.occurrences_dev, wl.mean_features_dev = st["occurrences"].data_ptr(), st["mean_features"].data_ptr()
            wl.mod_factor_dev, wl.steps_counter_dev = st["mod_factor"].data_ptr(), st["steps_counter"].data_ptr()
            for name, _, _ in self._wl_trace:
                field = {"entropy": "trace_entropy_dev", "histogram": "trace_histogram_dev",
                         "occurrences": "trace_occurrences_dev", "mod_factor": "trace_mod_factor_dev",
                         "cumulative_mean_features": "trace_mean_features_dev"}[name]
                setattr(wl, field, d[name].data_ptr())
        return cfg

    def _bytes_per_sample(self):
        """device / page-locked bytes of one sampling interval over all walkers (every trace array)"""
        W, N, F = self.nwalkers, self.engine.N, self.engine.F
        per_walker = (N if self.record_occupancy else 0) + 8 * F + 8 + 1 + 4
        if self.bias is not None:
            per_walker += 8
        for _, shape, dtype in self._wl_trace:
            per_walker += int(np.prod(shape)) * np.dtype(dtype).itemsize
        return W * per_walker

    def run(self, nsteps, initial_occupancies=None, thin_by=1, progress=False, stream_chunk=0,
            stream_file=None, keep_last_chunk=False, swmr_mode=False, max_chunk_bytes=2 << 30,
            pipeline_chunks=4, block=True):
        """``Sampler.run`` (sampler.py:212-301).

        ``initial_occupancies`` may be a host array, a page-locked ``torch.int32`` tensor (copied straight from it)
        or a CUDA tensor (never touches the host).  ``block=False`` returns as soon as every launch and copy of
        the run is enqueued: the last chunk of traces is handed to the sample container when the samples are
        first looked at (or by the next ``run``, after ITS first launch -- back-to-back runs then overlap the
        upload and initial evaluation of run k+1 with the tail of run k).  ``stream_chunk`` / ``stream_file`` /
        ``keep_last_chunk`` stream the samples to a file backend every ``stream_chunk`` samples
        (sampler.py:271-297, container.py:420-512)."""
        import torch
        eng = self.engine
        with torch.cuda.device(eng.device):
            return self._run(nsteps, initial_occupancies, thin_by, stream_chunk, stream_file, keep_last_chunk,
                             swmr_mode, max_chunk_bytes, pipeline_chunks, block, progress)

    def _run(self, nsteps, initial_occupancies, thin_by, stream_chunk, stream_file, keep_last_chunk, swmr_mode,
             max_chunk_bytes, pipeline_chunks, block, progress=False):
        import torch
        eng = self.engine
        if stream_chunk > 0:
            if nsteps % stream_chunk != 0:
                raise ValueError("streaming chunk must be a divisor of nsteps.")     # sampler.py:281-282
            if stream_chunk % thin_by != 0:
                raise ValueError("streaming chunk must be a multiple of thin_by.")
            backend = self.samples.get_backend(stream_file, nsteps // thin_by, swmr_mode=swmr_mode)
        # (a deferred tail of the previous run is resolved after this run's first launch, see below)
        ctx = self._begin_run(initial_occupancies, thin_by)
        if nsteps % thin_by != 0:
            warnings.warn(f"The number of steps {nsteps} is not a multiple of thin_by  {thin_by}. "
                          f"The last {nsteps % thin_by} will be ignored.", category=RuntimeWarning)
        S = nsteps // thin_by
        W, N, F = self.nwalkers, eng.N, eng.F
        dev = eng.device
        per_sample = self._bytes_per_sample()
        # The run is cut into a few launches so that the device->host copy and the host-side
        # bookkeeping of chunk i overlap the kernel of chunk i+1 (double-buffered trace slots, copies
        # on a side stream).  The chains are unaffected: the RNG is counter based.
        nmax = max(1, min(S, int(max_chunk_bytes // max(per_sample, 1)))) if S else 0
        if not block and stream_chunk == 0:
            # non-blocking runs overlap their copies with the NEXT run's kernel: one launch per run (every launch pays
            # the start-up of its blocks -- staging, bit-planes, environment words -- once more)
            pass
        elif S >= 2 * pipeline_chunks:
            nmax = min(nmax, -(-S // pipeline_chunks))
        elif S >= 2:
            nmax = min(nmax, -(-S // 2))
        if stream_chunk > 0:
            per_flush = stream_chunk // thin_by       # launches end on flush boundaries
            nmax = max(k for k in range(1, max(1, min(nmax, per_flush)) + 1) if per_flush % k == 0)
        self._kernel_events = []
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        if self.samples.has_deferred and self._slot_key(0, nmax, W, N, F) != self.__dict__.get("_slots", {}).get(0, {}).get("key"):
            self.samples.resolve_deferred()      # the trace slots are about to be re-allocated under a pending copy
        # non-blocking runs give every launch its own trace slot and defer EVERY chunk: the call returns once all
        # launches and copies are enqueued, so the next run's upload / initial evaluation / launches are queued long
        # before the device needs them (no bubble between back-to-back runs)
        async_all = (not block) and stream_chunk == 0 and S > 0
        nslots = max(2, min(8, -(-S // nmax))) if async_all else 2
        slots = [self._trace_slot(i, nmax, W, N, F, dev) for i in range(nslots)]
        # the previous run's deferred chunks may still be copied out of the slots they used: keep rotating
        phase = self.__dict__.get("_slot_phase", 0) % nslots

        def finalize(host, pooled, n, ev_copy):
            ev_copy.synchronize()
            owned = []
            if pooled:
                # the trace arrays ARE the page-locked blocks the copy engine wrote: no host copy
                arrs = {k: v.numpy() for k, v in host.items()}
                for k, v in host.items():
                    owned.append(((lambda recycle=True, t=v: _PINNED.release(t, recycle)), arrs[k]))
                traces = {"features": arrs["features"][:n], "enthalpy": arrs["enthalpy"][:n][:, :, None],
                          "accepted": arrs["accepted"][:n].view(np.bool_)[:, :, None],
                          "n_accepted": arrs["n_accepted"][:n]}
                if self.bias is not None:
                    traces["bias"] = arrs["bias"][:n][:, :, None]
                for name, _, _ in self._wl_trace:
                    traces[name] = arrs[name][:n][:, :, None] if name == "mod_factor" else arrs[name][:n]
                if self.record_occupancy:
                    traces["occupancy"] = arrs["occupancy"][:n]          # int8; int32 on access
                del arrs
            else:
                traces = {
                    "features": _fast_copy(host["features"][:n].numpy()),
                    "enthalpy": host["enthalpy"][:n].numpy().copy()[:, :, None],
                    "accepted": host["accepted"][:n].numpy().astype(bool)[:, :, None],
                    "n_accepted": host["n_accepted"][:n].numpy().copy(),
                }
                if self.bias is not None:
                    traces["bias"] = host["bias"][:n].numpy().copy()[:, :, None]
                for name, _, _ in self._wl_trace:
                    v = host[name][:n].numpy().copy()
                    traces[name] = v[:, :, None] if name == "mod_factor" else v
                if self.record_occupancy:
                    o = host["occupancy"][:n].numpy()
                    traces["occupancy"] = _fast_copy(o.reshape(n * W, N)).reshape(n, W, N)   # int8; int32 on access
            if not self.record_occupancy:
                traces["occupancy"] = np.zeros((n, W, 0), dtype=np.int8)
            if "temperature" in self.samples._shapes:
                traces["temperature"] = np.broadcast_to(temperature[None, :, None], (n, W, 1)).copy()
            if n and W:   # acceptance of the chunk: steers the kernel / Ewald path of the next run
                self._acc_est = float(traces["n_accepted"][-1].mean()) / thin_by
            return traces, thin_by, owned

        temperature = self._temperature.copy()
        done, ci, pending = 0, 0, None
        bar = None
        if progress:      # sampler.py:190-194 shows steps per walker; here the bar advances launch by launch
            try:
                from tqdm import tqdm
                bar = tqdm(total=S * thin_by, desc=f"Sampling {self.nwalkers} walker(s) on {eng.device}", unit="step")
            except ImportError:
                bar = None
        while done < S:
            n = min(nmax, S - done)
            slot = slots[(ci + phase) % nslots]
            d = slot["dev"]
            cfg = self._run_config(ctx, d, n)
            if slot.get("copy_event") is not None:
                main.wait_event(slot["copy_event"])     # (a deferred tail may still be reading this slot)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(main)
            eng.run(cfg)
            ev1.record(main)
            self._kernel_events.append((ev0, ev1))
            self._step_counter += n * thin_by
            # device -> pinned host staging on the copy stream, behind this chunk's kernel
            names = [k for k in d if not (k == "occupancy" and not self.record_occupancy)]
            host = {k: _PINNED.acquire((n, *d[k].shape[1:]), d[k].dtype) for k in names}
            pooled = all(v is not None for v in host.values())
            if not pooled:
                for v in host.values():
                    if v is not None:
                        _PINNED.release(v)
                self.samples.resolve_deferred()
                host = self._slot_staging(slot)
            ev_copy = torch.cuda.Event()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ev1)
                for k in names:
                    host[k][:n].copy_(d[k][:n], non_blocking=True)
                ev_copy.record(self._copy_stream)
            slot["copy_event"] = ev_copy
            if async_all and pooled:
                self.samples.defer(n, thin_by, lambda p=(host, pooled, n, ev_copy): finalize(*p))
                done += n
                ci += 1
                continue
            # tail of the PREVIOUS run: resolved only now, behind this run's upload, initial evaluation and
            # first launch, which therefore overlapped it
            self.samples.resolve_deferred()
            if pending is not None:
                self.samples.append(*finalize(*pending))        # overlaps the kernel just launched
                if stream_chunk > 0 and self.samples.num_samples >= per_flush:
                    self.samples.flush_to_backend(backend)      # sampler.py:286-291
            pending = (host, pooled, n, ev_copy)
            done += n
            ci += 1
            if bar is not None:
                bar.update(n * thin_by)
        if bar is not None:
            bar.update(S * thin_by - bar.n)
            bar.close()
        if not async_all:
            self.samples.resolve_deferred()
        self._slot_phase = (ci + phase) % nslots     # slot the next launch would take
        events, self._kernel_events = self._kernel_events, []
        self._last_events = events
        if pending is not None:
            if block or stream_chunk > 0 or not pending[1]:
                self.samples.append(*finalize(*pending))
            else:
                n_tail = pending[2]
                self.samples.defer(n_tail, thin_by, lambda p=pending: finalize(*p))
        if stream_chunk > 0:
            if self.samples.num_samples:
                self.samples.flush_to_backend(backend)
            backend.close()
            if keep_last_chunk is False:
                self.clear_samples()                             # sampler.py:293-297
        if block:
            torch.cuda.synchronize(dev)
        elif getattr(eng, "_pin_evt", None) is not None:
            eng._pin_evt.synchronize()       # the caller's (page-locked) input buffer has been consumed

    @property
    def last_kernel_ms(self):
        """device time of the lmc_run launches of the last ``run`` (CUDA events on the launching stream)"""
        ev = getattr(self, "_last_events", None)
        if not ev:
            return 0.0
        ev[-1][1].synchronize()
        return float(sum(a.elapsed_time(b) for a, b in ev))

    def run_device(self, nsteps, initial_occupancies=None, thin_by=1, out=None, reuse_state=False):
        """Device-resident form of ``run``: the same chains, but every trace stays in HBM.

        Returns a dict of CUDA tensors ``[S, W, ...]`` (``occupancy`` int8 codes, ``features`` / ``enthalpy``
        float64, ``accepted`` uint8, ``n_accepted`` int32, bias / Wang-Landau arrays when traced) for callers that
        post-process on the GPU; nothing is copied to the host and nothing is appended to ``samples``.
        ``initial_occupancies``: host array or CUDA tensor (int32 ``[W, N]``), ``None`` continues the chains.
        ``out``: the dict of a previous call with the same shape, to be overwritten (no allocation).
        ``reuse_state=True`` (with ``initial_occupancies=None``) continues from the running features / enthalpies
        the previous ``run_device`` left on the device instead of re-evaluating the occupancies in full (the
        reference re-evaluates at every ``run``, sampler.py:423-429; the kernels keep the running values
        current, so the chains are the same up to the rounding of the running sums)."""
        import torch
        eng = self.engine
        with torch.cuda.device(eng.device):
            self.samples.resolve_deferred()
            prev = getattr(self, "_resident_ctx", None)
            if reuse_state and initial_occupancies is None and prev is not None and prev.thin_by == thin_by:
                ctx = prev
            else:
                ctx = self._begin_run(initial_occupancies, thin_by)
                self._resident_ctx = ctx
            S = nsteps // thin_by
            W, N, F = self.nwalkers, eng.N, eng.F
            if out is None:
                slot = self._trace_slot("resident", S, W, N, F, eng.device)
                out = slot["dev"]
            if S:
                main = torch.cuda.current_stream(eng.device)
                cfg = self._run_config(ctx, out, S)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(main)
                eng.run(cfg)
                ev1.record(main)
                self._last_events = [(ev0, ev1)]
                self._step_counter += S * thin_by
            return out

    def anneal(self, temperatures, mcmc_steps, initial_occupancies=None, thin_by=1, progress=False,
               stream_chunk=0, stream_file=None, swmr_mode=True):
        """sampler.py:303-384: one ``run`` per temperature, each starting from the last configuration of the
        previous one; everything goes to the same container (or, streaming, to the same file)."""
        if temperatures[0] < temperatures[-1]:
            raise ValueError("End temperature is greater than start "
                             f"temperature {temperatures[-1]:.2f} > {temperatures[0]:.2f}.")
        for i, t in enumerate(temperatures):
            for k in self.mckernels:
                k.temperature = t
            self.run(mcmc_steps, initial_occupancies=initial_occupancies if i == 0 else None, thin_by=thin_by,
                     progress=progress, stream_chunk=stream_chunk, stream_file=stream_file, swmr_mode=swmr_mode,
                     keep_last_chunk=True)
        if stream_chunk > 0:       # if streaming to file was done then clear the samples now (sampler.py:382-384)
            self.clear_samples()

    def current_occupancies(self):
        """int32 ``[W, N]`` host copy of the walkers' current occupancies."""
        eng = self.engine
        return eng.occupancy_to_int32(self._occ_dev, self.nwalkers, eng.row_stride).cpu().numpy()
