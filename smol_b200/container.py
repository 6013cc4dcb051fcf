"""SampleContainer: in-memory mirror of ``smol/moca/sampler/container.py`` (array part).

Trace names, shapes ``[nsamples, nwalkers, ...]`` and dtypes follow
``smol/moca/trace.py`` / ``sampler.py:123-128``: ``occupancy`` int32, ``features`` /
``enthalpy`` / ``temperature`` float64, ``accepted`` bool (flag of the last step of each
thinning interval, ``sampler.py:199-201``).  ``n_accepted`` (accepted steps per interval) is an
engine extension.  Persistence (``container.py:420-692``): ``flush_to_backend`` / ``get_backend`` /
``to_hdf5`` / ``from_hdf5`` write the reference's layout (groups ``metadata`` and ``trace``, one
resizable dataset per trace, attributes ``nsamples`` / ``total_mc_steps``) through h5py when it is
installed and otherwise through :class:`DirectoryBackend`, a directory of ``.npy`` memory maps that
offers the same subset of the h5py interface; ``as_dict`` / ``from_dict`` use the reference's keys.
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np


class _Attrs(dict):
    """attribute dict of a group, written through to ``attrs.json`` of its directory"""

    def __init__(self, path):
        super().__init__()
        self._path = path
        if os.path.exists(path):
            with open(path) as f:
                super().update(json.load(f))

    def __setitem__(self, k, v):
        super().__setitem__(k, int(v) if isinstance(v, (int, np.integer)) else v)
        self.flush()

    def flush(self):
        tmp = self._path + ".tmp"
        with open(tmp, "w") as f:
            json.dump(dict(self), f)
        os.replace(tmp, self._path)        # readers (SWMR) always see a complete file


class _Dataset:
    """one ``.npy`` memory map; ``resize`` along axis 0 rewrites the file (rare: the sampler allocates up front)"""

    def __init__(self, path, shape=None, dtype=None, mode="r+"):
        self.path = path
        if shape is not None:
            self._mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.dtype(dtype), shape=tuple(shape))
        else:
            self._mm = np.load(path, mmap_mode=mode)

    shape = property(lambda self: self._mm.shape)
    dtype = property(lambda self: self._mm.dtype)

    def __len__(self):
        return self._mm.shape[0]

    def __getitem__(self, key):
        return np.asarray(self._mm[key])

    def __setitem__(self, key, value):
        self._mm[key] = value

    def resize(self, size, axis=0):
        assert axis == 0
        old = self._mm
        tmp = self.path + ".grow.npy"
        new = np.lib.format.open_memmap(tmp, mode="w+", dtype=old.dtype, shape=(int(size), *old.shape[1:]))
        n = min(len(old), int(size))
        new[:n] = old[:n]
        new.flush()
        del new, old
        self._mm = None
        os.replace(tmp, self.path)
        self._mm = np.load(self.path, mmap_mode="r+")

    def flush(self):
        if hasattr(self._mm, "flush"):
            self._mm.flush()


class _Group:
    def __init__(self, path, mode):
        self.path, self._mode = path, mode
        os.makedirs(path, exist_ok=True)
        self.attrs = _Attrs(os.path.join(path, "attrs.json"))
        self._items = {}

    def create_group(self, name):
        self._items[name] = _Group(os.path.join(self.path, name), self._mode)
        return self._items[name]

    def create_dataset(self, name, shape=None, dtype=None, maxshape=None, data=None):
        if data is not None and shape is None:          # scalar / string payload (metadata)
            with open(os.path.join(self.path, name + ".json"), "w") as f:
                json.dump(data, f)
            self._items[name] = data
        else:
            self._items[name] = _Dataset(os.path.join(self.path, name + ".npy"), shape, dtype)
        return self._items[name]

    def __contains__(self, name):
        return name in self.keys()

    def keys(self):
        names = set(self._items)
        for f in os.listdir(self.path):
            full = os.path.join(self.path, f)
            if os.path.isdir(full):
                names.add(f)
            elif f.endswith(".npy") and not f.endswith(".grow.npy"):
                names.add(f[:-4])
            elif f.endswith(".json") and f != "attrs.json":
                names.add(f[:-5])
        return sorted(names)

    def __iter__(self):
        return iter(self.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, name):
        if name not in self._items:
            base = os.path.join(self.path, name)
            if os.path.isdir(base):
                self._items[name] = _Group(base, self._mode)
            elif os.path.exists(base + ".npy"):
                self._items[name] = _Dataset(base + ".npy", mode="r" if self._mode == "r" else "r+")
            elif os.path.exists(base + ".json"):
                with open(base + ".json") as f:
                    self._items[name] = json.load(f)
            else:
                raise KeyError(name)
        return self._items[name]

    def flush(self):
        for v in self._items.values():
            if hasattr(v, "flush"):
                v.flush()


class DirectoryBackend(_Group):
    """File backend with the part of the ``h5py.File`` interface ``SampleContainer`` uses, stored as a directory:
    ``<path>/trace/<name>.npy`` (memory maps), ``<path>/trace/attrs.json``, ``<path>/metadata/*.json``.  Attributes
    are replaced atomically after the data they describe is flushed, so another process may read the directory while
    it is being written (the reference's single-writer / multiple-reader mode)."""

    def __init__(self, path, mode="r+"):
        if mode == "w-" and os.path.exists(path):
            raise FileExistsError(path)
        super().__init__(path, mode)
        self.swmr_mode = False

    def close(self):
        self.flush()
        self._items = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:
        return None


class SampleContainer:
    def __init__(self, ensemble, nwalkers, trace_shapes, sampling_metadata=None):
        self._ensemble = ensemble
        self.metadata = dict(sampling_metadata or {})
        self._nwalkers = nwalkers
        self._shapes = dict(trace_shapes)      # name -> (shape tuple, dtype)
        self._chunks = {name: [] for name in self._shapes}
        self._cache = {}
        self._thin = []
        self._owned = []          # (release callback, base ndarray) of page-locked blocks the chunks view
        self._deferred = []       # (n, thinned_by, resolve): chunks whose device->host copy is still in flight
        self._flushed = False     # samples were written to a backend: kept readable, replaced by the next append
        self._aux_checkpoint = None
        self.total_mc_steps = 0

    def __del__(self):
        # page-locked blocks of a dropped container stop counting against the pool's budget
        try:
            for entry in self._owned:
                if entry is not None:
                    try:
                        entry[0](False)
                    except TypeError:
                        pass
        except Exception:
            pass

    # ---- bookkeeping ----------------------------------------------------------------------
    @property
    def ensemble(self):
        return self._ensemble

    @property
    def sublattices(self):
        return self._ensemble.sublattices

    @property
    def natural_parameters(self):
        return self._ensemble.natural_parameters

    @property
    def num_samples(self):
        if self._flushed:            # container.py:435-437: a flush rewinds the write position
            return int(sum(d[0] for d in self._deferred))
        ch = self._chunks["enthalpy"]
        return int(sum(len(c) for c in ch)) + int(sum(d[0] for d in self._deferred))

    # ---- chunks still on their way from the device (Sampler.run(block=False)) --------------------
    @property
    def has_deferred(self):
        return bool(self._deferred)

    def defer(self, n, thinned_by, resolve):
        """``resolve()`` -> ``(traces, thinned_by, owned)`` once the chunk's copy has landed"""
        self._deferred.append((int(n), int(thinned_by), resolve))

    def resolve_deferred(self):
        while self._deferred:
            _, _, resolve = self._deferred.pop(0)
            self.append(*resolve())

    def __len__(self):
        return self.num_samples

    @property
    def shape(self):
        return (self._nwalkers, self._shapes["occupancy"][0][0])

    @property
    def traced_values(self):
        return list(self._shapes)

    def append(self, traces: dict, thinned_by: int, owned=None):
        """container.py:384-397 for a whole block of samples at once.

        ``owned``: ``(release, base)`` pairs for trace arrays that are views of page-locked blocks
        written directly by the device->host copy (no staging copy).  ``clear`` hands a block back
        through ``release`` only when nothing outside this container still refers to it."""
        n = len(traces["enthalpy"])
        if self._flushed:
            self._drop_chunks()
        for name in self._shapes:
            self._chunks[name].append(traces[name])
        if owned:
            self._owned.extend(owned)
        self._cache.clear()
        self.total_mc_steps += n * thinned_by
        self._thin.append((n, thinned_by))

    def clear(self):
        self.resolve_deferred()
        self._drop_chunks()

    def _drop_chunks(self):
        import sys
        for name in self._chunks:
            self._chunks[name] = []
        self._cache.clear()
        self.total_mc_steps = 0
        self._thin = []
        self._flushed = False
        owned, self._owned = self._owned, []
        for i in range(len(owned)):
            release, base = owned[i]
            owned[i] = None
            # references left: the local name and getrefcount's argument; anything more is a view
            # the caller still holds, and its memory must not be recycled under it
            if sys.getrefcount(base) <= 2:
                release()
            else:
                try:
                    release(False)     # not recycled, but no longer counted against the page-locked budget
                except TypeError:
                    pass

    def _full(self, name):
        self.resolve_deferred()
        if name not in self._cache:
            shape, dtype = self._shapes[name]
            ch = self._chunks[name]
            if not ch:
                arr = np.empty((0, self._nwalkers, *shape), dtype=dtype)
            else:
                arr = np.concatenate(ch, axis=0) if len(ch) > 1 else ch[0]
                if arr.dtype != dtype:
                    arr = arr.astype(dtype)
                self._chunks[name] = [arr]
            self._cache[name] = arr
        return self._cache[name]

    @staticmethod
    def _flatten(values):
        """container.py:514-519: samples x walkers merged AND size-one axes squeezed (flat enthalpies are 1-D)."""
        s = values.shape
        return np.squeeze(values.reshape((s[0] * s[1], *s[2:])))

    # ---- accessors (container.py:131-381) ---------------------------------------------------
    def get_trace_value(self, name, discard=0, thin_by=1, flat=True):
        value = self._full(name)[discard + thin_by - 1:: thin_by]
        return self._flatten(value) if flat else value

    def mean_trace_value(self, name, discard=0, thin_by=1, flat=True):
        return self.get_trace_value(name, discard, thin_by, flat).mean(axis=0)

    def trace_value_variance(self, name, discard=0, thin_by=1, flat=True):
        return self.get_trace_value(name, discard, thin_by, flat).var(axis=0)

    def sampling_efficiency(self, discard=0, flat=True):
        total_accepted = self._full("accepted")[discard:].sum(axis=0)
        eff = total_accepted / (self.num_samples - discard)
        return eff.mean() if flat else eff

    def step_efficiency(self, discard=0, flat=True):
        """Exact accepted fraction over ALL attempted steps (engine extension)."""
        nacc = self._full("n_accepted")[discard:].sum(axis=0).astype(np.float64)
        steps = 0
        seen = 0
        for n, thin in self._thin:
            lo = max(discard - seen, 0)
            steps += max(n - lo, 0) * thin
            seen += n
        eff = nacc / max(steps, 1)
        return eff.mean() if flat else eff

    def get_occupancies(self, discard=0, thin_by=1, flat=True):
        return self.get_trace_value("occupancy", discard, thin_by, flat)

    def get_enthalpies(self, discard=0, thin_by=1, flat=True):
        return self.get_trace_value("enthalpy", discard, thin_by, flat)

    def get_feature_vectors(self, discard=0, thin_by=1, flat=True):
        return self.get_trace_value("features", discard, thin_by, flat)

    def get_energies(self, discard=0, thin_by=1, flat=True):
        """container.py:208-229: the enthalpies when there are no extra terms, else features[:n_energy] . coefs;
        shape ``[S, W, 1]`` like the enthalpy trace (flattened and squeezed when ``flat``)."""
        n = self._ensemble.num_energy_coefs
        if len(self.natural_parameters) == n:
            return self.get_enthalpies(discard, thin_by, flat)
        feats = self.get_feature_vectors(discard, thin_by, flat=False)
        energies = np.tensordot(feats[..., :n], self.natural_parameters[:n], axes=([-1], [0]))[..., None]
        return self._flatten(energies) if flat else energies

    def get_temperatures(self, discard=0, thin_by=1):
        """container.py:231-233 (always flat)."""
        return self.get_trace_value("temperature", discard, thin_by)

    def mean_enthalpy(self, discard=0, thin_by=1, flat=True):
        return self.get_enthalpies(discard, thin_by, flat).mean(axis=0)

    def enthalpy_variance(self, discard=0, thin_by=1, flat=True):
        return self.get_enthalpies(discard, thin_by, flat).var(axis=0)

    def mean_energy(self, discard=0, thin_by=1, flat=True):
        return self.get_energies(discard, thin_by, flat).mean(axis=0)

    def energy_variance(self, discard=0, thin_by=1, flat=True):
        return self.get_energies(discard, thin_by, flat).var(axis=0)

    def mean_feature_vector(self, discard=0, thin_by=1, flat=True):
        return self.get_feature_vectors(discard, thin_by, flat).mean(axis=0)

    def feature_vector_variance(self, discard=0, thin_by=1, flat=True):
        return self.get_feature_vectors(discard, thin_by, flat).var(axis=0)

    def get_minimum_enthalpy(self, discard=0, thin_by=1, flat=True):
        return self.get_enthalpies(discard, thin_by, flat).min(axis=0)

    def get_minimum_enthalpy_occupancy(self, discard=0, thin_by=1, flat=True):
        inds = self.get_enthalpies(discard, thin_by, flat).argmin(axis=0)
        occus = self.get_occupancies(discard, thin_by, flat)
        return occus[inds] if flat else occus[inds, np.arange(self._nwalkers)][0]

    def get_minimum_energy(self, discard=0, thin_by=1, flat=True):
        """container.py:321-323."""
        return self.get_energies(discard, thin_by, flat).min(axis=0)

    def get_minimum_energy_occupancy(self, discard=0, thin_by=1, flat=True):
        """container.py:325-334."""
        inds = self.get_energies(discard, thin_by, flat).argmin(axis=0)
        occus = self.get_occupancies(discard, thin_by, flat)
        return occus[inds] if flat else occus[inds, np.arange(self._nwalkers)][0]

    def get_sublattice_species_counts(self, sublattice, discard=0, thin_by=1, flat=True):
        """container.py:349-382: counts of each species of a sublattice, last axis in the order of its site space
        (``sublattice.encoding``); one vectorised comparison per code instead of np.unique per sample."""
        if not any(sublattice is s for s in self.sublattices) and sublattice not in list(self.sublattices):
            raise ValueError("Sublattice provided is not recognized.\n Provide one included"
                             " in the sublattices attribute of this SampleContainer.")
        occus = self.get_occupancies(discard, thin_by, flat=False)[..., np.asarray(sublattice.sites, dtype=int)]
        counts = np.stack([np.count_nonzero(occus == code, axis=-1) for code in sublattice.encoding],
                          axis=-1).astype(np.float64)
        return self._flatten(counts) if flat else counts

    def get_species_counts(self, discard=0, thin_by=1, flat=True):
        """container.py:336-347: counts per species summed over the sublattices, keyed by species.  Like the
        reference (``subcounts.T``) the chain form is ``[walkers, samples]``, transposed w.r.t. the traces."""
        counts = {}
        for s in self.sublattices:
            sub = self.get_sublattice_species_counts(s, discard, thin_by, flat)
            for sp, c in zip(s.species, sub.T):
                counts[sp] = counts[sp] + c if sp in counts else c.copy()
        return counts

    def get_sublattice_compositions(self, sublattice, discard=0, thin_by=1, flat=True):
        """container.py:235-238."""
        return self.get_sublattice_species_counts(sublattice, discard, thin_by, flat) / len(sublattice.sites)

    def get_compositions(self, discard=0, thin_by=1, flat=True):
        """container.py:240-243: species counts over ALL sites of the supercell."""
        counts = self.get_species_counts(discard, thin_by, flat)
        return {sp: c / self.shape[1] for sp, c in counts.items()}

    def mean_composition(self, discard=0, thin_by=1, flat=True):
        """container.py:283-286."""
        return {sp: c.mean(axis=0) for sp, c in self.get_compositions(discard, thin_by, flat).items()}

    def composition_variance(self, discard=0, thin_by=1, flat=True):
        """container.py:288-291."""
        return {sp: c.var(axis=0) for sp, c in self.get_compositions(discard, thin_by, flat).items()}

    def mean_sublattice_composition(self, sublattice, discard=0, thin_by=1, flat=True):
        """container.py:293-297."""
        return self.get_sublattice_compositions(sublattice, discard, thin_by, flat).mean(axis=0)

    def sublattice_composition_variance(self, sublattice, discard=0, thin_by=1, flat=True):
        """container.py:299-305."""
        return self.get_sublattice_compositions(sublattice, discard, thin_by, flat).var(axis=0)

    def get_orbit_factors(self, function_orbit_ids, discard=0, thin_by=1, flat=True):
        """container.py:269-281 (sum of natural parameter x feature over the functions of each orbit id)."""
        vals = self.natural_parameters * self.get_feature_vectors(discard=discard, thin_by=thin_by, flat=flat)
        ids = np.asarray(function_orbit_ids)
        return np.array([np.sum(vals[..., ids == i]) for i in range(len(self.natural_parameters))])

    def vacuum(self):
        """container.py:399-411 trims unused pre-allocated rows; the chunks here hold sampled rows only."""
        return None

    # ---- persistence (container.py:420-692) ---------------------------------------------------
    def _trace_arrays(self):
        """name -> full array in the reference's dtype, reference names only"""
        return {name: self._full(name) for name in self._shapes if name != "n_accepted"}

    def flush_to_backend(self, backend):
        """container.py:420-437: append the samples held in memory to the backend and rewind; the flushed
        samples stay readable until the next one arrives (the reference overwrites its arrays in place)."""
        self.resolve_deferred()
        trace = backend["trace"]
        start = int(trace.attrs["nsamples"])
        nsamples = 0 if self._flushed else self.num_samples
        end = start + nsamples
        if nsamples:
            for name, value in self._trace_arrays().items():
                if len(trace[name]) < end:
                    trace[name].resize(end, axis=0)
                trace[name][start:end] = value
        backend.flush()
        trace.attrs["total_mc_steps"] = int(trace.attrs["total_mc_steps"]) + (0 if self._flushed else self.total_mc_steps)
        trace.attrs["nsamples"] = end
        backend.flush()
        self.total_mc_steps = 0
        self._thin = []
        self._flushed = True

    def get_backend(self, file_path, alloc_nsamples=0, swmr_mode=False):
        """container.py:439-475: open (or create) the backend file and make room for ``alloc_nsamples`` more.

        HDF5 through h5py when it is installed; without h5py -- or for a path ending in ``.lmc`` -- a
        :class:`DirectoryBackend` with the same layout."""
        if file_path is None:
            raise ValueError("a file path is needed to stream samples")
        h5 = None if str(file_path).endswith(".lmc") else _h5py()
        exists = os.path.isfile(file_path) if h5 is not None else os.path.isdir(file_path)
        if exists:
            backend = self._check_backend(file_path, h5)
            trace_grp = backend["trace"]
            available = len(trace_grp["occupancy"]) - int(trace_grp.attrs["nsamples"])
            if available < alloc_nsamples:
                self._grow_backend(backend, alloc_nsamples - available)
        else:
            backend = h5.File(file_path, "w-", libver="latest") if h5 is not None else DirectoryBackend(file_path, "w-")
            self._init_backend(backend, alloc_nsamples)
        if swmr_mode:
            backend.swmr_mode = swmr_mode
        return backend

    def _check_backend(self, file_path, h5):
        """container.py:477-489."""
        backend = h5.File(file_path, mode="r+", libver="latest") if h5 is not None else DirectoryBackend(file_path, "r+")
        shape = tuple(backend["trace"]["occupancy"].shape[1:])
        if tuple(self.shape) != shape:
            backend.close()
            raise RuntimeError(f"Backend file {file_path} has incompatible dimensions {self.shape}, {shape}.")
        return backend

    def _init_backend(self, backend, nsamples):
        """container.py:491-512."""
        metadata = backend.create_group("metadata")
        metadata.create_dataset("ensemble", data=json.dumps(_ensemble_dict(self._ensemble)))
        metadata.create_dataset("sampling_metadata", data=json.dumps(_jsonable(self.metadata)))
        trace_grp = backend.create_group("trace")
        for name, (shape, dtype) in self._shapes.items():
            if name == "n_accepted":
                continue
            trace_grp.create_dataset(name, shape=(nsamples, self._nwalkers, *shape), dtype=np.dtype(dtype),
                                     maxshape=(None, self._nwalkers, *shape))
        trace_grp.attrs["nsamples"] = 0
        trace_grp.attrs["total_mc_steps"] = 0
        backend.flush()

    @staticmethod
    def _grow_backend(backend, nsamples):
        """container.py:514-520."""
        for name in backend["trace"]:
            backend["trace"][name].resize(len(backend["trace"][name]) + nsamples, axis=0)

    def to_hdf5(self, file_path):
        """container.py:615-629: save (or append to) a backend file; the container keeps its samples."""
        self.resolve_deferred()
        keep = (self.total_mc_steps, list(self._thin), self._flushed)
        backend = self.get_backend(file_path, self.num_samples)
        self.flush_to_backend(backend)
        self.total_mc_steps, self._thin, self._flushed = keep
        backend.close()

    @classmethod
    def from_hdf5(cls, file_path, swmr_mode=True, ensemble=None):
        """container.py:631-692.  ``ensemble``: the Ensemble the samples came from (this build does not rebuild
        processors from their serialised form; without it the container carries sublattices and natural
        parameters only, like the reference's legacy files)."""
        h5 = None if (str(file_path).endswith(".lmc") or os.path.isdir(file_path)) else _h5py()
        f = h5.File(file_path, "r", swmr=swmr_mode) if h5 is not None else DirectoryBackend(file_path, "r")
        try:
            nsamples = int(f["trace"].attrs["nsamples"])
            if len(f["trace"]["occupancy"]) > nsamples:
                warnings.warn(f"The hdf5 file provided appears to be from an unifinished MC run.\n Only {nsamples} of "
                              f" {len(f['trace']['occupancy'])} samples have been written and will be loaded.",
                              UserWarning)
            trace = {name: np.array(value[:nsamples]) for name, value in f["trace"].items()}
            meta = json.loads(f["metadata"]["sampling_metadata"][()] if h5 is not None
                              else f["metadata"]["sampling_metadata"])
            ens_d = json.loads(f["metadata"]["ensemble"][()] if h5 is not None else f["metadata"]["ensemble"])
            total = int(f["trace"].attrs["total_mc_steps"])
        finally:
            f.close()
        return cls._from_arrays(trace, meta, ens_d, total, ensemble)

    @classmethod
    def _from_arrays(cls, trace, metadata, ensemble_d, total_mc_steps, ensemble=None):
        if ensemble is None:
            ensemble = _FrozenEnsemble(ensemble_d)
        elif ensemble_d is not None and len(ensemble_d.get("sublattices", [])) != len(ensemble.sublattices):
            raise ValueError("Sublattices in Ensemble object passed do not match, the once saved. \n "
                             "Make sure you are passing the correct Ensemble object.")     # container.py:567-574
        occ = trace["occupancy"]
        nwalkers = occ.shape[1]
        shapes = {name: (tuple(v.shape[2:]), v.dtype.type if v.dtype != np.bool_ else bool) for name, v in trace.items()}
        if "n_accepted" not in trace:
            trace = dict(trace)
            trace["n_accepted"] = trace["accepted"].reshape(occ.shape[0], nwalkers).astype(np.int32)
            shapes["n_accepted"] = ((), np.int32)
        container = cls(ensemble, nwalkers, shapes, metadata)
        if occ.shape[0]:
            nsteps = int(total_mc_steps)
            container.append(trace, nsteps // max(occ.shape[0], 1))
        container.total_mc_steps = int(total_mc_steps)
        return container

    def as_dict(self):
        """container.py:525-543 (same keys)."""
        trace = {name: value.tolist() for name, value in self._trace_arrays().items()}
        return {"@module": self.__class__.__module__, "@class": self.__class__.__name__,
                "ensemble": _ensemble_dict(self._ensemble), "metadata": _jsonable(self.metadata),
                "total_mc_steps": int(self.total_mc_steps), "nsamples": int(self.num_samples), "trace": trace,
                "aux_checkpoint": self._aux_checkpoint}

    @classmethod
    def from_dict(cls, d, ensemble=None):
        """container.py:577-613."""
        trace = {key: np.array(val) for key, val in d["trace"].items()}
        if "occupancy" in trace:
            trace["occupancy"] = trace["occupancy"].astype(np.int32)
        if "accepted" in trace:
            trace["accepted"] = trace["accepted"].astype(bool)
        container = cls._from_arrays(trace, d.get("metadata") or {}, d.get("ensemble"), d["total_mc_steps"], ensemble)
        container._aux_checkpoint = d.get("aux_checkpoint")
        return container


def _jsonable(obj):
    if isinstance(obj, dict):
        return {str(k): _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_jsonable(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.generic):
        return obj.item()
    return obj


def _ensemble_dict(ensemble):
    """what post-processing needs of the ensemble (sublattices, natural parameters, chemical potentials); the
    processor tables are model construction and are not serialised by this build"""
    if ensemble is None:
        return None
    if isinstance(ensemble, _FrozenEnsemble):
        return ensemble._d
    return {"@class": "Ensemble",
            "sublattices": [{"species": [str(sp) for sp in s.species], "sites": np.asarray(s.sites).tolist(),
                             "active_sites": np.asarray(s.active_sites).tolist(),
                             "encoding": np.asarray(s.encoding).tolist()} for s in ensemble.sublattices],
            "natural_parameters": np.asarray(ensemble.natural_parameters).tolist(),
            "num_energy_coefs": int(ensemble.num_energy_coefs), "num_sites": int(ensemble.num_sites),
            "chemical_potentials": _jsonable(getattr(ensemble, "chemical_potentials", None)),
            "thermo_boundaries": _jsonable(getattr(ensemble, "thermo_boundaries", {}))}


class _FrozenEnsemble:
    """stand-in for the Ensemble of a loaded container (the reference's legacy form: sublattices, natural parameters
    and the number of energy coefficients, container.py:545-575)"""

    def __init__(self, d):
        from .sublattice import Sublattice
        self._d = d or {}
        self.sublattices = []
        for s in self._d.get("sublattices", []):
            sub = Sublattice(tuple(s["species"]), np.array(s["sites"], dtype=np.int64))
            sub.active_sites = np.array(s["active_sites"], dtype=np.int64)
            sub.encoding = np.array(s["encoding"], dtype=np.int32)
            self.sublattices.append(sub)
        self.natural_parameters = np.array(self._d.get("natural_parameters", []), dtype=np.float64)
        self.num_energy_coefs = int(self._d.get("num_energy_coefs", len(self.natural_parameters)))
        self.num_sites = int(self._d.get("num_sites", 0))
        self.chemical_potentials = self._d.get("chemical_potentials")
        self.thermo_boundaries = self._d.get("thermo_boundaries", {})
