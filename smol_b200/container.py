"""SampleContainer: in-memory mirror of ``smol/moca/sampler/container.py`` (array part).

Trace names, shapes ``[nsamples, nwalkers, ...]`` and dtypes follow
``smol/moca/trace.py`` / ``sampler.py:123-128``: ``occupancy`` int32, ``features`` /
``enthalpy`` / ``temperature`` float64, ``accepted`` bool (flag of the last step of each
thinning interval, ``sampler.py:199-201``).  ``n_accepted`` (accepted steps per interval) is an
engine extension.  HDF5 streaming (``container.py:420-512``) is not part of this build.
"""
from __future__ import annotations

import numpy as np


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
        self.total_mc_steps = 0

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
        ch = self._chunks["enthalpy"]
        return int(sum(len(c) for c in ch))

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
        for name in self._shapes:
            self._chunks[name].append(traces[name])
        if owned:
            self._owned.extend(owned)
        self._cache.clear()
        self.total_mc_steps += n * thinned_by
        self._thin.append((n, thinned_by))

    def clear(self):
        import sys
        for name in self._chunks:
            self._chunks[name] = []
        self._cache.clear()
        self.total_mc_steps = 0
        self._thin = []
        owned, self._owned = self._owned, []
        for i in range(len(owned)):
            release, base = owned[i]
            owned[i] = None
            # references left: the local name and getrefcount's argument; anything more is a view
            # the caller still holds, and its memory must not be recycled under it
            if sys.getrefcount(base) <= 2:
                release()

    def _full(self, name):
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
