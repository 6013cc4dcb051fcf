"""Sublattice data model (mirror of ``smol/moca/sublattice.py:23-226``).

``site_space`` is the ordered tuple of species labels allowed on the sites (position ==
occupancy code); with smol installed a ``SiteSpace`` works as well (only ``len`` and
iteration are used).
"""
from __future__ import annotations

import warnings

import numpy as np


class Sublattice:
    """Sites sharing one site space; only ``active_sites`` may change during MC."""

    def __init__(self, site_space, sites):
        self.site_space = site_space
        self.sites = np.unique(np.asarray(sites, dtype=np.int64))          # sublattice.py:55
        self.active_sites = self.sites.copy()
        if len(self.site_space) <= 1:                                      # sublattice.py:57-59
            self.restrict_sites(self.sites)
        self.encoding = np.arange(len(self.site_space), dtype=np.int32)    # sublattice.py:61

    @property
    def is_active(self):
        if len(self.active_sites) == 0 and len(self.species) > 1:
            warnings.warn("Sub-lattice is inactive, but have multiple allowed species. "
                          "You'd better split it.")
        return len(self.active_sites) > 0

    @property
    def species(self):
        keys = getattr(self.site_space, "keys", None)
        return tuple(keys()) if keys is not None else tuple(self.site_space)

    @property
    def restricted_sites(self):
        return np.setdiff1d(self.sites, self.active_sites)

    def restrict_sites(self, sites):
        """sublattice.py:91-101."""
        sites = set(int(s) for s in np.atleast_1d(sites))
        self.active_sites = np.array([i for i in self.active_sites if int(i) not in sites],
                                     dtype=np.int64)

    def reset_restricted_sites(self):
        if len(self.site_space) > 1:
            self.active_sites = self.sites.copy()

    def split_by_species(self, occu, species_in_partitions):
        """sublattice.py:109-184 for integer codes: one new sublattice per code partition."""
        occu = np.asarray(occu)
        part_codes = [sorted(int(c) for c in part) for part in species_in_partitions]
        flat = [c for p in part_codes for c in p]
        if sorted(flat) != sorted(int(c) for c in self.encoding):
            raise ValueError("partitions must cover the sublattice encoding exactly once")
        out = []
        species = self.species
        for codes in part_codes:
            sel = self.sites[np.isin(occu[self.sites], codes)]
            sub = Sublattice.__new__(Sublattice)
            sub.site_space = tuple(species[list(self.encoding).index(c)] for c in codes)
            sub.sites = sel
            sub.active_sites = np.array([s for s in sel if s in set(self.active_sites.tolist())],
                                        dtype=np.int64)
            if len(codes) <= 1:
                sub.active_sites = np.array([], dtype=np.int64)
            sub.encoding = np.array(codes, dtype=np.int32)
            out.append(sub)
        return out

    def __repr__(self):
        return (f"Sublattice(species={self.species}, n_sites={len(self.sites)}, "
                f"n_active={len(self.active_sites)}, encoding={self.encoding.tolist()})")
