"""Extractor for live ``smol`` objects: build the GPU ensemble from a ``smol.moca.Ensemble``.

Only attributes of the reference's public / serialised interface are read:

* processors: ``cluster_subspace``, ``supercell_matrix``, ``coefs``, ``allowed_species`` (``processor/base.py:59-107``),
  ``ClusterDecompositionProcessor._interaction_tensors`` (``expansion.py:324``, kept "for serialization"),
  ``EwaldProcessor.ewald_matrix`` / ``_ewald_inds`` (``ewald.py:78, 92-101``), ``CompositeProcessor.processors``
  (``composite.py:56-59``);
* the subspace: ``orbits[i].{id, bit_id, flat_tensor_indices, flat_correlation_tensors}``, ``num_orbits``,
  ``num_corr_functions``, ``orbit_multiplicities``, ``get_orbit_indices(scm).arrays`` (``clusterspace.py``);
* the ensemble: ``processor``, ``sublattices`` (``site_space``, ``sites``, ``active_sites``, ``encoding``),
  ``chemical_potentials`` (``ensemble.py:219-321``).

smol and pymatgen are not importable in the build container; ``tests/test_host_logic.py`` drives this module with
objects that expose exactly these names.
"""
from __future__ import annotations

import numpy as np

from .ensemble import Ensemble
from .processor import (ClusterDecompositionProcessor, ClusterExpansionProcessor, CompositeProcessor,
                        EwaldProcessor)
from .sublattice import Sublattice


class _SubspaceView:
    """The reference subspace plus ``allowed_species(scm)``, which the reference keeps on the processor."""

    def __init__(self, subspace, allowed_species):
        self._subspace = subspace
        self._allowed = [tuple(sp) for sp in allowed_species]

    def allowed_species(self, scmatrix):
        return self._allowed

    def __getattr__(self, name):
        return getattr(self._subspace, name)

    def __eq__(self, other):
        return self is other or getattr(other, "_subspace", other) is self._subspace


def _convert_processor(p, view):
    kind = type(p).__name__
    scm = np.array(p.supercell_matrix)
    if kind == "ClusterDecompositionProcessor":
        return ClusterDecompositionProcessor(view, scm, p._interaction_tensors, coefficients=np.array(p.coefs))
    if kind == "ClusterExpansionProcessor":
        return ClusterExpansionProcessor(view, scm, np.array(p.coefs))
    if kind == "EwaldProcessor":
        return EwaldProcessor(view, scm, ewald_term=getattr(p, "_ewald_term", None),
                              coefficient=float(np.ravel(p.coefs)[0]), ewald_matrix=np.array(p.ewald_matrix),
                              ewald_inds=np.array(p._ewald_inds))
    raise NotImplementedError(f"{kind} has no GPU counterpart")


def from_smol_processor(processor):
    """GPU processor with the tables of a live ``smol.moca`` processor."""
    view = _SubspaceView(processor.cluster_subspace, processor.allowed_species)
    if type(processor).__name__ == "CompositeProcessor":
        out = CompositeProcessor(view, np.array(processor.supercell_matrix))
        for p in processor.processors:
            out.add_processor(_convert_processor(p, view))
        return out
    return _convert_processor(processor, view)


def from_smol_sublattice(sublattice):
    """``smol.moca.Sublattice`` -> ``smol_b200.Sublattice`` (same sites, active sites and encoding)."""
    out = Sublattice(sublattice.site_space, np.array(sublattice.sites))
    out.active_sites = np.array(sublattice.active_sites, dtype=np.int64)
    out.encoding = np.array(sublattice.encoding, dtype=np.int32)
    return out


def from_smol_ensemble(ensemble):
    """``smol.moca.Ensemble`` -> ``smol_b200.Ensemble``: then ``smol_b200.Sampler.from_ensemble(...)`` as usual."""
    gpu = from_smol_processor(ensemble.processor)
    subl = [from_smol_sublattice(s) for s in ensemble.sublattices]
    mus = ensemble.chemical_potentials
    if mus is not None:
        mus = {str(k): float(v) for k, v in dict(mus).items()}
        for s in subl:   # species keys as strings on both sides
            s.site_space = tuple(str(sp) for sp in s.species)
    return Ensemble(gpu, sublattices=subl, chemical_potentials=mus)
