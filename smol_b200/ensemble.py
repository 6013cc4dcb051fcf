"""Ensemble: mirror of ``smol/moca/ensemble.py`` in front of the GPU processors."""
from __future__ import annotations

import numpy as np

from .processor import (ClusterDecompositionProcessor, ClusterExpansionProcessor,
                        CompositeProcessor, EwaldProcessor, _as_occu)


class Ensemble:
    """``ensemble.py:102-430``: natural parameters, sublattices, chemical potentials."""

    def __init__(self, processor, sublattices=None, chemical_potentials=None):
        if sublattices is None:
            sublattices = processor.get_sublattices()
        self.thermo_boundaries = {}
        self._params = np.array(processor.coefs, dtype=np.float64)
        self._processor = processor
        self._sublattices = sublattices
        self._chemical_potentials = None
        self._engine = None
        self.chemical_potentials = chemical_potentials

    @classmethod
    def from_cluster_expansion(cls, cluster_expansion, supercell_matrix,
                               processor_type="decomposition", use_concentration=False, **kwargs):
        """ensemble.py:132-217."""
        from .processor import _reject_use_concentration
        _reject_use_concentration(use_concentration)
        subspace = cluster_expansion.cluster_subspace
        has_ext = len(getattr(subspace, "external_terms", [])) > 0
        coefs = cluster_expansion.coefs[:-1] if has_ext else cluster_expansion.coefs
        if processor_type == "decomposition":
            ce = ClusterDecompositionProcessor(subspace, supercell_matrix,
                                               cluster_expansion.cluster_interaction_tensors)
        elif processor_type == "expansion":
            ce = ClusterExpansionProcessor(subspace, supercell_matrix, coefs)
        else:
            raise ValueError(f"Processor type {processor_type} not supported!")
        if has_ext:
            processor = CompositeProcessor(subspace, supercell_matrix)
            processor.add_processor(ce)
            processor.add_processor(EwaldProcessor(subspace, supercell_matrix,
                                                   ewald_term=subspace.external_terms[0],
                                                   coefficient=cluster_expansion.coefs[-1]))
        else:
            processor = ce
        return cls(processor, **kwargs)

    # ---- simple properties ----------------------------------------------------------------
    @property
    def num_sites(self):
        return self.processor.num_sites

    @property
    def num_energy_coefs(self):
        return len(self._processor.coefs)

    @property
    def system_size(self):
        return self.processor.size

    @property
    def processor(self):
        return self._processor

    @property
    def sublattices(self):
        return self._sublattices

    @property
    def active_sublattices(self):
        return [s for s in self.sublattices if s.is_active]

    @property
    def restricted_sites(self):
        return np.concatenate([s.restricted_sites for s in self.sublattices])

    @property
    def species(self):
        """ensemble.py:256-265."""
        return list({sp for s in self.active_sublattices for sp in s.species})

    @property
    def natural_parameters(self):
        return self._params

    @natural_parameters.setter
    def natural_parameters(self, value):
        """ensemble.py:278-286."""
        if not np.array_equal(self.processor.coefs, value[: self.num_energy_coefs]):
            raise ValueError("The original expansion coefficients can not be changed!")
        self._params = np.array(value, dtype=np.float64)

    # ---- chemical potentials (ChemicalPotentialManager, ensemble.py:20-99) ------------------
    @property
    def chemical_potentials(self):
        return None if self._chemical_potentials is None else self._chemical_potentials["value"]

    @chemical_potentials.setter
    def chemical_potentials(self, value):
        self._engine = None
        if value is None:
            if self._chemical_potentials is not None:
                self._chemical_potentials = None
                self.thermo_boundaries.pop("chemical_potentials", None)
                if self.num_energy_coefs < len(self._params):
                    self._params = self._params[:-1]
            return
        value = {k: v for k, v in value.items() if k in self.species}
        if set(value.keys()) != set(self.species):
            raise ValueError("Chemical potentials given are missing species. Values must be given "
                             f"for each of the following: {self.species}")
        if self._chemical_potentials is None:
            self._params = np.append(self._params, -1.0)       # natural parameter, ensemble.py:25
        self._chemical_potentials = {"value": dict(value), "table": self._build_table(value)}
        self.thermo_boundaries["chemical_potentials"] = dict(value)

    def _build_table(self, value):
        """ensemble.py:89-99."""
        num_cols = max(max(s.encoding) for s in self.sublattices) + 1
        table = np.zeros((self.num_sites, num_cols))
        for s in self.active_sublattices:
            table[s.sites[:, None], s.encoding] = [value[sp] for sp in s.species]
        return table

    @property
    def mu_table(self):
        return None if self._chemical_potentials is None else self._chemical_potentials["table"]

    # ---- site restrictions (ensemble.py:378-398) ---------------------------------------------
    def restrict_sites(self, sites):
        for s in self.sublattices:
            s.restrict_sites(sites)
        self._engine = None

    def reset_restricted_sites(self):
        for s in self.sublattices:
            s.reset_restricted_sites()
        self._engine = None

    def split_sublattice_by_species(self, sublattice_id, occu, species_in_partitions):
        """ensemble.py:288-321."""
        splits = self.sublattices[sublattice_id].split_by_species(occu, species_in_partitions)
        self._sublattices = (self._sublattices[:sublattice_id] + splits
                             + self._sublattices[sublattice_id + 1:])
        if self.chemical_potentials is not None:
            self.chemical_potentials = {sp: self.chemical_potentials[sp] for sp in self.species}
        self._engine = None

    # ---- evaluation ----------------------------------------------------------------------------
    def packed_model(self, table_flip=None, sublattice_probabilities=None):
        from .model import PackedModel
        return PackedModel(self.num_sites, self.natural_parameters, self.sublattices,
                           mu_table=self.mu_table, table_flip=table_flip,
                           sublattice_probabilities=sublattice_probabilities,
                           **self.processor._tables())

    def _get_engine(self):
        if self._engine is None:
            from .engine import LmcEngine
            self._engine = LmcEngine(self.packed_model())
        return self._engine

    def compute_feature_vector_batch(self, occupancies):
        eng = self._get_engine()
        occ = eng.upload_occupancy(_as_occu(occupancies).reshape(-1, self.num_sites))
        return eng.full_features(occ)[0].cpu().numpy()

    def compute_feature_vector_change_batch(self, occupancies, sites, codes):
        import torch
        eng = self._get_engine()
        occ = eng.upload_occupancy(_as_occu(occupancies).reshape(-1, self.num_sites))
        s = torch.as_tensor(np.ascontiguousarray(sites, dtype=np.int32)).to(eng.device)
        c = torch.as_tensor(np.ascontiguousarray(codes, dtype=np.int32)).to(eng.device)
        return eng.delta_features(occ, s.reshape(occ.shape[0], -1).contiguous(),
                                  c.reshape(occ.shape[0], -1).contiguous()).cpu().numpy()

    def compute_feature_vector(self, occupancy):
        """ensemble.py:323-351."""
        return self.compute_feature_vector_batch(_as_occu(occupancy)[None, :])[0]

    def compute_feature_vector_change(self, occupancy, step):
        """ensemble.py:353-376."""
        if len(step) == 0:
            return np.zeros(len(self.natural_parameters))
        sites = np.array([[f[0] for f in step]], dtype=np.int32)
        codes = np.array([[f[1] for f in step]], dtype=np.int32)
        return self.compute_feature_vector_change_batch(_as_occu(occupancy)[None, :], sites, codes)[0]
