"""Processors: the reference's ``smol.moca.processor`` interface evaluated on the GPU.

Same class names, constructor arguments, attributes and error behaviour as
``smol/moca/processor/{base,expansion,ewald,composite}.py``; the arithmetic runs in the
CUDA kernels of ``liblmc`` (no CPU path).  Single-occupancy methods do a host round trip
(they exist for API compatibility and tests); the batched ``*_batch`` variants are the
efficient form.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from .model import ExpansionTables, PackedModel
from .sublattice import Sublattice


# Pauling electronegativities (pymatgen's Element.X; elements without a value compare as +inf there)
_PAULING_X = dict(
    H=2.20, Li=0.98, Be=1.57, B=2.04, C=2.55, N=3.04, O=3.44, F=3.98, Na=0.93, Mg=1.31, Al=1.61, Si=1.90, P=2.19,
    S=2.58, Cl=3.16, K=0.82, Ca=1.00, Sc=1.36, Ti=1.54, V=1.63, Cr=1.66, Mn=1.55, Fe=1.83, Co=1.88, Ni=1.91, Cu=1.90,
    Zn=1.65, Ga=1.81, Ge=2.01, As=2.18, Se=2.55, Br=2.96, Kr=3.00, Rb=0.82, Sr=0.95, Y=1.22, Zr=1.33, Nb=1.6, Mo=2.16,
    Tc=1.9, Ru=2.2, Rh=2.28, Pd=2.20, Ag=1.93, Cd=1.69, In=1.78, Sn=1.96, Sb=2.05, Te=2.1, I=2.66, Xe=2.6, Cs=0.79,
    Ba=0.89, La=1.10, Ce=1.12, Pr=1.13, Nd=1.14, Pm=1.13, Sm=1.17, Eu=1.2, Gd=1.2, Tb=1.1, Dy=1.22, Ho=1.23, Er=1.24,
    Tm=1.25, Yb=1.1, Lu=1.27, Hf=1.3, Ta=1.5, W=2.36, Re=1.9, Os=2.2, Ir=2.20, Pt=2.28, Au=2.54, Hg=2.00, Tl=1.62,
    Pb=2.33, Bi=2.02, Th=1.3, U=1.38)


def species_sort_key(label):
    """Ordering key of a species label ('Li+', 'Mn3+', 'O2-', 'Au', a pymatgen Species ...) equal to pymatgen's
    ``Species.__lt__``: electronegativity (unknown = +inf), symbol, oxidation state; labels that are not element
    symbols sort by the label itself behind every element of known electronegativity."""
    import re
    text = str(label)
    m = re.match(r"^([A-Z][a-z]?)(\d*\.?\d*)([+-]?)$", text)
    if not m or m.group(1) not in _PAULING_X:
        return (float("inf"), text, 0.0)
    oxi = float(m.group(2) or 1.0) * (1 if m.group(3) == "+" else -1) if m.group(3) else 0.0
    return (_PAULING_X[m.group(1)], m.group(1), oxi)


def _reject_use_concentration(flag):
    """base.py:60-62: the site-basis measure comes from the prim's site concentrations.  The bases (and so the
    correlation / interaction tensors) are inputs here; a subspace built with the concentration measure already
    carries them, the flag itself cannot be honoured after the fact."""
    if flag:
        raise NotImplementedError("use_concentration=True is not supported: build the cluster subspace (site bases) "
                                  "with the concentration measure instead")


def _as_occu(occupancy):
    """int32 conversion with the reference's error (expansion.py:176-189, 210-214)."""
    try:
        return np.array(occupancy, dtype=np.int32)
    except (ValueError, TypeError):
        types = {type(n) for n in occupancy}
        raise ValueError(f"occupancy contains {types}, but should be integers!")


class Processor(ABC):
    """``smol/moca/processor/base.py:26-312``."""

    def __init__(self, cluster_subspace, supercell_matrix, coefficients=None):
        self._subspace = cluster_subspace
        self._scmatrix = np.array(supercell_matrix)
        self.size = int(round(abs(np.linalg.det(self._scmatrix))))          # base.py:66
        self.coefs = None if coefficients is None else np.array(coefficients, dtype=np.float64)
        # species allowed on every supercell site, in the order that defines the occupancy codes.  The reference
        # reads them off the supercell structure (base.py:68-72, get_allowed_species); a subspace from
        # smol_b200.lattice provides them directly, and smol_b200.interop hands over a live smol processor's list
        # through the attribute ``_allowed_species_override`` of the subspace wrapper.
        self.allowed_species = list(cluster_subspace.allowed_species(self._scmatrix))
        self.num_sites = len(self.allowed_species)
        self._engine = None

    @property
    def cluster_subspace(self):
        return self._subspace

    @property
    def supercell_matrix(self):
        return self._scmatrix

    # ---- encoding (base.py:192-243) ------------------------------------------------------
    def encode_occupancy(self, occupancy):
        return np.array([species.index(sp) for species, sp in zip(self.allowed_species, occupancy)],
                        dtype=np.int32)

    def decode_occupancy(self, encoded_occupancy):
        return [species[i] for i, species in zip(encoded_occupancy, self.allowed_species)]

    def get_sublattices(self):
        """base.py:75-82, 245-268: one sublattice per unique site space, in the reference's FIXED order -- the site
        spaces sorted by their species lists (``SiteSpace.__lt__``, cofe/space/domain.py:201-203), species compared as
        pymatgen compares them (electronegativity, then symbol, then oxidation state).  Positional arguments such as
        ``sublattice_probabilities``, the table-flip dimensions and hyperplane columns follow this order."""
        spaces = []
        for sp in self.allowed_species:
            if sp not in spaces:
                spaces.append(sp)
        spaces.sort(key=lambda space: [species_sort_key(sp) for sp in space])
        return [Sublattice(space, np.array([i for i, sp in enumerate(self.allowed_species)
                                            if sp == space])) for space in spaces]

    # ---- tables ----------------------------------------------------------------------------
    @abstractmethod
    def _tables(self):
        """Return dict(expansion=ExpansionTables|None, ewald=(matrix, inds)|None)."""

    def _get_engine(self):
        if self._engine is None:
            from .engine import LmcEngine
            packed = PackedModel(self.num_sites, self.coefs, self.get_sublattices(), **self._tables())
            self._engine = LmcEngine(packed)
        return self._engine

    # ---- evaluation --------------------------------------------------------------------------
    def compute_feature_vector_batch(self, occupancies):
        """``[W, N]`` int occupancies -> ``[W, F]`` float64 feature vectors (GPU)."""
        eng = self._get_engine()
        occ = eng.upload_occupancy(_as_occu(occupancies).reshape(-1, self.num_sites))
        feat, _ = eng.full_features(occ)
        return feat.cpu().numpy()

    def compute_feature_vector_change_batch(self, occupancies, sites, codes):
        """Batched flips: ``sites``/``codes`` ``[W, k]`` applied sequentially per walker."""
        import torch
        eng = self._get_engine()
        occ = eng.upload_occupancy(_as_occu(occupancies).reshape(-1, self.num_sites))
        s = torch.as_tensor(np.ascontiguousarray(sites, dtype=np.int32)).to(eng.device)
        c = torch.as_tensor(np.ascontiguousarray(codes, dtype=np.int32)).to(eng.device)
        if s.ndim == 1:
            s, c = s[:, None].contiguous(), c[:, None].contiguous()
        return eng.delta_features(occ, s, c).cpu().numpy()

    def compute_feature_vector(self, occupancy):
        out = self.compute_feature_vector_batch(_as_occu(occupancy)[None, :])[0]
        return out if out.size > 1 or not self._scalar_feature else float(out[0])

    def compute_feature_vector_change(self, occupancy, flips):
        occu = _as_occu(occupancy)
        nfeat = len(self.coefs)
        if len(flips) == 0:
            out = np.zeros(nfeat)
        else:
            sites = np.array([[f[0] for f in flips]], dtype=np.int32)
            codes = np.array([[f[1] for f in flips]], dtype=np.int32)
            out = self.compute_feature_vector_change_batch(occu[None, :], sites, codes)[0]
        return out if not self._scalar_feature else float(out[0])

    _scalar_feature = False

    def compute_property(self, occupancy):
        """base.py:165-176."""
        return np.dot(self.coefs, self.compute_feature_vector(occupancy))

    def compute_property_change(self, occupancy, flips):
        """base.py:178-190."""
        return np.dot(self.coefs, self.compute_feature_vector_change(occupancy, flips))

    def compute_average_drift(self, iterations=1000, seed=None):
        """base.py:270-312 (delta vs full, forward and reverse), batched on the GPU."""
        rng = np.random.default_rng(seed)
        occu = np.array([rng.integers(len(sp)) for sp in self.allowed_species], dtype=np.int32)
        active = [i for i, sp in enumerate(self.allowed_species) if len(sp) > 1]
        occs, sites, codes = [], [], []
        for _ in range(iterations):
            site = int(rng.choice(active))
            choices = [c for c in range(len(self.allowed_species[site])) if c != occu[site]]
            new = int(rng.choice(choices))
            occs.append(occu.copy())
            sites.append(site)
            codes.append(new)
            occu[site] = new
        occs.append(occu.copy())
        occs = np.array(occs)
        coefs = np.atleast_1d(self.coefs)
        props = self.compute_feature_vector_batch(occs) @ coefs
        sites_a, codes_a = np.array(sites)[:, None], np.array(codes)[:, None]
        dfw = self.compute_feature_vector_change_batch(occs[:-1], sites_a, codes_a) @ coefs
        back = occs[:-1][np.arange(iterations), sites][:, None]
        drv = self.compute_feature_vector_change_batch(occs[1:], sites_a, back) @ coefs
        forward = np.sum((props[1:] - props[:-1]) - dfw) / iterations
        reverse = np.sum((props[:-1] - props[1:]) - drv) / iterations
        return forward, reverse


class ClusterExpansionProcessor(Processor):
    """``smol/moca/processor/expansion.py:39-241``: features = correlation vector * size."""

    def __init__(self, cluster_subspace, supercell_matrix, coefficients, use_concentration=False,
                 num_threads=None, num_threads_full=None):
        _reject_use_concentration(use_concentration)
        super().__init__(cluster_subspace, supercell_matrix, coefficients)
        if len(coefficients) != cluster_subspace.num_corr_functions:
            raise ValueError(
                f"The provided coefficients are not the right length. Got {len(coefficients)} "
                f"coefficients, the length must be {cluster_subspace.num_corr_functions} based on "
                f"the provided cluster subspace.")
        self.num_threads = num_threads            # kept for API compatibility (OpenMP knob)
        self.num_threads_full = num_threads_full
        self._exp = None

    def _tables(self):
        if self._exp is None:
            self._exp = ExpansionTables(self._subspace, self._scmatrix, "correlation",
                                        num_sites=self.num_sites)
        return dict(expansion=self._exp)


class ClusterDecompositionProcessor(Processor):
    """``expansion.py:243-489``: features = mean cluster interactions * size."""

    def __init__(self, cluster_subspace, supercell_matrix, interaction_tensors, coefficients=None,
                 use_concentration=False, num_threads=None, num_threads_full=None):
        if len(interaction_tensors) != cluster_subspace.num_orbits:
            raise ValueError(
                f"The number of cluster interaction tensors must match the number  of orbits in "
                f"the subspace. Got {len(interaction_tensors)} interaction tensors, but need "
                f"{cluster_subspace.num_orbits}  for the given cluster_subspace.")
        _reject_use_concentration(use_concentration)
        coefficients = (cluster_subspace.orbit_multiplicities if coefficients is None
                        else coefficients)                                   # expansion.py:311-316
        super().__init__(cluster_subspace, supercell_matrix, coefficients)
        self._interaction_tensors = interaction_tensors
        self.num_threads = num_threads
        self.num_threads_full = num_threads_full
        self._exp = None

    def _tables(self):
        if self._exp is None:
            self._exp = ExpansionTables(self._subspace, self._scmatrix, "interaction",
                                        interaction_tensors=self._interaction_tensors,
                                        num_sites=self.num_sites)
        return dict(expansion=self._exp)


def _orbits_by_diameter(cluster_subspace):
    """``clusterspace.py:367-381``: orbits grouped by diameter (rounded to 6 decimals), ascending."""
    obd = getattr(cluster_subspace, "orbits_by_diameter", None)
    if obd is not None:
        return dict(obd)
    from itertools import groupby

    def diam(orb):
        base = getattr(orb, "base_cluster", None)
        d = getattr(base, "diameter", None)
        return float(np.round(orb.diameter if d is None else d, 6))
    return {size: tuple(orbs) for size, orbs in groupby(sorted(cluster_subspace.orbits, key=diam), key=diam)}


class DistanceProcessor:
    """Mixin of ``smol/moca/processor/distance.py:20-200``: features ``[L, |f_i - f_T,i| ...]`` with ``f`` the
    correlation / cluster-interaction vector PER SUPERCELL, ``L`` the largest orbit diameter up to which all
    features match the target within ``match_tol``, coefficients ``[-match_weight, target_weights ...]``.

    On the device the running vector ``f`` is kept per walker and every proposal folds its per-record differences
    into the change of ``f`` (``LmcRunConfig.dist_*``); the reference re-evaluates both full vectors per proposal
    (``evaluator.pyx:319-435``)."""

    def _init_distance(self, target_vector, match_weight, match_tol, target_weights, feature_indices):
        if match_weight < 0:                                                     # distance.py:80-81
            raise ValueError("The match weight must be a positive number.")
        if len(target_weights) != len(target_vector) - 1:                       # distance.py:83-88
            raise ValueError(f"The length of target_weights must be equal to the length ofthe target vector minus "
                             f"one {len(target_vector) - 1}. \nGot {len(target_weights)} instead.")
        self.target_vector = np.array(target_vector, dtype=np.float64)
        self.match_tol = float(match_tol)
        self.coefs = np.concatenate([[-float(match_weight)], np.asarray(target_weights, dtype=np.float64)])
        groups = _orbits_by_diameter(self._subspace)
        self._group_diam = np.array(list(groups.keys()), dtype=np.float64)
        idx = [np.array(feature_indices(orbs), dtype=np.int32) for orbs in groups.values()]
        self._group_off = np.concatenate([[0], np.cumsum([len(i) for i in idx])]).astype(np.int32)
        self._group_idx = np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, dtype=np.int32)

    def distance_tables(self):
        """(target, tol, group offsets, group feature indices, group diameters) for ``LmcRunConfig.dist_*``."""
        return self.target_vector, self.match_tol, self._group_off, self._group_idx, self._group_diam

    def exact_match_max_diameter(self, distance_vector):
        """``distance.py:309-331, 452-472`` (host helper for a vector already on the host)."""
        out = 0.0
        for q, diam in enumerate(self._group_diam):
            ids = self._group_idx[self._group_off[q]:self._group_off[q + 1]]
            if np.all(np.asarray(distance_vector)[ids] <= self.match_tol):
                out = float(diam)
            else:
                break
        return out

    def compute_feature_vector_batch(self, occupancies):
        """``[W, N]`` occupancies -> ``[W, F]`` distance vectors (``distance.py:133-154``) on the device."""
        eng = self._get_engine()
        occ = eng.upload_occupancy(_as_occu(occupancies).reshape(-1, self.num_sites))
        feat, _ = eng.full_features(occ)
        eng.distance_init(self, feat)
        return feat.cpu().numpy()

    def compute_feature_vector_change_batch(self, occupancies, sites, codes):
        """distance vectors with the flips applied minus the current ones (``distance.py:156-180``)"""
        occ = _as_occu(occupancies).reshape(-1, self.num_sites)
        sites, codes = np.atleast_2d(sites), np.atleast_2d(codes)
        nxt = occ.copy()
        for k in range(sites.shape[1]):
            nxt[np.arange(len(occ)), sites[:, k]] = codes[:, k]
        both = self.compute_feature_vector_batch(np.concatenate([occ, nxt]))
        return both[len(occ):] - both[:len(occ)]


class CorrelationDistanceProcessor(DistanceProcessor, ClusterExpansionProcessor):
    """``distance.py:209-331``."""

    def __init__(self, cluster_subspace, supercell_matrix, target_vector=None, match_weight=1.0, match_tol=1e-8,
                 target_weights=None, use_concentration=False, **processor_kwargs):
        n = cluster_subspace.num_corr_functions
        target_vector = np.zeros(n) if target_vector is None else target_vector          # distance.py:262-264
        target_weights = np.ones(n - 1) if target_weights is None else target_weights
        ClusterExpansionProcessor.__init__(self, cluster_subspace, supercell_matrix, np.zeros(n),
                                           use_concentration=use_concentration, **processor_kwargs)
        self._init_distance(target_vector, match_weight, match_tol, target_weights,
                            lambda orbs: [i for o in orbs for i in range(o.bit_id, o.bit_id + len(o))])


class ClusterInteractionDistanceProcessor(DistanceProcessor, ClusterDecompositionProcessor):
    """``distance.py:334-472``."""

    def __init__(self, cluster_subspace, supercell_matrix, interaction_tensors, target_vector=None,
                 match_weight=1.0, match_tol=1e-8, target_weights=None, use_concentration=False,
                 **processor_kwargs):
        n = cluster_subspace.num_orbits
        target_vector = np.zeros(n) if target_vector is None else target_vector
        target_weights = np.ones(n - 1) if target_weights is None else target_weights
        ClusterDecompositionProcessor.__init__(self, cluster_subspace, supercell_matrix, interaction_tensors,
                                               use_concentration=use_concentration, **processor_kwargs)
        self._init_distance(target_vector, match_weight, match_tol, target_weights, lambda orbs: [o.id for o in orbs])


class EwaldProcessor(Processor):
    """``smol/moca/processor/ewald.py:26-208``.

    The reference builds the matrix with pymatgen's ``EwaldSummation``; here it is either given
    (``ewald_matrix`` + ``ewald_inds`` -- e.g. extracted from a live smol ``EwaldProcessor``) or
    computed by ``smol_b200.lattice.ewald_matrix`` for the same overlaid-species structure.
    """

    _scalar_feature = True

    def __init__(self, cluster_subspace, supercell_matrix, ewald_term=None, coefficient=1.0,
                 ewald_matrix=None, ewald_inds=None):
        super().__init__(cluster_subspace, supercell_matrix, np.array([coefficient]))
        self._ewald_term = ewald_term
        if ewald_matrix is None:
            from .lattice import ewald_matrix as _ewm
            kw = {}
            if ewald_term is not None:      # EwaldTerm(eta, real_space_cut, recip_space_cut, use_term), cofe/extern/ewald.py:35-62
                kw = dict(eta=getattr(ewald_term, "eta", None),
                          real_cut=getattr(ewald_term, "real_space_cut", None),
                          recip_cut=getattr(ewald_term, "recip_space_cut", None),
                          term=getattr(ewald_term, "use_term", "total"))
            ewald_matrix, ewald_inds = _ewm(cluster_subspace, self._scmatrix, **kw)
        self._matrix = np.ascontiguousarray(ewald_matrix, dtype=np.float64)
        self._ewald_inds = np.ascontiguousarray(ewald_inds, dtype=np.int32)

    @property
    def ewald_matrix(self):
        return self._matrix

    def _tables(self):
        return dict(ewald=(self._matrix, self._ewald_inds))

    def compute_property(self, occupancy):
        return self.coefs * self.compute_feature_vector(occupancy)            # ewald.py:103-113

    def compute_property_change(self, occupancy, flips):
        return self.coefs * self.compute_feature_vector_change(occupancy, flips)


class CompositeProcessor(Processor):
    """``smol/moca/processor/composite.py:26-180``: concatenated features of its members.

    Supported composition (what the reference builds in ``Ensemble.from_cluster_expansion``,
    ensemble.py:180-214): one expansion-type processor optionally followed by one
    ``EwaldProcessor``.
    """

    def __init__(self, cluster_subspace, supercell_matrix, use_concentration=False):
        _reject_use_concentration(use_concentration)
        super().__init__(cluster_subspace, supercell_matrix, None)
        self._processors = []
        self.coefs = np.empty(0)

    @property
    def processors(self):
        return self._processors

    def add_processor(self, processor):
        if isinstance(processor, CompositeProcessor):
            raise AttributeError("A CompositeProcessor can not be added into another "
                                 "CompositeProcessor")
        if self.cluster_subspace is not processor.cluster_subspace and \
                self.cluster_subspace != processor.cluster_subspace:
            raise ValueError("The cluster subspace of the processor to be added does not match "
                             "the one of this CompositeProcessor.")
        if not np.array_equal(self._scmatrix, processor.supercell_matrix):
            raise ValueError("The supercell matrix of the processor to be added does not match "
                             "the one of this CompositeProcessor.")
        if isinstance(processor, EwaldProcessor):
            if any(isinstance(p, EwaldProcessor) for p in self._processors):
                raise NotImplementedError("only one EwaldProcessor per composite is supported")
        elif self._processors:
            raise NotImplementedError("the expansion-type processor must be added first and only once")
        self._processors.append(processor)
        self.coefs = np.append(self.coefs, processor.coefs)
        self._engine = None

    def _tables(self):
        out = {}
        for p in self._processors:
            out.update(p._tables())
        return out

    def compute_property(self, occupancy):
        return float(np.dot(self.coefs, self.compute_feature_vector(occupancy)))

    def compute_property_change(self, occupancy, flips):
        return float(np.dot(self.coefs, self.compute_feature_vector_change(occupancy, flips)))
