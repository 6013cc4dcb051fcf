"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Plain Python/numpy restatement of smol's lattice-MC hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this
module; the product (``smol_b200``) never does.

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  Arithmetic is done in the reference's order (orbit -> bit
combo -> cluster, sequential double adds) so that it can be compared bit-for-bit
with the C restatement ``oracle/lmc_oracle.c`` and, within the ``-ffast-math``
slack of the reference build, with the reference's own compiled evaluators in
``oracle/_ref`` (pass ``use_ref=True``).

Parity pinning (see tests/test_oracle_*.py):
  * against the compiled reference evaluators (oracle/_ref, built here from
    /root/reference by oracle/build_ref.py) on seeded random inputs, and the
    golden vectors generated from them in tests/golden/ (script committed);
  * against the reference's known-answer table-flip a-priori factors
    (tests/test_moca/test_mcushers.py:199-234) and the correlation vectors stored in its tests
    (LiCaBr and the CASM-generated ones, tests/test_cofe/test_clusterspace.py:627-725, 881-996);
  * against the reference's OWN Python classes, imported unmodified from /root/reference behind package shells
    (tests/golden/make_reference_python_golden.py -> tests/golden/ref_python_steps.npz): Metropolis /
    UniformlyRandom / WangLandau / MulticellMetropolis kernels, Flip / Swap / TableFlip / Composite / MultiStep
    ushers, the three bias terms, the Sampler loop and SampleContainer, the expansion / decomposition / distance
    processors -- step by step with the kernels' generator scripted to the Philox word positions below.

RNG: the reference uses numpy PCG64 with a data-dependent number of draws per step
(``kernel/mcusher.py:146-200``), which cannot be reproduced lane-wise on a GPU.  The
engine and this oracle share a counter-based Philox4x32-10 stream instead:
``key = (seed_lo, seed_hi)`` per walker, ``counter = (step_lo, step_hi, block,
walker_id)``; see ``StepRandom`` for the fixed meaning of every 32-bit word.
"""
from __future__ import annotations

import math
import os
import sys
from types import SimpleNamespace

import numpy as np

kB = 8.617333262145e-5  # smol/constants.py:4

# ======================================================================================
# Philox4x32-10 (Salmon et al., SC'11) -- counter based RNG shared with the CUDA kernels
# ======================================================================================
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Return 4 uint32 words for ``counter`` (4 words) and ``key`` (2 words)."""
    c0, c1, c2, c3 = (int(c) & _MASK for c in counter)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK
        hi1, lo1 = p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def mulhi32(r: int, n: int) -> int:
    """Bounded integer in [0, n): high word of the 32x32 product."""
    return (int(r) * int(n)) >> 32


def u01(r: int) -> float:
    """Uniform double in (0, 1) from one 32-bit word: (r + 0.5) * 2^-32."""
    return (int(r) + 0.5) * 2.0 ** -32


class StepRandom:
    """Random words of ONE attempted step of ONE walker.

    block 0: [r0 sublattice choice | r1 first site | r2 second site rank / new code | r3 accept]
             (TableFlip: r0 = swap-vs-table decision, r1 = flip direction choice, r3 = accept)
    block 1+: TableFlip only -- either the (sublattice, site1, rank) words of the fallback
             swap (block 1) or one word per sequential site pick (block 1 + pick//4, pick%4).
    """

    def __init__(self, seed: int, walker: int, step: int):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.walker = walker
        self.step = step
        self._blocks = {}

    def block(self, b: int):
        if b not in self._blocks:
            self._blocks[b] = philox4x32_10(
                (self.step & _MASK, (self.step >> 32) & _MASK, b, self.walker), self.key)
        return self._blocks[b]

    def word(self, i: int) -> int:
        return self.block(i // 4)[i % 4]


# ======================================================================================
# Reference compiled evaluators (oracle/_ref), optional
# ======================================================================================
def load_ref():
    """Import the reference's compiled Cython modules from oracle/_ref or return None."""
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if not os.path.isdir(os.path.join(ref, "smol")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        from smol.utils.cluster.container import IntArray2DContainer
        from smol.utils.cluster.evaluator import ClusterSpaceEvaluator
        from smol.utils.cluster.ewald import delta_ewald_single_flip
    except Exception:  # pragma: no cover
        return None
    return SimpleNamespace(ClusterSpaceEvaluator=ClusterSpaceEvaluator,
                           IntArray2DContainer=IntArray2DContainer,
                           delta_ewald_single_flip=delta_ewald_single_flip)


# ======================================================================================
# L1: evaluator restatements (smol/utils/cluster/evaluator.pyx, ewald.pyx)
# ======================================================================================
def _seq_sum(values) -> float:
    p = 0.0
    for v in values:
        p = p + float(v)
    return p


def correlations_from_occupancy(orbit_data, num_corr, occu, indices):
    """evaluator.pyx:121-168."""
    out = np.zeros(num_corr)
    out[0] = 1.0
    for (oid, bit_id, tensors, strides), idx in zip(orbit_data, indices):
        J = idx.shape[0]
        ind = (occu[idx] * strides[None, :]).sum(axis=1)
        for k in range(tensors.shape[0]):
            out[bit_id + k] = _seq_sum(tensors[k, ind]) / J
    return out


def interactions_from_occupancy(orbit_data, num_orbits, offset, inter_tensors, occu, indices):
    """evaluator.pyx:170-209."""
    out = np.zeros(num_orbits)
    out[0] = offset
    for (oid, bit_id, tensors, strides), idx, inter in zip(orbit_data, indices, inter_tensors):
        J = idx.shape[0]
        ind = (occu[idx] * strides[None, :]).sum(axis=1)
        out[oid] = _seq_sum(inter[ind]) / J
    return out


def delta_correlations_from_occupancies(orbit_data, num_corr, occu_f, occu_i, ratio, indices):
    """evaluator.pyx:211-265: out[bit_id+k] = (sum_j T[k,ind_f]-T[k,ind_i]) / ratio / J."""
    out = np.zeros(num_corr)
    for n, ((oid, bit_id, tensors, strides), idx) in enumerate(zip(orbit_data, indices)):
        J = idx.shape[0]
        ind_i = (occu_i[idx] * strides[None, :]).sum(axis=1)
        ind_f = (occu_f[idx] * strides[None, :]).sum(axis=1)
        for k in range(tensors.shape[0]):
            p = _seq_sum(tensors[k, ind_f] - tensors[k, ind_i])
            out[bit_id + k] = p / ratio[n] / J
    return out


def delta_interactions_from_occupancies(orbit_data, num_orbits, inter_tensors, occu_f, occu_i,
                                        ratio, indices):
    """evaluator.pyx:267-317."""
    out = np.zeros(num_orbits)
    for n, ((oid, bit_id, tensors, strides), idx, inter) in enumerate(
            zip(orbit_data, indices, inter_tensors)):
        J = idx.shape[0]
        ind_i = (occu_i[idx] * strides[None, :]).sum(axis=1)
        ind_f = (occu_f[idx] * strides[None, :]).sum(axis=1)
        p = _seq_sum(inter[ind_f] - inter[ind_i])
        out[oid] = p / ratio[n] / J
    return out


def corr_distances_from_occupancies(orbit_data, num_corr, occu_f, occu_i, ref_corr_vector, indices):
    """evaluator.pyx:319-372: |corr - ref| of the two occupancies, rows [before, after]; column 0 stays 0."""
    out = np.zeros((2, num_corr))
    for (oid, bit_id, tensors, strides), idx in zip(orbit_data, indices):
        J = idx.shape[0]
        ind_i = (occu_i[idx] * strides[None, :]).sum(axis=1)
        ind_f = (occu_f[idx] * strides[None, :]).sum(axis=1)
        for k in range(tensors.shape[0]):
            out[1, bit_id + k] = abs(_seq_sum(tensors[k, ind_f]) / J - ref_corr_vector[bit_id + k])
            out[0, bit_id + k] = abs(_seq_sum(tensors[k, ind_i]) / J - ref_corr_vector[bit_id + k])
    return out


def interaction_distances_from_occupancies(orbit_data, num_orbits, inter_tensors, occu_f, occu_i,
                                           ref_interaction_vector, indices):
    """evaluator.pyx:374-435."""
    out = np.zeros((2, num_orbits))
    for (oid, bit_id, tensors, strides), idx, inter in zip(orbit_data, indices, inter_tensors):
        J = idx.shape[0]
        ind_i = (occu_i[idx] * strides[None, :]).sum(axis=1)
        ind_f = (occu_f[idx] * strides[None, :]).sum(axis=1)
        out[1, oid] = abs(_seq_sum(inter[ind_f]) / J - ref_interaction_vector[oid])
        out[0, oid] = abs(_seq_sum(inter[ind_i]) / J - ref_interaction_vector[oid])
    return out


def delta_ewald_single_flip(occu_f, occu_i, ewald_matrix, ewald_indices, site_ind):
    """ewald.pyx:9-59 (sequential k loop, per-k partial ``out_k`` then ``out += out_k``)."""
    add = ewald_indices[site_ind, occu_f[site_ind]]
    sub = ewald_indices[site_ind, occu_i[site_ind]]
    n = occu_f.shape[0]
    rows = np.arange(n)
    i = ewald_indices[rows, occu_f]
    j = ewald_indices[rows, occu_i]
    out_k = np.zeros(n)
    if add != -1:
        ok = i != -1
        v = ewald_matrix[i[ok], add]
        out_k[ok] += np.where(i[ok] != add, 2 * v, v)
    if sub != -1:
        ok = j != -1
        v = ewald_matrix[j[ok], sub]
        out_k[ok] -= np.where(j[ok] != sub, 2 * v, v)
    return _seq_sum(out_k)


# ======================================================================================
# L2: processors (smol/moca/processor/*.py)
# ======================================================================================
def get_orbit_data(orbits):
    """utils/cluster/__init__.py:4-15."""
    return tuple((o.id, o.bit_id, np.ascontiguousarray(o.flat_correlation_tensors),
                  np.ascontiguousarray(o.flat_tensor_indices, dtype=np.int32)) for o in orbits)


class _LocalData(SimpleNamespace):
    pass


class _ExpansionBase:
    """Shared table construction of expansion.py:120-156 / 344-392."""

    def __init__(self, cluster_subspace, supercell_matrix, coefficients, use_ref=False):
        self.cluster_subspace = cluster_subspace
        self.supercell_matrix = np.asarray(supercell_matrix)
        self.size = int(round(abs(np.linalg.det(self.supercell_matrix))))  # base.py:66
        self.coefs = np.asarray(coefficients, dtype=np.float64)
        self._orbit_data = get_orbit_data(cluster_subspace.orbits)
        self._indices = tuple(cluster_subspace.get_orbit_indices(supercell_matrix).arrays)
        # Processor.num_sites = len(supercell structure) (processor/base.py:88-91), which also counts
        # sites without configurational freedom that appear in no cluster
        nss = getattr(cluster_subspace, "num_supercell_sites", None)
        self.num_sites = (int(nss(self.supercell_matrix)) if nss is not None
                          else 1 + max(int(a.max()) for a in self._indices))
        self._ref = load_ref() if use_ref else None
        if use_ref and self._ref is None:
            raise RuntimeError("oracle/_ref is not built; run python oracle/build_ref.py")
        data_by_sites = {}
        for n, (odata, cluster_indices) in enumerate(zip(self._orbit_data, self._indices)):
            for site_ind in np.unique(cluster_indices):
                in_inds = np.any(cluster_indices == site_ind, axis=-1)
                ratio = len(cluster_indices) / np.sum(in_inds)
                data_by_sites.setdefault(int(site_ind), []).append(
                    (n, odata, np.ascontiguousarray(cluster_indices[in_inds]), ratio))
        self._data_by_sites = data_by_sites
        self._local = {}

    def _local_data(self, site):
        if site not in self._local:
            data = self._data_by_sites[site]
            ld = _LocalData(orbit_ns=[d[0] for d in data], orbit_data=tuple(d[1] for d in data),
                            indices=tuple(d[2] for d in data),
                            ratio=np.array([d[3] for d in data]))
            self._local[site] = ld
        return self._local[site]


class ClusterExpansionProcessor(_ExpansionBase):
    """processor/expansion.py:39-241 (correlation-vector features)."""

    def __init__(self, cluster_subspace, supercell_matrix, coefficients, use_ref=False):
        super().__init__(cluster_subspace, supercell_matrix, coefficients, use_ref)
        self.num_corr = cluster_subspace.num_corr_functions
        self.num_orbits = cluster_subspace.num_orbits
        if len(self.coefs) != self.num_corr:
            raise ValueError("The provided coefficients are not the right length.")
        if self._ref is not None:
            R = self._ref
            self._evaluator = R.ClusterSpaceEvaluator(self._orbit_data, self.num_orbits,
                                                      self.num_corr)
            self._container = R.IntArray2DContainer(self._indices)

    def num_features(self):
        return self.num_corr

    def compute_feature_vector(self, occupancy):
        """expansion.py:165-189: corr * size."""
        occupancy = np.array(occupancy, dtype=np.int32)
        if self._ref is not None:
            return self._evaluator.correlations_from_occupancy(occupancy, self._container) \
                * self.size
        return correlations_from_occupancy(self._orbit_data, self.num_corr, occupancy,
                                           self._indices) * self.size

    def compute_feature_vector_change(self, occupancy, flips):
        """expansion.py:191-231: flips applied sequentially, result * size."""
        occu_i = np.array(occupancy, dtype=np.int32)
        delta = np.zeros(self.num_corr)
        for f in flips:
            occu_f = occu_i.copy()
            occu_f[f[0]] = f[1]
            ld = self._local_data(int(f[0]))
            if self._ref is not None:
                if not hasattr(ld, "ev"):
                    R = self._ref
                    ld.ev = R.ClusterSpaceEvaluator(ld.orbit_data, self.num_orbits, self.num_corr)
                    ld.cont = R.IntArray2DContainer(ld.indices)
                delta += ld.ev.delta_correlations_from_occupancies(occu_f, occu_i, ld.ratio,
                                                                   ld.cont)
            else:
                delta += delta_correlations_from_occupancies(
                    ld.orbit_data, self.num_corr, occu_f, occu_i, ld.ratio, ld.indices)
            occu_i = occu_f
        return delta * self.size

    def compute_property(self, occupancy):
        return np.dot(self.coefs, self.compute_feature_vector(occupancy))  # base.py:176

    def compute_property_change(self, occupancy, flips):
        return np.dot(self.coefs, self.compute_feature_vector_change(occupancy, flips))


class ClusterDecompositionProcessor(_ExpansionBase):
    """processor/expansion.py:243-489 (cluster-interaction features, coefs = multiplicities)."""

    def __init__(self, cluster_subspace, supercell_matrix, interaction_tensors,
                 coefficients=None, use_ref=False):
        coefficients = (cluster_subspace.orbit_multiplicities if coefficients is None
                        else coefficients)
        super().__init__(cluster_subspace, supercell_matrix, coefficients, use_ref)
        self.num_corr = cluster_subspace.num_corr_functions
        self.num_orbits = cluster_subspace.num_orbits
        if len(interaction_tensors) != self.num_orbits:
            raise ValueError("The number of cluster interaction tensors must match orbits.")
        self.offset = float(interaction_tensors[0])
        self._flat = tuple(np.ravel(np.asarray(t, dtype=np.float64), order="C")
                           for t in interaction_tensors[1:])
        if self._ref is not None:
            R = self._ref
            self._evaluator = R.ClusterSpaceEvaluator(self._orbit_data, self.num_orbits,
                                                      self.num_corr, 1, self.offset, self._flat)
            self._container = R.IntArray2DContainer(self._indices)

    def num_features(self):
        return self.num_orbits

    def compute_feature_vector(self, occupancy):
        """expansion.py:392-418."""
        occupancy = np.array(occupancy, dtype=np.int32)
        if self._ref is not None:
            return self._evaluator.interactions_from_occupancy(occupancy, self._container) \
                * self.size
        return interactions_from_occupancy(self._orbit_data, self.num_orbits, self.offset,
                                           self._flat, occupancy, self._indices) * self.size

    def compute_feature_vector_change(self, occupancy, flips):
        """expansion.py:420-464."""
        occu_i = np.array(occupancy, dtype=np.int32)
        delta = np.zeros(self.num_orbits)
        for f in flips:
            occu_f = occu_i.copy()
            occu_f[f[0]] = f[1]
            ld = self._local_data(int(f[0]))
            flat = tuple(self._flat[n] for n in ld.orbit_ns)
            if self._ref is not None:
                if not hasattr(ld, "ev"):
                    R = self._ref
                    ld.ev = R.ClusterSpaceEvaluator(ld.orbit_data, self.num_orbits, self.num_corr,
                                                    1, self.offset, flat)
                    ld.cont = R.IntArray2DContainer(ld.indices)
                delta += ld.ev.delta_interactions_from_occupancies(occu_f, occu_i, ld.ratio,
                                                                   ld.cont)
            else:
                delta += delta_interactions_from_occupancies(
                    ld.orbit_data, self.num_orbits, flat, occu_f, occu_i, ld.ratio, ld.indices)
            occu_i = occu_f
        return delta * self.size

    def compute_property(self, occupancy):
        return np.dot(self.coefs, self.compute_feature_vector(occupancy))

    def compute_property_change(self, occupancy, flips):
        return np.dot(self.coefs, self.compute_feature_vector_change(occupancy, flips))


def orbits_by_diameter(cluster_subspace):
    """clusterspace.py:367-381: {diameter rounded to 6 decimals: orbits}, ascending."""
    from itertools import groupby

    def diam(orb):
        base = getattr(orb, "base_cluster", None)
        d = getattr(base, "diameter", None)
        return float(np.round(orb.diameter if d is None else d, 6))
    return {size: tuple(orbs) for size, orbs in groupby(sorted(cluster_subspace.orbits, key=diam), key=diam)}


class _DistanceMixin:
    """processor/distance.py:20-200: d = -w L + |W (f - f_T)|_1 as features [L, |f_i - f_T,i| ...] with
    coefficients [-w, W...]; features are per supercell (NOT multiplied by the size)."""

    def _init_distance(self, target_vector, match_weight, match_tol, target_weights):
        if match_weight < 0:
            raise ValueError("The match weight must be a positive number.")          # distance.py:80-81
        if len(target_weights) != len(target_vector) - 1:
            raise ValueError("The length of target_weights must be equal to the length of"
                             f"the target vector minus one {len(target_vector) - 1}.")
        self.target_vector = np.array(target_vector, dtype=np.float64)
        self.match_tol = match_tol
        self.coefs = np.concatenate([[-match_weight], np.asarray(target_weights, dtype=np.float64)])
        self._by_diameter = orbits_by_diameter(self.cluster_subspace)

    def exact_match_max_diameter(self, distance_vector):
        """distance.py:309-331 / 452-472."""
        max_matched_diameter = 0.0
        for diameter, orbits in self._by_diameter.items():
            indices = self._orbit_feature_indices(orbits)
            if np.all(distance_vector[indices] <= self.match_tol):
                max_matched_diameter = diameter
            else:
                break
        return max_matched_diameter

    def compute_feature_vector(self, occupancy):
        """distance.py:133-154."""
        occupancy = np.array(occupancy, dtype=np.int32)
        fv = self._base_feature_vector(occupancy) / self.size
        fv[:] = np.abs(fv - self.target_vector)
        fv[0] = self.exact_match_max_diameter(fv) if self.coefs[0] != 0 else 0.0
        return fv

    def compute_feature_vector_change(self, occupancy, flips):
        """distance.py:156-180."""
        occupancy = np.array(occupancy, dtype=np.int32)
        dv = self.compute_feature_vector_distances(occupancy, flips)
        if self.coefs[0] != 0:
            dv[0, 0] = self.exact_match_max_diameter(dv[0])
            dv[1, 0] = self.exact_match_max_diameter(dv[1])
        return dv[1] - dv[0]


class CorrelationDistanceProcessor(_DistanceMixin, ClusterExpansionProcessor):
    """processor/distance.py:209-331."""

    def __init__(self, cluster_subspace, supercell_matrix, target_vector=None, match_weight=1.0,
                 match_tol=1e-8, target_weights=None, use_ref=False):
        n = cluster_subspace.num_corr_functions
        target_vector = np.zeros(n) if target_vector is None else target_vector
        target_weights = np.ones(n - 1) if target_weights is None else target_weights
        ClusterExpansionProcessor.__init__(self, cluster_subspace, supercell_matrix, np.zeros(n), use_ref)
        self._init_distance(target_vector, match_weight, match_tol, target_weights)

    def _base_feature_vector(self, occupancy):
        return ClusterExpansionProcessor.compute_feature_vector(self, occupancy)

    def _orbit_feature_indices(self, orbits):
        return [i for orb in orbits for i in range(orb.bit_id, orb.bit_id + len(orb))]

    def compute_feature_vector_distances(self, occupancy, flips):
        """distance.py:281-307."""
        occu_f = occupancy.copy()
        for f in flips:
            occu_f[f[0]] = f[1]
        if self._ref is not None:
            return np.array(self._evaluator.corr_distances_from_occupancies(
                occu_f, occupancy, np.ascontiguousarray(self.target_vector), self._container))
        return corr_distances_from_occupancies(self._orbit_data, self.num_corr, occu_f, occupancy,
                                               self.target_vector, self._indices)

    def compute_property(self, occupancy):
        return np.dot(self.coefs, self.compute_feature_vector(occupancy))

    def compute_property_change(self, occupancy, flips):
        return np.dot(self.coefs, self.compute_feature_vector_change(occupancy, flips))


class ClusterInteractionDistanceProcessor(_DistanceMixin, ClusterDecompositionProcessor):
    """processor/distance.py:334-472."""

    def __init__(self, cluster_subspace, supercell_matrix, interaction_tensors, target_vector=None,
                 match_weight=1.0, match_tol=1e-8, target_weights=None, use_ref=False):
        n = cluster_subspace.num_orbits
        target_vector = np.zeros(n) if target_vector is None else target_vector
        target_weights = np.ones(n - 1) if target_weights is None else target_weights
        ClusterDecompositionProcessor.__init__(self, cluster_subspace, supercell_matrix, interaction_tensors,
                                               None, use_ref)
        self._init_distance(target_vector, match_weight, match_tol, target_weights)

    def _base_feature_vector(self, occupancy):
        return ClusterDecompositionProcessor.compute_feature_vector(self, occupancy)

    def _orbit_feature_indices(self, orbits):
        return [orb.id for orb in orbits]

    def compute_feature_vector_distances(self, occupancy, flips):
        """distance.py:424-450."""
        occu_f = occupancy.copy()
        for f in flips:
            occu_f[f[0]] = f[1]
        if self._ref is not None:
            return np.array(self._evaluator.interaction_distances_from_occupancies(
                occu_f, occupancy, np.ascontiguousarray(self.target_vector), self._container))
        return interaction_distances_from_occupancies(self._orbit_data, self.num_orbits, self._flat, occu_f,
                                                      occupancy, self.target_vector, self._indices)

    def compute_property(self, occupancy):
        return np.dot(self.coefs, self.compute_feature_vector(occupancy))

    def compute_property_change(self, occupancy, flips):
        return np.dot(self.coefs, self.compute_feature_vector_change(occupancy, flips))


class EwaldProcessor:
    """processor/ewald.py:26-208 with the matrix and index table given as inputs."""

    def __init__(self, ewald_matrix, ewald_inds, coefficient=1.0, use_ref=False):
        self.ewald_matrix = np.ascontiguousarray(ewald_matrix, dtype=np.float64)
        self._ewald_inds = np.ascontiguousarray(ewald_inds, dtype=np.int32)
        self.coefs = np.array([coefficient], dtype=np.float64)
        self.num_sites = self._ewald_inds.shape[0]
        self._ref = load_ref() if use_ref else None

    def num_features(self):
        return 1

    def compute_feature_vector(self, occupancy):
        """ewald.py:128-145 + cofe/extern/ewald.py:102-130."""
        occupancy = np.array(occupancy, dtype=np.int32)
        i_inds = self._ewald_inds[np.arange(len(occupancy)), occupancy]
        b_inds = np.zeros(self.ewald_matrix.shape[0] + 1, dtype=bool)
        b_inds[i_inds] = True
        ew = b_inds[:-1]
        return np.sum(self.ewald_matrix[ew, :][:, ew])

    def compute_feature_vector_change(self, occupancy, flips):
        """ewald.py:147-182."""
        occu_i = np.array(occupancy, dtype=np.int32)
        delta = 0
        for f in flips:
            occu_f = occu_i.copy()
            occu_f[f[0]] = f[1]
            if self._ref is not None:
                delta += self._ref.delta_ewald_single_flip(occu_f, occu_i, self.ewald_matrix,
                                                           self._ewald_inds, int(f[0]))
            else:
                delta += delta_ewald_single_flip(occu_f, occu_i, self.ewald_matrix,
                                                 self._ewald_inds, int(f[0]))
            occu_i = occu_f
        return delta

    def compute_property(self, occupancy):
        return self.coefs * self.compute_feature_vector(occupancy)

    def compute_property_change(self, occupancy, flips):
        return self.coefs * self.compute_feature_vector_change(occupancy, flips)


class CompositeProcessor:
    """processor/composite.py:26-180."""

    def __init__(self, processors=()):
        self._processors = []
        self.coefs = np.empty(0)
        for p in processors:
            self.add_processor(p)

    def add_processor(self, processor):
        self._processors.append(processor)
        self.coefs = np.append(self.coefs, processor.coefs)
        self.size = getattr(self._processors[0], "size", None)
        self.num_sites = self._processors[0].num_sites

    @property
    def processors(self):
        return self._processors

    def compute_feature_vector(self, occupancy):
        occupancy = np.array(occupancy, dtype=np.int32)
        feats = [np.array(p.compute_feature_vector(occupancy)) for p in self._processors]
        return np.append(feats[0], feats[1:])

    def compute_feature_vector_change(self, occupancy, flips):
        occupancy = np.array(occupancy, dtype=np.int32)
        ups = [np.array(p.compute_feature_vector_change(occupancy, flips))
               for p in self._processors]
        return np.append(ups[0], ups[1:])

    def compute_property(self, occupancy):
        return sum(p.compute_property(occupancy) for p in self._processors)

    def compute_property_change(self, occupancy, flips):
        return sum(p.compute_property_change(occupancy, flips) for p in self._processors)


# ======================================================================================
# L3: ensemble (smol/moca/ensemble.py) and sublattices (smol/moca/sublattice.py)
# ======================================================================================
class Sublattice(SimpleNamespace):
    """Plain-data sublattice: ``species`` (labels), ``sites``, ``active_sites``, ``encoding``."""

    def __init__(self, species, sites, active_sites=None, encoding=None):
        sites = np.unique(np.asarray(sites, dtype=np.int64))  # sublattice.py:55
        if active_sites is None:
            active_sites = sites.copy() if len(species) > 1 else np.array([], dtype=np.int64)
        if encoding is None:
            encoding = np.arange(len(species), dtype=np.int32)
        super().__init__(species=tuple(species), sites=sites,
                         active_sites=np.asarray(active_sites, dtype=np.int64),
                         encoding=np.asarray(encoding, dtype=np.int32))

    @property
    def is_active(self):
        return len(self.active_sites) > 0


class Ensemble:
    """ensemble.py:102-430 restricted to what the step loop reads."""

    def __init__(self, processor, sublattices, chemical_potentials=None):
        self.processor = processor
        self.sublattices = list(sublattices)
        self.num_sites = processor.num_sites
        self.natural_parameters = np.array(processor.coefs, dtype=np.float64)
        self.num_energy_coefs = len(processor.coefs)
        self.mu_table = None
        if chemical_potentials is not None:
            # ensemble.py:61-65, 89-99: table[site, code] = mu(species); natural parameter -1
            self.natural_parameters = np.append(self.natural_parameters, -1.0)
            num_cols = max(max(sl.encoding) for sl in self.sublattices) + 1
            table = np.zeros((self.num_sites, num_cols))
            for sl in self.active_sublattices:
                pots = [chemical_potentials[sp] for sp in sl.species]
                table[sl.sites[:, None], sl.encoding] = pots
            self.mu_table = table

    @property
    def active_sublattices(self):
        return [s for s in self.sublattices if s.is_active]

    def compute_feature_vector(self, occupancy):
        """ensemble.py:323-351."""
        feats = self.processor.compute_feature_vector(occupancy)
        if self.mu_table is not None:
            work = sum(self.mu_table[site][sp] for site, sp in enumerate(occupancy))
            feats = np.append(feats, work)
        return feats

    def compute_feature_vector_change(self, occupancy, step):
        """ensemble.py:353-376 (pre-step occupancy for every flip of the step)."""
        delta = self.processor.compute_feature_vector_change(occupancy, step)
        if self.mu_table is not None:
            dwork = sum(self.mu_table[f[0]][f[1]] - self.mu_table[f[0]][occupancy[f[0]]]
                        for f in step)
            delta = np.append(delta, dwork)
        return delta


# ======================================================================================
# L4: ushers (smol/moca/kernel/mcusher.py)
# ======================================================================================
class _Usher:
    def __init__(self, sublattices, sublattice_probabilities=None):
        self.sublattices = list(sublattices)
        self.active_sublattices = [s for s in self.sublattices if s.is_active]
        n = len(self.active_sublattices)
        probs = (np.full(n, 1.0 / n) if sublattice_probabilities is None
                 else np.asarray(sublattice_probabilities, dtype=np.float64))  # mcusher.py:61-75
        self._sublatt_probs = probs
        self._cum = np.cumsum(probs)
        self._cum[-1] = 1.0

    def get_random_sublattice(self, r0):
        """mcusher.py:146-148 with u = u01(r0): first s with cum[s] > u."""
        if len(self.active_sublattices) == 1:
            return self.active_sublattices[0]
        u = u01(r0)
        for s, c in enumerate(self._cum):
            if c > u:
                return self.active_sublattices[s]
        return self.active_sublattices[-1]

    def compute_log_priori_factor(self, occupancy, step):
        return 0.0


class Flip(_Usher):
    """mcusher.py:151-170."""

    def propose_step(self, occupancy, rnd: StepRandom, word0=0):
        r = [rnd.word(word0 + i) for i in range(3)]
        sl = self.get_random_sublattice(r[0])
        site = int(sl.active_sites[mulhi32(r[1], len(sl.active_sites))])
        choices = [int(c) for c in sl.encoding if c != occupancy[site]]  # encoding order
        return [(site, choices[mulhi32(r[2], len(choices))])]


class Swap(_Usher):
    """mcusher.py:173-200."""

    def propose_step(self, occupancy, rnd: StepRandom, word0=0):
        r = [rnd.word(word0 + i) for i in range(3)]
        sl = self.get_random_sublattice(r[0])
        site1 = int(sl.active_sites[mulhi32(r[1], len(sl.active_sites))])
        species1 = occupancy[site1]
        swap_options = sl.active_sites[occupancy[sl.active_sites] != species1]
        if swap_options.size > 0:
            site2 = int(swap_options[mulhi32(r[2], swap_options.size)])
            return [(site1, int(occupancy[site2])), (site2, int(species1))]
        return []


class Composite(_Usher):
    """mcusher.py:307-394: one of the sub-ushers, picked by weight with random word 4 of the step (the
    reference draws ``rng.choice(mcushers, p=p)``), proposes with its own sublattices / probabilities."""

    def __init__(self, sublattices, mcushers, mcusher_weights=None):
        super().__init__(sublattices)
        self.mcushers = list(mcushers)
        weights = [1] * len(self.mcushers) if mcusher_weights is None else list(mcusher_weights)
        total = sum(weights)                                           # mcusher.py:382-390
        self._p = [w / total for w in weights]
        self._pcum = np.cumsum(self._p)
        self._pcum[-1] = 1.0

    def propose_step(self, occupancy, rnd: StepRandom, word0=0):
        u = u01(rnd.word(4))
        pick = len(self.mcushers) - 1
        for i, c in enumerate(self._pcum):
            if c > u:
                pick = i
                break
        return self.mcushers[pick].propose_step(occupancy, rnd, word0)


class MultiStep(_Usher):
    """mcusher.py:203-304.  The length comes from random word 4 of the step (the reference draws
    ``rng.choice(step_lens, p=step_p)``); proposal j of the chain draws from words 8 + 4 j .. (block 2 + j)."""

    def __init__(self, sublattices, mcusher, step_lengths, step_probabilities=None):
        super().__init__(sublattices)
        self._mcusher = mcusher
        self._step_lens = [step_lengths] if isinstance(step_lengths, int) else list(step_lengths)
        p = ([1.0 / len(self._step_lens)] * len(self._step_lens) if step_probabilities is None
             else list(step_probabilities))                              # mcusher.py:244-247
        self._pcum = np.cumsum(p)
        self._pcum[-1] = 1.0

    def propose_step(self, occupancy, rnd: StepRandom, word0=0):
        u = u01(rnd.word(4))
        pick = len(self._step_lens) - 1
        for i, c in enumerate(self._pcum):
            if c > u:
                pick = i
                break
        step_length = self._step_lens[pick]
        occu = np.array(occupancy).copy()                                # mcusher.py:287
        steps = [self._mcusher.propose_step(occu, rnd, 8)]
        for f in steps[-1]:
            occu[f[0]] = f[1]
        for j in range(1, step_length):                                  # mcusher.py:293-301
            step = self._mcusher.propose_step(occu, rnd, 8 + 4 * j)
            if all(s not in (s for st in steps for s, _ in st) for s, _ in step):
                steps.append(step)
                for f in steps[-1]:
                    occu[f[0]] = f[1]
        return [flip for step in steps for flip in step]


def flip_weights_mask(flip_vectors, n, max_n):
    """utils/math.py:832-867."""
    fv = np.array(flip_vectors, dtype=int)
    directions = np.concatenate([(u, -u) for u in fv], axis=0)
    max_n = np.array(max_n, dtype=int)
    return ~(np.any(directions + n < 0, axis=-1) | np.any(directions + n > max_n, axis=-1))


def get_dim_ids_table(sublattices, active_only=False):
    """moca/occu_utils.py:27-58."""
    n_row = sum(len(s.sites) for s in sublattices)
    n_col = max(max(s.encoding) for s in sublattices) + 1
    table = np.zeros((n_row, n_col), dtype=int) - 1
    dim_id = 0
    for s in sublattices:
        for code in s.encoding:
            sites = (s.active_sites if active_only else s.sites).astype(int)
            table[sites, code] = dim_id
            dim_id += 1
    return table


class TableFlip(_Usher):
    """mcusher.py:397-711 with the flip table given (``counts`` format, one row per vector)."""

    def __init__(self, sublattices, flip_table, flip_weights=None, swap_weight=0.1):
        super().__init__(sublattices)
        self.flip_table = np.array(flip_table, dtype=int)
        self.swap_weight = swap_weight
        self.dim_ids, d = [], 0
        for s in self.sublattices:  # occu_utils.py:20-25
            self.dim_ids.append(list(range(d, d + len(s.species))))
            d += len(s.species)
        self.d = d
        self.max_n = [len(s.active_sites) for s in self.sublattices for _ in s.species]
        if flip_weights is None:
            self.flip_weights = np.ones(len(self.flip_table) * 2)
        elif len(flip_weights) == len(self.flip_table):
            self.flip_weights = np.repeat(np.asarray(flip_weights, float), 2)
        else:
            self.flip_weights = np.asarray(flip_weights, float)
        self._swapper = Swap(self.sublattices)
        self._dim_ids_table = get_dim_ids_table(self.sublattices, active_only=True)

    def _counts(self, occupancy):
        """occu_utils.py:96-128 (active sites only)."""
        occu = np.asarray(occupancy, dtype=int)
        dim = self._dim_ids_table[np.arange(len(occu)), occu]
        n = np.zeros(self.d, dtype=int)
        for x in dim[dim >= 0]:
            n[x] += 1
        return n

    def propose_step(self, occupancy, rnd: StepRandom, word0=0):
        """mcusher.py:553-639.

        Random words: block0.r0 swap decision, block0.r1 direction choice; a fallback /
        scheduled swap reads words 4,5,6 (block 1); site picks read words 4,5,6,... in
        pick order.  ``rng.choice(list, size=m, replace=False)`` is restated as m sequential
        bounded draws from the shrinking list (list order kept).
        """
        if u01(rnd.word(0)) < self.swap_weight:
            return self._swapper.propose_step(occupancy, rnd, word0=4)
        occu = np.asarray(occupancy, dtype=int)
        dim = self._dim_ids_table[np.arange(len(occu)), occu]
        species_list = [np.where(dim == i)[0].tolist() for i in range(self.d)]
        species_n = [len(s) for s in species_list]
        mask = flip_weights_mask(self.flip_table, species_n, self.max_n).astype(int)
        masked = self.flip_weights * mask
        if np.allclose(masked, 0):
            return self._swapper.propose_step(occupancy, rnd, word0=4)
        # utils/math.py:870-893: choose a section of the normalised partition
        p = masked / masked.sum()
        u = u01(rnd.word(1))
        cum = np.cumsum(p)
        idx = len(p) - 1
        for i, c in enumerate(cum):
            if c > u and p[i] > 0:
                idx = i
                break
        uvec = self.flip_table[idx // 2] * (-1 if idx % 2 == 1 else 1)
        step, w = [], 4
        for s, dim_ids in zip(self.sublattices, self.dim_ids):
            if not s.is_active:
                continue
            dim_ids = np.array(dim_ids, dtype=int)
            u_sl = uvec[dim_ids]
            site_ids = []
            for dd in dim_ids[u_sl < 0]:
                pool = list(species_list[dd])
                for _ in range(-uvec[dd]):
                    site_ids.append(pool.pop(mulhi32(rnd.word(w), len(pool))))
                    w += 1
            for dd, code in zip(dim_ids[u_sl > 0], s.encoding[u_sl > 0]):
                for _ in range(uvec[dd]):
                    site = site_ids.pop(mulhi32(rnd.word(w), len(site_ids)))
                    w += 1
                    step.append((int(site), int(code)))
            assert len(site_ids) == 0
        return step

    def _get_flip_id(self, occupancy, step):
        """mcusher.py:641-654 + occu_utils.py:131-169."""
        occu_now = np.array(occupancy, dtype=int)
        dn = np.zeros(self.d, dtype=int)
        for site, code in step:
            dn[self._dim_ids_table[site, occu_now[site]]] -= 1
            dn[self._dim_ids_table[site, code]] += 1
            occu_now[site] = code
        if np.all(dn == 0):
            return -1, 0
        for fid, v in enumerate(self.flip_table):
            if np.array_equal(v, dn):
                return fid, 0
            if np.array_equal(-v, dn):
                return fid, 1
        return None, None

    def compute_log_priori_factor(self, occupancy, step):
        """mcusher.py:656-711."""
        fid, direction = self._get_flip_id(occupancy, step)
        if fid is None:
            raise ValueError(f"Step {step} is not in flip table.")
        if fid < 0:
            return 0
        u = (-2 * direction + 1) * self.flip_table[fid]
        n_now = self._counts(occupancy)
        w_now = self.flip_weights * flip_weights_mask(self.flip_table, n_now, self.max_n)
        p_now = (1 - self.swap_weight) * w_now[fid * 2 + direction] / w_now.sum()
        n_next = n_now + u
        w_next = self.flip_weights * flip_weights_mask(self.flip_table, n_next, self.max_n)
        p_next = (1 - self.swap_weight) * w_next[fid * 2 + (1 - direction)] / w_next.sum()
        log_factor = math.log(p_next / p_now)
        for dim in np.nonzero(u)[0]:
            log_factor += math.lgamma(n_now[dim] + 1) - math.lgamma(n_next[dim] + 1)
        return log_factor


# ======================================================================================
# L4: kernels (smol/moca/kernel/base.py, metropolis.py, random.py, wanglandau.py)
# ======================================================================================
# ------------------------------------------------------------------------------------------------
# bias terms (smol/moca/kernel/bias.py)
# ------------------------------------------------------------------------------------------------
def _oxi_state(label) -> float:
    """get_oxi_state (smol/moca/composition/space.py) for species labels such as 'Mn3+', 'O2-', 'A'."""
    import re
    m = re.search(r"(\d*\.?\d*)([+-])$", str(label))
    if not m:
        return 0.0
    mag = float(m.group(1)) if m.group(1) else 1.0
    return mag if m.group(2) == "+" else -mag


class _Bias:
    def __init__(self, sublattices):
        self.sublattices = list(sublattices)
        self.active_sublattices = [s for s in self.sublattices if s.is_active]

    def _table(self, fill):
        num_cols = max(int(max(sl.encoding)) for sl in self.sublattices) + 1     # bias.py:226-229
        num_rows = sum(len(sl.sites) for sl in self.sublattices)
        return np.full((num_rows, num_cols), fill, dtype=np.float64)

    def compute_bias_change(self, occupancy, step):
        """MCBias.compute_bias_change, bias.py:79-93."""
        occu_next = np.array(occupancy).copy()
        for site, code in step:
            occu_next[site] = code
        return self.compute_bias(occu_next) - self.compute_bias(occupancy)


class FugacityBias(_Bias):
    """bias.py:96-233."""

    def __init__(self, sublattices, fugacity_fractions):
        super().__init__(sublattices)
        table = self._table(1.0)
        for fus, sl in zip(fugacity_fractions, self.active_sublattices):           # bias.py:230-233
            ordered = np.array([fus[sp] for sp in sl.species], dtype=np.float64)
            table[np.asarray(sl.sites)[:, None], np.asarray(sl.encoding)] = ordered[None, :]
        self._fu_table = table

    def compute_bias(self, occupancy):
        """bias.py:180-191."""
        return sum(math.log(self._fu_table[site, sp]) for site, sp in enumerate(occupancy))

    def compute_bias_change(self, occupancy, step):
        """bias.py:193-214: only the last flip of a site counts."""
        steps = {site: code for site, code in step}
        return sum(math.log(self._fu_table[site, code] / self._fu_table[site, occupancy[site]])
                   for site, code in steps.items())


class SquareChargeBias(_Bias):
    """bias.py:236-287."""

    def __init__(self, sublattices, penalty=0.5):
        super().__init__(sublattices)
        self.penalty = penalty
        table = self._table(0.0)
        for sl in self.sublattices:                                                # bias.py:267-270
            cs = np.array([_oxi_state(sp) for sp in sl.species], dtype=np.float64)
            table[np.asarray(sl.sites)[:, None], np.asarray(sl.encoding)] = cs[None, :]
        self._c_table = table

    def compute_bias(self, occupancy):
        """bias.py:276-287."""
        occupancy = np.asarray(occupancy)
        c = np.sum(self._c_table[np.arange(len(occupancy), dtype=int), occupancy])
        return -self.penalty * c ** 2


class SquareHyperplaneBias(_Bias):
    """bias.py:290-353 with get_dim_ids_table / occu_to_counts (occu_utils.py:27-57, 96-130)."""

    def __init__(self, sublattices, hyperplane_normals, hyperplane_intercepts, penalty=0.5):
        super().__init__(sublattices)
        self.penalty = penalty
        self._A = np.array(hyperplane_normals, dtype=int)
        self._b = np.array(hyperplane_intercepts, dtype=int)
        self._dim_ids_table = get_dim_ids_table(self.sublattices)
        self.d = sum(len(sl.species) for sl in self.sublattices)

    def compute_bias(self, occupancy):
        occu = np.array(occupancy, dtype=int)
        dim_ids = self._dim_ids_table[np.arange(len(occu), dtype=int), occu]     # occu_to_counts
        n = np.zeros(self.d, dtype=int)
        ids, cnt = np.unique(dim_ids[dim_ids >= 0], return_counts=True)
        n[ids] = cnt
        return -self.penalty * np.sum((self._A @ n - self._b) ** 2)


def _dot_seq(a, b) -> float:
    """Sequential dot product (the engine's order; np.dot's BLAS order is unspecified)."""
    p = 0.0
    for x, y in zip(a, b):
        p = p + float(x) * float(y)
    return p


class Metropolis:
    """kernel/metropolis.py:31-60 + kernel/base.py:145-166, 291-343, 368-436."""

    def __init__(self, ensemble, usher, temperature, seed=0, walker=0, kB_=kB, bias=None):
        self.ensemble = ensemble
        self.usher = usher
        self.bias = bias
        self.natural_params = ensemble.natural_parameters
        self.seed, self.walker = int(seed), int(walker)
        self.kB = kB_
        self.temperature = temperature
        self.step_index = 0

    @property
    def beta(self):
        return 1.0 / (self.kB * self.temperature)

    def compute_initial_trace(self, occupancy):
        feats = np.array(self.ensemble.compute_feature_vector(occupancy), dtype=np.float64)
        tr = SimpleNamespace(occupancy=np.array(occupancy), features=feats,
                             enthalpy=np.array([_dot_seq(self.natural_params, feats)]),
                             accepted=np.array([True]),
                             temperature=np.array([self.temperature], dtype=np.float64))
        if self.bias is not None:                                    # base.py:362-363
            tr.bias = np.array([self.bias.compute_bias(occupancy)], dtype=np.float64)
        return tr

    def set_aux_state(self, occupancy):
        return

    def single_step(self, occupancy):
        """base.py:145-166; occupancy modified in place on accept."""
        rnd = StepRandom(self.seed, self.walker, self.step_index)
        self.step_index += 1
        step = self.usher.propose_step(occupancy, rnd)
        dfeat = np.array(self.ensemble.compute_feature_vector_change(occupancy, step),
                         dtype=np.float64)
        dH = _dot_seq(self.natural_params, dfeat)
        log_factor = self.usher.compute_log_priori_factor(occupancy, step)
        exponent = -self.beta * dH + log_factor                      # metropolis.py:41
        dbias = 0.0
        if self.bias is not None:                                     # base.py:307-311, metropolis.py:43-44
            dbias = float(self.bias.compute_bias_change(occupancy, step))
            exponent += dbias
        # metropolis.py:46-48 (the uniform of this step is word 3 of block 0, drawn or not)
        accepted = True if exponent >= 0 else exponent > math.log(u01(rnd.word(3)))
        if accepted:
            for site, sp in step:                                     # base.py:339-340
                occupancy[site] = sp
        return SimpleNamespace(accepted=accepted, dfeatures=dfeat, denthalpy=dH, step=step,
                               exponent=exponent, dbias=dbias)


class UniformlyRandom(Metropolis):
    """kernel/random.py:14-36: accept iff log_priori >= 0 or > log(u) (beta = 0)."""

    def __init__(self, ensemble, usher, seed=0, walker=0):
        super().__init__(ensemble, usher, temperature=math.inf, seed=seed, walker=walker)

    @property
    def beta(self):
        return 0.0


class WangLandau:
    """kernel/wanglandau.py:20-305."""

    def __init__(self, ensemble, usher, min_enthalpy, max_enthalpy, bin_size, flatness=0.8,
                 mod_factor=1.0, check_period=1000, update_period=1, mod_update=2.0, seed=0,
                 walker=0):
        self.ensemble, self.usher = ensemble, usher
        self.natural_params = ensemble.natural_parameters
        self.seed, self.walker = int(seed), int(walker)
        self.flatness, self.check_period, self.update_period = flatness, check_period, update_period
        self._m = float(mod_factor)
        self._mod_update = mod_update
        self._window = (min_enthalpy, max_enthalpy, bin_size)
        self._levels = np.arange(min_enthalpy, max_enthalpy, bin_size)
        nb, nf = len(self._levels), len(self.natural_params)
        self._current_enthalpy = np.inf
        self._current_features = np.zeros(nf)
        self._entropy = np.zeros(nb)
        self._histogram = np.zeros(nb, dtype=np.int64)
        self._occurrences = np.zeros(nb, dtype=np.int64)
        self._mean_features = np.zeros((nb, nf))
        self._steps_counter = 0
        self.step_index = 0

    def _get_bin_id(self, e):
        """wanglandau.py:175-180."""
        if e == np.inf:
            return np.inf
        return int((e - self._window[0]) // self._window[2])

    def set_aux_state(self, occupancy):
        """wanglandau.py:290-300."""
        feats = np.array(self.ensemble.compute_feature_vector(occupancy), dtype=np.float64)
        self._current_features = feats
        self._current_enthalpy = _dot_seq(self.natural_params, feats)

    def compute_initial_trace(self, occupancy):
        feats = np.array(self.ensemble.compute_feature_vector(occupancy), dtype=np.float64)
        return SimpleNamespace(occupancy=np.array(occupancy), features=feats,
                               enthalpy=np.array([_dot_seq(self.natural_params, feats)]),
                               accepted=np.array([True]))

    def single_step(self, occupancy):
        rnd = StepRandom(self.seed, self.walker, self.step_index)
        self.step_index += 1
        step = self.usher.propose_step(occupancy, rnd)
        dfeat = np.array(self.ensemble.compute_feature_vector_change(occupancy, step),
                         dtype=np.float64)
        dH = _dot_seq(self.natural_params, dfeat)
        # _accept_step, wanglandau.py:186-202
        bin_id = self._get_bin_id(self._current_enthalpy)
        new_enthalpy = self._current_enthalpy + dH
        if new_enthalpy < self._window[0] or new_enthalpy >= self._window[1]:
            accepted = False
        else:
            new_bin = self._get_bin_id(new_enthalpy)
            exponent = self._entropy[bin_id] - self._entropy[new_bin] + \
                self.usher.compute_log_priori_factor(occupancy, step)
            accepted = True if exponent >= 0 else exponent > math.log(u01(rnd.word(3)))
        if accepted:                                                  # wanglandau.py:204-220
            for site, sp in step:
                occupancy[site] = sp
            self._current_features = self._current_features + dfeat
            self._current_enthalpy = self._current_enthalpy + dH
        # _do_post_step, wanglandau.py:222-266
        bin_id = self._get_bin_id(self._current_enthalpy)
        if 0 <= bin_id < len(self._levels):
            self._steps_counter += 1
            total = self._occurrences[bin_id]
            self._mean_features[bin_id, :] = (
                1 / (total + 1) * (self._current_features + total * self._mean_features[bin_id, :]))
            if self._steps_counter % self.update_period == 0:
                self._entropy[bin_id] += self._m
                self._histogram[bin_id] += 1
                self._occurrences[bin_id] += 1
        # trace (wanglandau.py:247-251): the arrays themselves -- a reset below shows in the sample -- but a COPY
        # of the modification factor taken before the flatness check
        self.trace_mod_factor = self._m
        if self._steps_counter % self.check_period == 0:
            histogram = self._histogram[self._entropy > 0]
            if len(histogram) >= 2 and (histogram > self.flatness * histogram.mean()).all():
                self._histogram[:] = 0
                self._m = self._mod_update(self._m) if callable(self._mod_update) else self._m / self._mod_update   # wanglandau.py:100-105
        return SimpleNamespace(accepted=accepted, dfeatures=dfeat, denthalpy=dH, step=step)


# ======================================================================================
# L5: sampler loop (smol/moca/sampler/sampler.py:164-210, 386-440)
# ======================================================================================
def run_sampler(kernels, initial_occupancies, nsteps, thin_by=1):
    """Return dict of sampled arrays ``[S, W, ...]`` (trace.py / sampler.py:123-128).

    ``accepted`` is the flag of the LAST step of each thinning interval
    (sampler.py:199-201); ``n_accepted`` (engine extension) counts accepts per interval.
    """
    occus = np.array(initial_occupancies, dtype=np.int32).copy()     # sampler.py:401-406
    if occus.ndim == 1:
        occus = occus[None, :]
    W = len(kernels)
    for k, o in zip(kernels, occus):
        k.set_aux_state(o)
    traces = [k.compute_initial_trace(o) for k, o in zip(kernels, occus)]
    feats = np.stack([t.features for t in traces])
    enth = np.stack([t.enthalpy for t in traces])
    S = nsteps // thin_by
    out = dict(occupancy=np.zeros((S, W, occus.shape[1]), dtype=np.int32),
               features=np.zeros((S, W, feats.shape[1])), enthalpy=np.zeros((S, W, 1)),
               accepted=np.zeros((S, W, 1), dtype=bool),
               n_accepted=np.zeros((S, W), dtype=np.int64))
    is_wl = all(hasattr(k, "_entropy") for k in kernels)
    if is_wl:
        nb, nf = kernels[0]._mean_features.shape
        out.update(entropy=np.zeros((S, W, nb)), histogram=np.zeros((S, W, nb), dtype=np.int64),
                   occurrences=np.zeros((S, W, nb), dtype=np.int64), mod_factor=np.zeros((S, W, 1)),
                   cumulative_mean_features=np.zeros((S, W, nb, nf)))
    has_bias = all(hasattr(t, "bias") for t in traces)
    bias = np.stack([t.bias for t in traces]) if has_bias else None
    if has_bias:
        out["bias"] = np.zeros((S, W, 1))
    acc = np.ones(W, dtype=bool)
    for s in range(S):
        nacc = np.zeros(W, dtype=np.int64)
        for _ in range(thin_by):
            for i, k in enumerate(kernels):
                st = k.single_step(occus[i])
                acc[i] = st.accepted
                if st.accepted:                                       # sampler.py:204-207
                    feats[i] += st.dfeatures
                    enth[i] += st.denthalpy
                    nacc[i] += 1
                    if has_bias:
                        bias[i] += st.dbias
        out["occupancy"][s] = occus
        out["features"][s] = feats
        out["enthalpy"][s] = enth
        out["accepted"][s, :, 0] = acc
        out["n_accepted"][s] = nacc
        if is_wl:
            for i, k in enumerate(kernels):
                out["entropy"][s, i], out["histogram"][s, i] = k._entropy, k._histogram
                out["occurrences"][s, i], out["mod_factor"][s, i, 0] = k._occurrences, k.trace_mod_factor
                out["cumulative_mean_features"][s, i] = k._mean_features
        if has_bias:
            out["bias"][s] = bias
    return out


class MulticellMetropolis:
    """kernel/base.py:439-722 (MulticellKernel) + kernel/metropolis.py:102-175 for ONE chain.

    ``mckernels``: oracle ``Metropolis`` kernels, one per supercell shape, all of one walker.  The shape and
    hop-period choices come from ``numpy.random.default_rng(seed).choice(..., p=...)`` in the reference's order
    (base.py:530-533, 664, 678-680).  Proposals and acceptance uniforms are counter based (Philox keyed by the
    sub-kernel's seed, counter = the chain's global step index); the reference draws a hop's uniform from the
    multicell generator (metropolis.py:46-48), data dependent.

    ``share_visited`` (default, what the reference DOES): ``MCKernel.single_step`` stores the array it was handed as
    ``trace.occupancy`` without copying (base.py:162), and the sampler hands every kernel the chain's one live row
    (sampler.py:436-440).  So once a shape has taken an ordinary step (or had a hop away from it rejected,
    base.py:671-674) its "own" occupancy IS the chain's live occupancy: a later hop into it starts from the current
    occupancy string re-read in that shape's supercell, and only shapes never visited keep the occupancy given at
    set_aux_state.  Pinned against the reference's class in tests/golden/ref_python_steps.npz.
    ``share_visited=False`` is the documented intent (base.py:665-666: every shape keeps an occupancy of its own)."""

    def __init__(self, mckernels, temperature, kernel_probabilities=None, kernel_hop_periods=5,
                 kernel_hop_probabilities=None, seed=None, kB_=kB, share_visited=True):
        self._kernels = list(mckernels)
        nk = len(self._kernels)
        self._kernel_p = np.array(kernel_probabilities if kernel_probabilities is not None else [1.0 / nk] * nk)
        self._hop_periods = np.array([kernel_hop_periods] if isinstance(kernel_hop_periods, int)
                                     else kernel_hop_periods, dtype=int)
        nh = len(self._hop_periods)
        self._hop_p = np.array(kernel_hop_probabilities if kernel_hop_probabilities is not None else [1.0 / nh] * nh)
        self._rng = np.random.default_rng(seed)
        self.temperature, self.kB = temperature, kB_
        self.natural_params = self._kernels[0].natural_params
        self._current_hop_period = self._rng.choice(self._hop_periods, p=self._hop_p)     # base.py:532
        self._kernel_hop_counter = 1
        self._current_kernel_index = 0
        self._features = np.zeros((nk, len(self.natural_params)))
        self._occupancies = None
        self.share_visited = bool(share_visited)
        self._live, self._alias = None, [False] * nk
        self.step_index = 0

    @property
    def beta(self):
        return 1.0 / (self.kB * self.temperature)

    def set_aux_state(self, occupancies):
        """base.py:694-716: one occupancy per shape; the chain's live row starts as a copy of shape 0's
        (sampler.py:411-418)."""
        self._occupancies = [np.array(o, dtype=np.int32) for o in occupancies]
        for i, (k, o) in enumerate(zip(self._kernels, self._occupancies)):
            self._features[i] = k.ensemble.compute_feature_vector(o)
        self._live = self._occupancies[self._current_kernel_index].copy()
        self._alias = [False] * len(self._kernels)

    def occupancy_of(self, k):
        """the array shape k would read as its own"""
        return self._live if (self.share_visited and self._alias[k]) else self._occupancies[k]

    @property
    def current_occupancy(self):
        return self._live if self.share_visited else self._occupancies[self._current_kernel_index]

    def single_step(self):
        """base.py:645-692.  Returns (accepted, current kernel index)."""
        t = self.step_index
        self.step_index += 1
        if self._kernel_hop_counter % self._current_hop_period == 0:
            new = int(self._rng.choice(len(self._kernels), p=self._kernel_p))              # base.py:664
            k = self._kernels[new]
            occ = self.occupancy_of(new)
            rnd = StepRandom(k.seed, k.walker, t)
            step = k.usher.propose_step(occ, rnd)
            trial = occ.copy()
            for site, sp in step:                                                            # base.py:612-614
                trial[site] = sp
            new_features = np.array(k.ensemble.compute_feature_vector(trial), dtype=np.float64)
            dfeat = new_features - self._features[self._current_kernel_index]                # base.py:616-619
            dH = _dot_seq(self.natural_params, dfeat)
            exponent = -self.beta * dH + k.usher.compute_log_priori_factor(occ, step)       # metropolis.py:40-41
            accepted = True if exponent >= 0 else exponent > math.log(u01(rnd.word(3)))
            if accepted:                                                                     # base.py:637-642, 665-669
                occ[:] = trial
                self._features[new] = new_features
                self._current_kernel_index = new
                if self.share_visited:
                    self._live[:] = occ                                                      # base.py:669
            elif self.share_visited:
                self._alias[self._current_kernel_index] = True                               # base.py:672-673
            self._current_hop_period = self._rng.choice(self._hop_periods, p=self._hop_p)   # base.py:678-680
            self._kernel_hop_counter = 1
        else:
            cur = self._current_kernel_index
            k = self._kernels[cur]
            k.step_index = t
            st = k.single_step(self.current_occupancy)                                       # base.py:684 / 162
            if self.share_visited:
                self._alias[cur] = True
            self._kernel_hop_counter += 1
            accepted = st.accepted
            if accepted:
                self._features[cur] += st.dfeatures                                          # base.py:686-689
        return accepted, self._current_kernel_index


def run_multicell(chains, initial_occupancies, nsteps, thin_by=1):
    """``chains``: one ``MulticellMetropolis`` per walker; ``initial_occupancies [W][K][N]``.  Sampled arrays
    ``[S, W, ...]`` of the CURRENT shape's state (occupancy, features, enthalpy) plus ``kernel_index``."""
    W = len(chains)
    for c, o in zip(chains, initial_occupancies):
        c.set_aux_state(o)
    S = nsteps // thin_by
    N, F = len(initial_occupancies[0][0]), len(chains[0].natural_params)
    out = dict(occupancy=np.zeros((S, W, N), dtype=np.int32), features=np.zeros((S, W, F)),
               enthalpy=np.zeros((S, W, 1)), accepted=np.zeros((S, W, 1), dtype=bool),
               n_accepted=np.zeros((S, W), dtype=np.int64), kernel_index=np.zeros((S, W, 1), dtype=np.int64))
    for s in range(S):
        for i, c in enumerate(chains):
            nacc, acc = 0, True
            for _ in range(thin_by):
                acc, _ = c.single_step()
                nacc += bool(acc)
            cur = c._current_kernel_index
            out["occupancy"][s, i] = c.current_occupancy
            out["features"][s, i] = c._features[cur]
            out["enthalpy"][s, i, 0] = _dot_seq(c.natural_params, c._features[cur])
            out["accepted"][s, i, 0] = acc
            out["n_accepted"][s, i] = nacc
            out["kernel_index"][s, i, 0] = cur
    return out
