"""Shared model builders for the tests: the SAME tables feed the product (smol_b200) and the
oracle (oracle/lmc_oracle.py)."""
from __future__ import annotations

import functools

import numpy as np

from smol_b200 import lattice as L

S_FCC = {2: 6.0, 3: 3.5, 4: 3.0}          # point, pairs 1-4NN, NN triangle, NN tetrahedron


@functools.lru_cache(maxsize=None)
def fcc_subspace(basis="sinusoid"):
    return L.ClusterSubspace.from_cutoffs(L.fcc_prim(), S_FCC, basis=basis)


@functools.lru_cache(maxsize=None)
def rocksalt_subspace(cations=("Li+", "Mn3+", "Ti4+"), anions=("O2-",), cutoffs=None, basis="sinusoid"):
    cut = dict(cutoffs) if cutoffs else ({2: 6.0, 3: 3.5, 4: 3.0} if len(anions) == 1
                                         else {2: 4.3, 3: 3.5, 4: 3.0})
    return L.ClusterSubspace.from_cutoffs(L.rocksalt_prim(cations=cations, anions=anions), cut,
                                          basis=basis)


def fcc_coefs(subspace, seed=2024):
    """default_rng(2024).normal(0, 0.02) * multiplicity, coef[0] = 0."""
    rng = np.random.default_rng(seed)
    c = rng.normal(0, 0.02, subspace.num_corr_functions) * subspace.function_total_multiplicities
    c[0] = 0.0
    return c


def random_occupancies(subspace, scm, W, seed=0, balanced=False):
    rng = np.random.default_rng(seed)
    spaces = subspace.allowed_species(scm)
    N = len(spaces)
    occ = np.zeros((W, N), dtype=np.int32)
    ns = np.array([len(s) for s in spaces])
    if balanced:
        for w in range(W):
            for m in np.unique(ns):
                if m < 2:
                    continue
                sites = np.where(ns == m)[0]
                vals = np.arange(len(sites)) % m
                occ[w, sites] = rng.permutation(vals)
    else:
        occ = (rng.random((W, N)) * ns[None, :]).astype(np.int32)
    return occ


def oracle_sublattices(O, product_sublattices):
    return [O.Sublattice(s.species, s.sites, s.active_sites, s.encoding) for s in product_sublattices]
