"""Step proposal descriptors (mirror of ``smol/moca/kernel/mcusher.py``).

The proposals themselves run inside the CUDA kernel; these classes only carry what the kernel needs -- the
usher type, its sublattices and their pick probabilities -- with the constructor arguments of the reference.
"""
from __future__ import annotations

import numpy as np

from . import _capi as capi


class MCUsher:
    """``mcusher.py:24-148``: active sublattices and the probabilities of picking each."""

    code = None

    def __init__(self, sublattices, sublattice_probabilities=None, rng=None):
        self.sublattices = list(sublattices)
        self.active_sublattices = [s for s in self.sublattices if s.is_active]
        n = len(self.active_sublattices)
        if sublattice_probabilities is None:
            probs = np.full(n, 1.0 / n)
        else:
            if len(sublattice_probabilities) != n:                               # mcusher.py:65-69
                raise AttributeError("sublattice_probabilities needs to be the same length as the number "
                                     f"of active sublattices. Got {len(sublattice_probabilities)} and {n}.")
            if sum(sublattice_probabilities) != 1:                               # mcusher.py:70-71
                raise ValueError("sublattice_probabilities must sum to one.")
            probs = np.asarray(sublattice_probabilities, dtype=np.float64)
        self.sublattice_probabilities = probs


class Flip(MCUsher):
    """``mcusher.py:151-170``."""
    code = capi.LMC_USHER_FLIP


class Swap(MCUsher):
    """``mcusher.py:173-200``."""
    code = capi.LMC_USHER_SWAP


class Composite(MCUsher):
    """``mcusher.py:307-394``: every step one of the sub-ushers, picked by weight, proposes."""

    code = capi.LMC_USHER_COMPOSITE

    def __init__(self, sublattices, mcushers=None, mcusher_weights=None, rng=None):
        super().__init__(sublattices)
        self._mcushers, self._weights, self._p = [], [], []
        mcushers = list(mcushers or [])
        if mcusher_weights is None:
            mcusher_weights = [1] * len(mcushers)
        if len(mcusher_weights) != len(mcushers):
            raise ValueError("mcusher_weights must have one entry per mcusher")
        for u, w in zip(mcushers, mcusher_weights):
            self.add_mcusher(u, w)

    @property
    def mcushers(self):
        return self._mcushers

    @property
    def weights(self):
        return self._weights

    def add_mcusher(self, mcusher, weight=1):
        """``mcusher.py:366-380``; names are built on the composite's own sublattices."""
        if isinstance(mcusher, str):
            mcusher = mcusher_factory(mcusher, self.sublattices)
        if not isinstance(mcusher, (Flip, Swap)):
            raise NotImplementedError("composite sub-ushers must be Flip or Swap on the GPU path")
        if len(self._mcushers) >= capi.LMC_MAX_COMPOSITE:
            raise ValueError(f"at most {capi.LMC_MAX_COMPOSITE} sub-ushers")
        self._mcushers.append(mcusher)
        self._weights.append(weight)
        total = sum(self._weights)                                               # mcusher.py:382-390
        self._p = [w / total for w in self._weights]

    def device_tables(self, model_sublattices):
        """(codes, cumulative pick probabilities, cumulative sublattice probabilities over the MODEL's active
        sublattices; a sublattice a sub-usher does not serve has zero width)"""
        active = [s for s in model_sublattices if s.is_active]
        cum = np.cumsum(self._p)
        cum[-1] = 1.0
        sl_cum = np.ones((len(self._mcushers), capi.LMC_MAX_SUBLATTICES))
        for i, u in enumerate(self._mcushers):
            width = np.zeros(len(active))
            for sl, p in zip(u.active_sublattices, u.sublattice_probabilities):
                hit = [k for k, t in enumerate(active)
                       if t is sl or (np.array_equal(t.sites, sl.sites) and tuple(t.species) == tuple(sl.species))]
                if len(hit) != 1:
                    raise ValueError("a sub-usher sublattice is not one of the ensemble's active sublattices")
                width[hit[0]] = p
            c = np.cumsum(width)
            last = int(np.nonzero(width)[0][-1])
            c[last:] = 1.0
            sl_cum[i, :len(active)] = c
        return [u.code for u in self._mcushers], cum, sl_cum


class MultiStep(MCUsher):
    """``mcusher.py:203-304``: a step chains ``step_length`` proposals of one sub-usher; a proposal that touches
    an already changed site is dropped.  (With ``step_probabilities`` the reference stores them in a misspelt
    attribute, ``mcusher.py:243``, and then fails in ``propose_step``; here they are used as documented.)"""

    code = capi.LMC_USHER_MULTISTEP

    def __init__(self, sublattices, mcusher, step_lengths, step_probabilities=None, rng=None):
        super().__init__(sublattices)
        lens = np.array([step_lengths] if isinstance(step_lengths, (int, np.integer)) else step_lengths, dtype=int)
        if step_probabilities is not None:
            if sum(step_probabilities) != 1.0:                                   # mcusher.py:237-238
                raise ValueError("The step_probabilities do not sum to 1.")
            if len(step_probabilities) != len(lens):                             # mcusher.py:239-242
                raise ValueError("The length of step_lengths and step_probabilities does not match.")
            probs = np.array(step_probabilities, dtype=np.float64)
        else:
            probs = np.full(len(lens), 1.0 / len(lens))
        if isinstance(mcusher, str):
            mcusher = mcusher_factory(mcusher, self.sublattices)
        if not isinstance(mcusher, (Flip, Swap)):
            raise NotImplementedError("the multi-step sub-usher must be Flip or Swap on the GPU path")
        per = 2 if isinstance(mcusher, Swap) else 1
        if len(lens) < 1 or len(lens) > capi.LMC_MAX_COMPOSITE or (lens < 1).any() or (lens * per > capi.LMC_MAX_FLIPS).any():
            raise ValueError("step lengths must change at most 4 sites per step (<= 4 flips or <= 2 swaps), "
                             "at most 4 different lengths")
        self._mcusher, self._step_lens, self._step_p = mcusher, lens, probs
        self.sublattice_probabilities = mcusher.sublattice_probabilities

    def device_tables(self):
        cum = np.cumsum(self._step_p)
        cum[-1] = 1.0
        return self._mcusher.code, [int(x) for x in self._step_lens], cum


_USHER_CLASSES = {"flip": Flip, "swap": Swap, "composite": Composite, "multistep": MultiStep}


def mcusher_factory(usher_type, sublattices, *args, **kwargs):
    """``mcusher.py`` ``mcusher_factory``."""
    key = str(usher_type).lower().replace("-", "").replace("_", "")
    if key not in _USHER_CLASSES:
        raise ValueError(f"{usher_type} is not a supported MCUsher here (Flip, Swap, Composite, MultiStep)")
    return _USHER_CLASSES[key](sublattices, *args, **kwargs)
