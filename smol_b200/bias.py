"""Bias terms of the Metropolis kernel (mirror of ``smol/moca/kernel/bias.py``).

A bias adds ``bias(after) - bias(before)`` to the Metropolis exponent (``kernel/metropolis.py:43-44``).
On the device both supported terms are a per-site table and a running table sum per walker
(``LMC_BIAS_*`` in ``include/lmc.h``):

* ``FugacityBias`` (``bias.py:96-233``): ``sum_k log(fugacity_fraction[k][occ[k]])`` -> table of log fractions;
* ``SquareChargeBias`` (``bias.py:236-287``): ``-penalty * (sum_k oxidation_state[k][occ[k]])**2``;
* ``SquareHyperplaneBias`` (``bias.py:290-353``): ``-penalty * ||A n - b||**2``, one table row per hyperplane.

``compute_bias`` / ``compute_bias_change`` run the same device kernel as the sampler.
"""
from __future__ import annotations

import re

import numpy as np

from . import _capi as capi


def get_oxi_state(species) -> float:
    """Oxidation state of a species (``smol/moca/composition/space.py`` ``get_oxi_state``): the
    ``oxi_state`` attribute of a pymatgen ``Species``, else parsed from a label such as ``Mn3+``."""
    ox = getattr(species, "oxi_state", None)
    if ox is not None:
        return float(ox)
    m = re.search(r"(\d*\.?\d*)([+-])$", str(species))
    if not m:
        return 0.0
    mag = float(m.group(1)) if m.group(1) else 1.0
    return mag if m.group(2) == "+" else -mag


class MCBias:
    """Base class (``bias.py:26-93``): a per-site table, a device mode and an optional penalty."""

    mode = capi.LMC_BIAS_NONE
    penalty = 0.0
    intercepts = None      # per-row constants subtracted from the table sums (SquareHyperplaneBias: b)

    def __init__(self, sublattices, rng=None, **kwargs):
        self.sublattices = list(sublattices)
        self.active_sublattices = [s for s in self.sublattices if s.is_active]
        self._table = None

    @property
    def table(self) -> np.ndarray:
        """float64 ``[num_sites, max_code + 1]`` (``[..., rows]`` for several hyperplanes) table summed over the
        occupancy on the device."""
        return self._table

    @property
    def rows(self) -> int:
        return 1 if self._table.ndim == 2 else int(self._table.shape[2])

    def _intercept_array(self):
        ic = np.zeros(capi.LMC_MAX_BIAS_ROWS, dtype=np.float64)
        if self.intercepts is not None:
            ic[:self.rows] = np.asarray(self.intercepts, dtype=np.float64)
        return ic

    def _blank_table(self, fill):
        num_cols = max(int(max(sl.encoding)) for sl in self.sublattices) + 1
        num_rows = sum(len(sl.sites) for sl in self.sublattices)
        return np.full((num_rows, num_cols), fill, dtype=np.float64)

    def _device_values(self, occupancies):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("smol_b200 needs a CUDA device (there is no CPU fallback)")
        lib = capi.load()
        occ = np.atleast_2d(np.asarray(occupancies)).astype(np.int8)
        W, N = occ.shape
        stride = int(lib.lmc_row_stride(N))
        rows = np.zeros((W, stride), dtype=np.int8)
        rows[:, :N] = occ
        dev = torch.device("cuda", torch.cuda.current_device())
        occ_d = torch.from_numpy(rows).to(dev)
        tab_d = torch.from_numpy(np.ascontiguousarray(self.table)).to(dev)
        bias = torch.empty(W, dtype=torch.float64, device=dev)
        tsum = torch.empty((W, self.rows), dtype=torch.float64, device=dev)
        ic = self._intercept_array()
        capi.check(lib.lmc_bias_init(occ_d.data_ptr(), W, N, self.mode, self.table.shape[1], self.rows,
                                     float(self.penalty), ic.ctypes.data, tab_d.data_ptr(), bias.data_ptr(),
                                     tsum.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return bias.cpu().numpy()

    def compute_bias(self, occupancy):
        return float(self._device_values(occupancy)[0])

    def compute_bias_change(self, occupancy, step):
        """``bias.py:79-93``: bias of the occupancy with the step applied minus the current bias."""
        occu_next = np.array(occupancy).copy()
        for site, code in step:
            occu_next[site] = code
        vals = self._device_values(np.stack([np.asarray(occupancy), occu_next]))
        return float(vals[1] - vals[0])


class FugacityBias(MCBias):
    """``bias.py:96-233``."""

    mode = capi.LMC_BIAS_TABLE_SUM

    def __init__(self, sublattices, fugacity_fractions=None, **kwargs):
        super().__init__(sublattices, **kwargs)
        self._species = [set(sl.species) for sl in self.active_sublattices]
        if fugacity_fractions is None:
            # the reference takes the site-space compositions; label-only site spaces carry none: equal fractions
            fugacity_fractions = []
            for sl in self.active_sublattices:
                items = getattr(sl.site_space, "items", None)
                fugacity_fractions.append(dict(items()) if items is not None
                                          else {sp: 1.0 / len(sl.species) for sp in sl.species})
        self.fugacity_fractions = fugacity_fractions

    @property
    def fugacity_fractions(self):
        return self._fus

    @fugacity_fractions.setter
    def fugacity_fractions(self, value):
        value = [dict(sub) for sub in value]
        if not all(sum(fus.values()) == 1 for fus in value):                 # bias.py:167-168
            raise ValueError("Fugacity ratios must add to one.")
        for spec, vals in zip(self._species, value):                          # bias.py:169-176
            if spec != set(vals.keys()):
                raise ValueError("Fugacity fractions given are missing or not valid species.\n"
                                 f"Values must be given for each  of the following: {self._species}")
        self._fus = value
        fu = self._blank_table(1.0)                                           # bias.py:216-233
        for fus, sl in zip(value, self.active_sublattices):
            ordered = np.array([fus[sp] for sp in sl.species], dtype=np.float64)
            fu[np.asarray(sl.sites)[:, None], np.asarray(sl.encoding)] = ordered[None, :]
        self._fu_table = fu
        self._table = np.log(fu)


class SquareChargeBias(MCBias):
    """``bias.py:236-287``."""

    mode = capi.LMC_BIAS_SQUARE_SUM

    def __init__(self, sublattices, penalty=0.5, **kwargs):
        super().__init__(sublattices, **kwargs)
        if penalty <= 0:
            raise ValueError("Penalty factor should be > 0!")
        self.penalty = float(penalty)
        table = self._blank_table(0.0)
        for sl in self.sublattices:
            cs = np.array([get_oxi_state(sp) for sp in sl.species], dtype=np.float64)
            table[np.asarray(sl.sites)[:, None], np.asarray(sl.encoding)] = cs[None, :]
        self._c_table = table
        self._table = table


class SquareHyperplaneBias(MCBias):
    """``bias.py:290-353``: ``-penalty * ||A n - b||^2`` with ``n`` the species counts in "counts" format
    (one entry per (sublattice, species), ``occu_utils.py:27-57``)."""

    mode = capi.LMC_BIAS_SQUARE_SUM

    def __init__(self, sublattices, hyperplane_normals, hyperplane_intercepts, penalty=0.5, **kwargs):
        super().__init__(sublattices, **kwargs)
        if penalty <= 0:
            raise ValueError("Penalty factor should be > 0!")
        self.penalty = float(penalty)
        self._A = np.atleast_2d(np.array(hyperplane_normals, dtype=int))
        self._b = np.atleast_1d(np.array(hyperplane_intercepts, dtype=int))
        self.d = sum(len(sl.species) for sl in self.sublattices)
        if self._A.shape != (len(self._b), self.d):
            raise ValueError(f"hyperplane_normals must have shape [{len(self._b)}, {self.d}]")
        if len(self._b) > capi.LMC_MAX_BIAS_ROWS:
            raise ValueError(f"at most {capi.LMC_MAX_BIAS_ROWS} hyperplanes")
        # get_dim_ids_table (occu_utils.py:27-57), then column A[:, dim] per (site, code)
        num_cols = max(int(max(sl.encoding)) for sl in self.sublattices) + 1
        num_rows = sum(len(sl.sites) for sl in self.sublattices)
        dim_ids = np.full((num_rows, num_cols), -1, dtype=int)
        dim = 0
        for sl in self.sublattices:
            for code in sl.encoding:
                dim_ids[np.asarray(sl.sites, dtype=int), int(code)] = dim
                dim += 1
        self._dim_ids_table = dim_ids
        table = np.zeros((num_rows, num_cols, len(self._b)), dtype=np.float64)
        ok = dim_ids >= 0
        table[ok] = self._A.T[dim_ids[ok]]
        self._table = table
        self.intercepts = self._b.astype(np.float64)


_BIAS = {"fugacitybias": FugacityBias, "fugacity": FugacityBias,
         "squarechargebias": SquareChargeBias, "squarecharge": SquareChargeBias,
         "squarehyperplanebias": SquareHyperplaneBias, "squarehyperplane": SquareHyperplaneBias}


def mcbias_factory(bias_type, sublattices, *args, **kwargs):
    """``bias.py:355-372``."""
    key = str(bias_type).lower().replace("-", "").replace("_", "").replace(" ", "")
    if key not in _BIAS:
        raise ValueError(f"{bias_type} is not a supported MCBias (available: FugacityBias, SquareChargeBias, SquareHyperplaneBias)")
    return _BIAS[key](sublattices, *args, **kwargs)
