"""ctypes driver of oracle/lmc_oracle.c (TEST INFRASTRUCTURE ONLY).

Packs the tables of the *Python oracle objects* (``oracle.lmc_oracle``) -- the reference
layout: full cluster-index arrays per orbit and, per site, the rows containing the site with
their ``cluster_ratio`` (processor/expansion.py:120-156) -- into flat arrays for the C
restatement.  Independent of the product's packer (smol_b200/model.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_c
from . import lmc_oracle as O

_P = C.c_void_p


class OModel(C.Structure):
    _fields_ = [("N", C.c_int32), ("F", C.c_int32), ("Fce", C.c_int32), ("n_orb", C.c_int32),
                ("size", C.c_int32), ("feature0", C.c_double), ("nat", _P),
                ("orb_fidx", _P), ("orb_K", _P), ("orb_T", _P), ("orb_I", _P), ("orb_stride", _P),
                ("orb_tab_off", _P), ("tab", _P), ("full_off", _P), ("full_idx_off", _P),
                ("full_rows", _P), ("site_ptr", _P), ("ent_orb", _P), ("ent_row_off", _P),
                ("ent_J", _P), ("ent_ratio", _P), ("local_rows", _P),
                ("E", C.c_int32), ("ewW", C.c_int32), ("ewF", C.c_int32), ("ewM", _P), ("ewInds", _P),
                ("muW", C.c_int32), ("muF", C.c_int32), ("mu", _P),
                ("n_sl", C.c_int32), ("sl_off", _P), ("sl_sites", _P), ("sl_ncodes", _P),
                ("sl_codes", _P), ("sl_cum", _P)]


class ORun(C.Structure):
    _fields_ = [("W", C.c_int32), ("walker_base", C.c_int32), ("usher", C.c_int32),
                ("kernel", C.c_int32), ("thin", C.c_int32), ("S", C.c_int64), ("step0", C.c_uint64),
                ("seeds", _P), ("beta", _P), ("occ", _P), ("features", _P), ("enthalpy", _P),
                ("tr_occ", _P), ("tr_feat", _P), ("tr_enth", _P), ("tr_acc", _P), ("tr_nacc", _P),
                ("wl_min", C.c_double), ("wl_max", C.c_double), ("wl_bin", C.c_double),
                ("wl_flat", C.c_double), ("wl_modupd", C.c_double),
                ("wl_nb", C.c_int32), ("wl_check", C.c_int32), ("wl_update", C.c_int32),
                ("wl_S", _P), ("wl_H", _P), ("wl_O", _P), ("wl_M", _P), ("wl_m", _P), ("wl_cnt", _P),
                ("nthreads", C.c_int32)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_c.build())
        _LIB.o_run.argtypes = [C.POINTER(OModel), C.POINTER(ORun)]
        _LIB.o_full_features.argtypes = [C.POINTER(OModel), _P, _P]
        _LIB.o_delta_features.argtypes = [C.POINTER(OModel), _P, _P, _P, C.c_int, _P]
    return _LIB


class COracle:
    """C restatement bound to one oracle ``Ensemble`` (expansion [+ Ewald] [+ mu])."""

    def __init__(self, ensemble: "O.Ensemble", sublattice_probabilities=None):
        self.keep = []
        m = OModel()
        proc = ensemble.processor
        procs = proc.processors if isinstance(proc, O.CompositeProcessor) else [proc]
        exp = next((p for p in procs if isinstance(p, (O.ClusterExpansionProcessor,
                                                       O.ClusterDecompositionProcessor))), None)
        ew = next((p for p in procs if isinstance(p, O.EwaldProcessor)), None)
        nat = np.ascontiguousarray(ensemble.natural_parameters, dtype=np.float64)
        m.N, m.F = ensemble.num_sites, len(nat)
        m.nat = self._p(nat)
        fcur = 0
        if exp is not None:
            inter = isinstance(exp, O.ClusterDecompositionProcessor)
            n_orb = len(exp._orbit_data)
            fidx, K, T, I, tab_off, tabs = [], [], [], [], [], []
            strides = np.zeros((n_orb, 8), dtype=np.int32)
            off = 0
            for n, (oid, bit_id, tensors, st) in enumerate(exp._orbit_data):
                tens = exp._flat[n][None, :] if inter else tensors
                fidx.append(oid if inter else bit_id)
                K.append(tens.shape[0]); T.append(tens.shape[1]); I.append(len(st))
                strides[n, :len(st)] = st
                tab_off.append(off); tabs.append(np.ravel(tens)); off += tens.size
            full_off, full_idx_off, rows = [0], [], []
            pos = 0
            for idx in exp._indices:
                full_off.append(full_off[-1] + len(idx)); full_idx_off.append(pos)
                rows.append(np.ravel(idx)); pos += idx.size
            site_ptr = [0]
            ent_orb, ent_row_off, ent_J, ent_ratio, local = [], [], [], [], []
            lpos = 0
            for site in range(m.N):
                for (n, odata, lrows, ratio) in exp._data_by_sites.get(site, []):
                    ent_orb.append(n); ent_row_off.append(lpos); ent_J.append(len(lrows))
                    ent_ratio.append(ratio); local.append(np.ravel(lrows)); lpos += lrows.size
                site_ptr.append(len(ent_orb))
            m.n_orb, m.size = n_orb, exp.size
            m.Fce = exp.num_features()
            m.feature0 = float(exp.offset) * exp.size if inter else float(exp.size)
            for name, arr, dt in (("orb_fidx", fidx, np.int32), ("orb_K", K, np.int32),
                                  ("orb_T", T, np.int32), ("orb_I", I, np.int32),
                                  ("orb_stride", strides, np.int32), ("orb_tab_off", tab_off, np.int64),
                                  ("tab", np.concatenate(tabs), np.float64),
                                  ("full_off", full_off, np.int64), ("full_idx_off", full_idx_off, np.int64),
                                  ("full_rows", np.concatenate(rows), np.int32),
                                  ("site_ptr", site_ptr, np.int64), ("ent_orb", ent_orb, np.int32),
                                  ("ent_row_off", ent_row_off, np.int64), ("ent_J", ent_J, np.int32),
                                  ("ent_ratio", ent_ratio, np.float64),
                                  ("local_rows", np.concatenate(local), np.int32)):
                setattr(m, name, self._p(np.ascontiguousarray(arr, dtype=dt)))
            fcur = m.Fce
        else:
            m.n_orb, m.size, m.Fce = 0, 1, 0
            m.site_ptr = self._p(np.zeros(m.N + 1, dtype=np.int64))
        if ew is not None:
            m.E, m.ewW, m.ewF = ew.ewald_matrix.shape[0], ew._ewald_inds.shape[1], fcur
            m.ewM, m.ewInds = self._p(ew.ewald_matrix), self._p(ew._ewald_inds)
            fcur += 1
        if ensemble.mu_table is not None:
            mu = np.ascontiguousarray(ensemble.mu_table, dtype=np.float64)
            m.muW, m.muF, m.mu = mu.shape[1], fcur, self._p(mu)
            fcur += 1
        assert fcur == m.F
        active = ensemble.active_sublattices
        off = np.zeros(len(active) + 1, dtype=np.int32)
        codes = np.zeros((len(active), 8), dtype=np.int32)
        ncodes = np.zeros(len(active), dtype=np.int32)
        for i, s in enumerate(active):
            off[i + 1] = off[i] + len(s.active_sites)
            ncodes[i] = len(s.encoding); codes[i, :len(s.encoding)] = s.encoding
        probs = (np.full(len(active), 1.0 / len(active)) if sublattice_probabilities is None
                 else np.asarray(sublattice_probabilities, dtype=np.float64))
        cum = np.cumsum(probs); cum[-1] = 1.0
        m.n_sl = len(active)
        m.sl_off, m.sl_ncodes, m.sl_codes = self._p(off), self._p(ncodes), self._p(codes)
        m.sl_sites = self._p(np.ascontiguousarray(np.concatenate([s.active_sites for s in active]),
                                                  dtype=np.int32))
        m.sl_cum = self._p(np.ascontiguousarray(cum))
        self.model = m
        self.N, self.F = m.N, m.F

    def _p(self, arr):
        arr = np.ascontiguousarray(arr)
        self.keep.append(arr)
        return arr.ctypes.data_as(_P)

    def full_features(self, occ):
        occ = np.ascontiguousarray(occ, dtype=np.int32)
        out = np.zeros(self.F)
        lib().o_full_features(C.byref(self.model), occ.ctypes.data_as(_P), out.ctypes.data_as(_P))
        return out

    def delta_features(self, occ, flips):
        occ = np.ascontiguousarray(occ, dtype=np.int32).copy()
        sites = np.ascontiguousarray([f[0] for f in flips], dtype=np.int32)
        codes = np.ascontiguousarray([f[1] for f in flips], dtype=np.int32)
        out = np.zeros(self.F)
        lib().o_delta_features(C.byref(self.model), occ.ctypes.data_as(_P), sites.ctypes.data_as(_P),
                               codes.ctypes.data_as(_P), len(flips), out.ctypes.data_as(_P))
        return out

    def run(self, occ0, nsteps, thin_by, seeds, usher="swap", temperature=None, wl=None,
            walker_base=0, step0=0, nthreads=1, record=True, kB=O.kB):
        """Same semantics as ``lmc_oracle.run_sampler``; returns (trace dict, final state dict)."""
        occ = np.ascontiguousarray(occ0, dtype=np.int32).copy()
        W, N, F = occ.shape[0], self.N, self.F
        S = nsteps // thin_by
        feats = np.stack([self.full_features(o) for o in occ])
        enth = np.array([O._dot_seq(np.ctypeslib.as_array(
            (C.c_double * F).from_address(self.model.nat), shape=(F,)), f) for f in feats])
        r = ORun()
        r.W, r.walker_base, r.thin, r.S, r.step0 = W, walker_base, thin_by, S, step0
        r.usher = {"flip": 0, "swap": 1}[usher]
        r.kernel = 0 if wl is None else 1
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        beta = np.zeros(W)
        if wl is None:
            t = np.broadcast_to(np.asarray(temperature, dtype=np.float64), (W,))
            beta = np.ascontiguousarray(1.0 / (kB * t))
        keep = [seeds, beta, occ, feats, enth]
        r.seeds, r.beta = seeds.ctypes.data_as(_P), beta.ctypes.data_as(_P)
        r.occ, r.features, r.enthalpy = (occ.ctypes.data_as(_P), feats.ctypes.data_as(_P),
                                         enth.ctypes.data_as(_P))
        out = {}
        if record:
            out = dict(occupancy=np.zeros((S, W, N), dtype=np.int32), features=np.zeros((S, W, F)),
                       enthalpy=np.zeros((S, W)), accepted=np.zeros((S, W), dtype=np.uint8),
                       n_accepted=np.zeros((S, W), dtype=np.int32))
            r.tr_occ, r.tr_feat = out["occupancy"].ctypes.data_as(_P), out["features"].ctypes.data_as(_P)
            r.tr_enth, r.tr_acc = out["enthalpy"].ctypes.data_as(_P), out["accepted"].ctypes.data_as(_P)
            r.tr_nacc = out["n_accepted"].ctypes.data_as(_P)
        else:
            nacc = np.zeros((S, W), dtype=np.int32)
            out = dict(n_accepted=nacc)
            r.tr_nacc = nacc.ctypes.data_as(_P)
        state = {}
        if wl is not None:
            levels = np.arange(wl["min"], wl["max"], wl["bin"])
            nb = len(levels)
            state = dict(entropy=np.zeros((W, nb)), histogram=np.zeros((W, nb), dtype=np.int64),
                         occurrences=np.zeros((W, nb), dtype=np.int64),
                         mean_features=np.zeros((W, nb, F)),
                         mod_factor=np.full(W, float(wl.get("mod_factor", 1.0))),
                         steps_counter=np.zeros(W, dtype=np.int64))
            r.wl_min, r.wl_max, r.wl_bin = wl["min"], wl["max"], wl["bin"]
            r.wl_flat, r.wl_modupd = wl.get("flatness", 0.8), wl.get("mod_update", 2.0)
            r.wl_nb, r.wl_check, r.wl_update = nb, wl.get("check", 1000), wl.get("update", 1)
            r.wl_S, r.wl_H = state["entropy"].ctypes.data_as(_P), state["histogram"].ctypes.data_as(_P)
            r.wl_O, r.wl_M = state["occurrences"].ctypes.data_as(_P), state["mean_features"].ctypes.data_as(_P)
            r.wl_m, r.wl_cnt = state["mod_factor"].ctypes.data_as(_P), state["steps_counter"].ctypes.data_as(_P)
        r.nthreads = nthreads
        lib().o_run(C.byref(self.model), C.byref(r))
        if record:
            out["enthalpy"] = out["enthalpy"][:, :, None]
            out["accepted"] = out["accepted"].astype(bool)[:, :, None]
        state.update(occupancy=occ, features=feats, enthalpy=enth)
        del keep
        return out, state
