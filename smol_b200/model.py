"""Table packer: processor / ensemble tables -> the flat arrays of ``LmcModelDesc``.

This is the device-facing counterpart of the per-site evaluator construction in
``smol/moca/processor/expansion.py:120-156`` (and ``:344-392`` for the decomposition
processor): for every site the rows of each orbit's cluster-index array that contain the
site become *records*; ``cluster_ratio = total_rows / rows_with_site`` is folded into the
orbit weight ``size / total_rows`` (``p / ratio / J`` of ``evaluator.pyx:262`` equals
``p / total_rows``).

A record stores the (up to three) OTHER sites of the row plus a class id.  The class carries
the flat-tensor strides of those sites and the *self stride* -- the sum of the strides of all
positions of the row that hold the flipped site (more than one position on small, aliased
supercells, ``clusterspace.py:1353-1359``), so that
``ind_f - ind_i = (new_code - old_code) * self_stride``.
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import _capi as capi


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class ExpansionTables:
    """Host tables of the cluster part (shared by CE and decomposition processors)."""

    def __init__(self, cluster_subspace, supercell_matrix, mode, coefs=None,
                 interaction_tensors=None, num_sites=None):
        scm = np.asarray(supercell_matrix)
        self.size = int(round(abs(np.linalg.det(scm))))
        orbits = list(cluster_subspace.orbits)
        indices = [np.ascontiguousarray(a, dtype=np.int64)
                   for a in cluster_subspace.get_orbit_indices(scm).arrays]
        self.num_sites = int(num_sites) if num_sites is not None else (
            1 + max(int(a.max()) for a in indices) if indices else 0)
        self.mode = mode
        n_orb = len(orbits)
        tab_off, tab_len, nfunc, fidx, csize = [], [], [], [], []
        strides = np.zeros((n_orb, capi.LMC_MAX_CLUSTER_SITES), dtype=np.int32)
        weight = np.zeros(n_orb)
        ftab = []
        off = 0
        for n, (orb, idx) in enumerate(zip(orbits, indices)):
            st = np.asarray(orb.flat_tensor_indices, dtype=np.int64)
            if len(st) > capi.LMC_MAX_CLUSTER_SITES:
                raise ValueError("clusters with more than 4 sites are not supported by liblmc")
            if mode == "correlation":
                tens = np.ascontiguousarray(orb.flat_correlation_tensors, dtype=np.float64)
                f0 = orb.bit_id
            else:
                tens = np.ravel(np.asarray(interaction_tensors[n + 1], dtype=np.float64),
                                order="C")[None, :]
                f0 = orb.id
            tab_off.append(off)
            tab_len.append(tens.shape[1])
            nfunc.append(tens.shape[0])
            fidx.append(f0)
            csize.append(len(st))
            strides[n, :len(st)] = st
            weight[n] = self.size / len(idx)
            ftab.append(tens.ravel())
            off += tens.size
        self.orb_tab_off, self.orb_tab_len = _i32(tab_off), _i32(tab_len)
        self.orb_nfunc, self.orb_fidx, self.orb_csize = _i32(nfunc), _i32(fidx), _i32(csize)
        self.orb_stride, self.orb_weight = _i32(strides), _f64(weight)
        self.ftab = _f64(np.concatenate(ftab)) if ftab else np.zeros(0)
        self.num_orbits = n_orb
        if mode == "correlation":
            self.num_features = cluster_subspace.num_corr_functions
            self.feature0 = float(self.size)                      # corr[0] = 1 (evaluator.pyx:143)
        else:
            self.num_features = cluster_subspace.num_orbits
            self.feature0 = float(interaction_tensors[0]) * self.size  # offset (evaluator.pyx:191)
        # ---- full rows, padded to 4 columns with site 0 / stride 0 ------------------------
        row_off = [0]
        rows = []
        for idx in indices:
            pad = np.zeros((len(idx), capi.LMC_MAX_CLUSTER_SITES), dtype=np.uint16)
            pad[:, :idx.shape[1]] = idx
            rows.append(pad)
            row_off.append(row_off[-1] + len(idx))
        self.orb_row_off = np.ascontiguousarray(row_off, dtype=np.int64)
        self.full_rows = (np.ascontiguousarray(np.concatenate(rows)) if rows
                          else np.zeros((0, 4), dtype=np.uint16))
        self._pack_records(indices, strides)

    def _pack_records(self, indices, strides):
        N = self.num_sites
        classes = {}
        rec_site, rec_orb, rec_row, rec_oth, rec_cls = [], [], [], [], []
        for n, idx in enumerate(indices):
            J, I = idx.shape
            st = strides[n, :I].astype(np.int64)
            rows_all = np.arange(J)
            for r in range(1, I + 1):
                for mask in itertools.combinations(range(I), r):
                    p0 = mask[0]
                    sel = np.ones(J, dtype=bool)
                    for p in range(I):
                        same = idx[:, p] == idx[:, p0]
                        sel &= same if p in mask else ~same
                    if not sel.any():
                        continue
                    others = [p for p in range(I) if p not in mask]
                    key = (n, tuple(int(st[p]) for p in others), int(st[list(mask)].sum()))
                    cid = classes.setdefault(key, len(classes))
                    k = int(sel.sum())
                    oth = np.zeros((k, 3), dtype=np.int64)
                    for q, p in enumerate(others):
                        oth[:, q] = idx[sel, p]
                    rec_site.append(idx[sel, p0])
                    rec_orb.append(np.full(k, n))
                    rec_row.append(rows_all[sel])
                    rec_oth.append(oth)
                    rec_cls.append(np.full(k, cid))
        if rec_site:
            site = np.concatenate(rec_site)
            orb = np.concatenate(rec_orb)
            row = np.concatenate(rec_row)
            oth = np.concatenate(rec_oth)
            cls = np.concatenate(rec_cls)
            order = np.lexsort((row, orb, site))
            site, orb, oth, cls = site[order], orb[order], oth[order], cls[order]
        else:
            site = orb = cls = np.zeros(0, dtype=np.int64)
            oth = np.zeros((0, 3), dtype=np.int64)
        rec = np.zeros((len(site), 4), dtype=np.uint16)
        rec[:, :3] = oth
        rec[:, 3] = cls
        self.site_rec = np.ascontiguousarray(rec)
        self.site_rec_off = np.zeros(N + 1, dtype=np.int64)
        np.add.at(self.site_rec_off, site + 1, 1)
        self.site_rec_off = np.cumsum(self.site_rec_off)
        # orbit segments per site
        if len(site):
            change = np.ones(len(site), dtype=bool)
            change[1:] = (site[1:] != site[:-1]) | (orb[1:] != orb[:-1])
            starts = np.where(change)[0]
            counts = np.diff(np.append(starts, len(site)))
            seg_site = site[starts]
            seg = np.stack([starts - self.site_rec_off[seg_site], counts, orb[starts]], axis=1)
        else:
            seg_site = np.zeros(0, dtype=np.int64)
            seg = np.zeros((0, 3), dtype=np.int64)
        self.site_seg = _i32(seg)
        self.site_seg_off = np.zeros(N + 1, dtype=np.int64)
        np.add.at(self.site_seg_off, seg_site + 1, 1)
        self.site_seg_off = np.cumsum(self.site_seg_off)
        ncls = len(classes)
        self.cls_orbit = np.zeros(ncls, dtype=np.int32)
        self.cls_stride = np.zeros((ncls, 4), dtype=np.int32)
        for (n, oth_st, self_st), cid in classes.items():
            self.cls_orbit[cid] = n
            self.cls_stride[cid, :len(oth_st)] = oth_st
            self.cls_stride[cid, 3] = self_st
        self.num_classes = ncls


class PackedModel:
    """All host arrays of one ``LmcModelDesc`` (kept alive while the descriptor is in use)."""

    def __init__(self, num_sites, natural_parameters, sublattices, expansion=None,
                 ewald=None, mu_table=None, table_flip=None, sublattice_probabilities=None):
        self.keep = []
        d = capi.LmcModelDesc()
        d.abi_version = capi.LMC_ABI_VERSION
        nat = _f64(natural_parameters)
        self.natural_parameters = nat
        F = len(nat)
        d.num_sites = int(num_sites)
        d.num_features = F
        d.natural_parameters = self._ptr(nat)
        fcur = 0
        if expansion is not None:
            e = expansion
            d.num_ce_features = e.num_features
            d.supercell_size = e.size
            d.feature0 = e.feature0
            d.num_orbits = e.num_orbits
            for name in ("orb_tab_off", "orb_tab_len", "orb_nfunc", "orb_fidx", "orb_csize",
                         "orb_stride", "orb_weight", "ftab", "orb_row_off", "full_rows",
                         "cls_orbit", "cls_stride", "site_rec_off", "site_rec", "site_seg_off",
                         "site_seg"):
                setattr(d, name, self._ptr(getattr(e, name)))
            d.ftab_len = len(e.ftab)
            d.num_classes = e.num_classes
            fcur = e.num_features
        else:
            z64 = np.zeros(int(num_sites) + 1, dtype=np.int64)
            d.num_ce_features = 0
            d.supercell_size = 1
            d.num_orbits = 0
            d.orb_row_off = self._ptr(np.zeros(1, dtype=np.int64))
            d.site_rec_off = self._ptr(z64)
            d.site_seg_off = self._ptr(z64)
            d.ftab_len = 0
        if ewald is not None:
            matrix, inds = ewald
            matrix, inds = _f64(matrix), _i32(inds)
            d.ewald_size = matrix.shape[0]
            d.ewald_width = inds.shape[1]
            d.ewald_matrix = self._ptr(matrix)
            d.ewald_inds = self._ptr(inds)
            d.ewald_feature = fcur
            fcur += 1
        if mu_table is not None:
            mu = _f64(mu_table)
            d.mu_width = mu.shape[1]
            d.mu_table = self._ptr(mu)
            d.mu_feature = fcur
            fcur += 1
        if fcur != F:
            raise ValueError(f"natural_parameters has {F} entries, tables define {fcur} features")
        # active sublattices
        active = [s for s in sublattices if len(s.active_sites) > 0]
        if not 1 <= len(active) <= capi.LMC_MAX_SUBLATTICES:
            raise ValueError("between 1 and 8 active sublattices are supported")
        off = np.zeros(len(active) + 1, dtype=np.int32)
        codes = np.zeros((len(active), capi.LMC_MAX_CODES), dtype=np.int32)
        ncodes = np.zeros(len(active), dtype=np.int32)
        for i, s in enumerate(active):
            off[i + 1] = off[i] + len(s.active_sites)
            enc = np.asarray(s.encoding, dtype=np.int32)
            if len(enc) > capi.LMC_MAX_CODES or enc.max() >= capi.LMC_MAX_CODES:
                raise ValueError("at most 8 species codes (< 8) per sublattice")
            ncodes[i] = len(enc)
            codes[i, :len(enc)] = enc
        probs = (np.full(len(active), 1.0 / len(active)) if sublattice_probabilities is None
                 else _f64(sublattice_probabilities))          # mcusher.py:61-75
        if len(probs) != len(active):
            raise AttributeError("Sublattice probabilities needs to be the same length as "
                                 "sublattices.")
        if abs(float(np.sum(probs)) - 1.0) > 1e-12:
            raise ValueError("Sublattice probabilities must sum to one.")
        d.num_sublattices = len(active)
        d.sl_site_off = self._ptr(off)
        d.sl_sites = self._ptr(_i32(np.concatenate([s.active_sites for s in active])))
        d.sl_ncodes = self._ptr(ncodes)
        d.sl_codes = self._ptr(codes)
        d.sl_prob = self._ptr(_f64(probs))
        self.active_sublattices = active
        if table_flip is not None:
            tf = table_flip
            d.tf_num_dims = tf["num_dims"]
            d.tf_num_flips = len(tf["table"])
            d.tf_table = self._ptr(_i32(tf["table"]))
            d.tf_weights = self._ptr(_f64(tf["weights"]))
            d.tf_max_n = self._ptr(_i32(tf["max_n"]))
            d.tf_dim_sl = self._ptr(_i32(tf["dim_sl"]))
            d.tf_dim_code = self._ptr(_i32(tf["dim_code"]))
            d.tf_swap_weight = float(tf["swap_weight"])
        self.desc = d

    def _ptr(self, arr):
        arr = np.ascontiguousarray(arr)
        self.keep.append(arr)
        return arr.ctypes.data_as(C.c_void_p)
