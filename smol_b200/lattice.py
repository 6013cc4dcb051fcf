"""Cluster-subspace tables built by direct lattice geometry (no pymatgen).

The Monte-Carlo hot path consumes *tables*: per-orbit correlation tensors, tensor
strides and the supercell cluster-index arrays.  In smol those come from
``ClusterSubspace`` (``smol/cofe/space/clusterspace.py``) which needs pymatgen.  This
module produces the same tables from first principles so that the engine, the tests
and the benchmark can build models on a machine that only has numpy:

* orbit enumeration under the full space group of the primitive cell
  (semantics of ``clusterspace.py:1368-1560`` -- clusters of ``size`` sites whose
  diameter is within ``cutoffs[size]``, grouped into symmetry orbits, sorted by size,
  then diameter, then decreasing multiplicity);
* site bases ``sinusoid`` / ``indicator`` with the QR orthonormalisation of
  ``smol/cofe/space/basis.py:207-258, 573-586``;
* bit combos, correlation tensors and flat tensor strides with the semantics of
  ``smol/cofe/space/orbit.py:137-155, 217-275``;
* supercell cluster index arrays with the layout of
  ``clusterspace.py:1329-1366`` -- rows ordered (equivalent cluster, translation),
  duplicate rows kept;
* cluster-interaction tensors (``smol/cofe/expansion.py:171-201``);
* an Ewald pair matrix for the "all species overlaid" structure with the index
  layout of ``smol/cofe/extern/ewald.py:64-100`` (the arithmetic itself is
  pymatgen's in the reference and therefore an *input* to the engine).

The objects are duck-type compatible with the attributes the reference processors
read from ``ClusterSubspace`` / ``Orbit`` (``orbits``, ``num_orbits``,
``num_corr_functions``, ``orbit_multiplicities``, ``get_orbit_indices`` ...).
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from functools import reduce

import numpy as np

EPS_MULT = 10  # smol/cofe/space/basis.py (threshold multiplier used in orthonormalize)
SITE_TOL = 1e-6


# --------------------------------------------------------------------------------------
# site bases
# --------------------------------------------------------------------------------------
def _sinusoid(n: int, m: int, s: int) -> float:
    """basis.py:573-586 -- n-th (1-based) sinusoid site function at species index s."""
    a = -(-n // 2)
    if n % 2 == 0:
        return -math.sin(2 * math.pi * a * s / m)
    return -math.cos(2 * math.pi * a * s / m)


def site_function_array(n_species: int, basis: str = "sinusoid", orthonormal: bool = True,
                        measure=None) -> np.ndarray:
    """Non-constant site functions as rows, ``[n_species-1, n_species]``.

    Follows ``StandardBasis._construct_function_array`` and ``orthonormalize``
    (basis.py:207-258).
    """
    m = n_species
    if basis == "sinusoid":
        funcs = [[_sinusoid(n, m, s) for s in range(m)] for n in range(1, m)]
    elif basis == "indicator":
        funcs = [[float(s == n) for s in range(m)] for n in range(m - 1)]
    else:
        raise ValueError(f"unknown site basis {basis}")
    f_array = np.vstack((np.ones(m), np.array(funcs, dtype=np.float64).reshape(m - 1, m)))
    if orthonormal:
        measure = np.full(m, 1.0 / m) if measure is None else np.asarray(measure, float)
        q_mat, r_mat = np.linalg.qr((np.sqrt(measure) * f_array).T, mode="complete")
        q_mat[abs(q_mat) < EPS_MULT * np.finfo(np.float64).eps] = 0.0
        f_array = (q_mat.T / q_mat[:, 0]).astype(np.float64)
    return np.ascontiguousarray(f_array[1:])


# --------------------------------------------------------------------------------------
# primitive cell + symmetry
# --------------------------------------------------------------------------------------
@dataclass
class PrimCell:
    """Primitive cell: lattice rows (A), fractional coords and per-site species spaces.

    ``site_spaces[b]`` is the *sorted* tuple of species labels allowed on basis site b
    (position in the tuple == occupancy code, ``cofe/space/domain.py:157-161``).
    ``charges`` maps a species label to its oxidation state (used only for Ewald).
    """

    lattice: np.ndarray
    frac_coords: np.ndarray
    site_spaces: list
    charges: dict = field(default_factory=dict)

    def __post_init__(self):
        self.lattice = np.asarray(self.lattice, dtype=np.float64).reshape(3, 3)
        self.frac_coords = np.asarray(self.frac_coords, dtype=np.float64).reshape(-1, 3)
        self.site_spaces = [tuple(s) for s in self.site_spaces]

    @property
    def num_sites(self):
        return len(self.frac_coords)

    def space_group(self):
        """All (R, t) in fractional coordinates mapping the decorated cell onto itself."""
        L = self.lattice
        G = L @ L.T
        vals = np.array(list(itertools.product((-1, 0, 1), repeat=9)), dtype=np.int64)
        Rs = vals.reshape(-1, 3, 3)
        # f' = R f (column vectors): cart = f^T L -> metric preserved iff R^T G R = G
        ok = np.all(np.abs(np.einsum("nji,jk,nkl->nil", Rs, G, Rs) - G) < 1e-8, axis=(1, 2))
        rots = Rs[ok]
        ops = []
        f = self.frac_coords
        for R in rots:
            Rf = f @ R.T
            for j in range(len(f)):
                if self.site_spaces[j] != self.site_spaces[0]:
                    continue
                t = f[j] - Rf[0]
                t = t - np.round(t)
                img = Rf + t
                good = True
                for b in range(len(f)):
                    d = img[b][None, :] - f
                    d -= np.round(d)
                    hit = np.where(np.all(np.abs(d) < SITE_TOL, axis=1))[0]
                    if len(hit) != 1 or self.site_spaces[hit[0]] != self.site_spaces[b]:
                        good = False
                        break
                if good and not any(np.array_equal(R, R2) and np.allclose(t, t2, atol=SITE_TOL)
                                    for R2, t2 in ops):
                    ops.append((R.copy(), t.copy()))
        return ops


def fcc_prim(a: float = 4.09, species=("A", "B"), charges=None) -> PrimCell:
    """FCC primitive cell, rows (0,a/2,a/2),(a/2,0,a/2),(a/2,a/2,0) as tests/data/AuPd_prim.json."""
    lat = 0.5 * a * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float)
    return PrimCell(lat, [[0, 0, 0]], [tuple(species)], charges or {})


def rocksalt_prim(a: float = 4.2, cations=("Li+", "Mn3+", "Ti4+"), anions=("O2-",),
                  charges=None) -> PrimCell:
    """Rocksalt primitive cell: cation (0,0,0), anion (1/2,1/2,1/2)."""
    lat = 0.5 * a * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float)
    if charges is None:
        charges = {"Li+": 1, "Mn3+": 3, "Ti4+": 4, "O2-": -2, "F-": -1, "Mn2+": 2,
                   "Zr4+": 4}
    return PrimCell(lat, [[0, 0, 0], [0.5, 0.5, 0.5]], [tuple(cations), tuple(anions)], charges)


# --------------------------------------------------------------------------------------
# orbits
# --------------------------------------------------------------------------------------
class Orbit:
    """Symmetry orbit of clusters with the table attributes the evaluators consume.

    ``clusters``: list of ordered site lists ``[(b, n1, n2, n3), ...]`` -- one per
    symmetry-equivalent cluster per primitive cell, site order carried through the
    symmetry operation so that bit orderings stay consistent (orbit.py:157-215).
    """

    def __init__(self, base_cluster, clusters, permutations, diameter, function_arrays):
        self.base_cluster = base_cluster
        self.clusters = clusters
        self.cluster_permutations = permutations
        self.diameter = diameter
        self.basis_arrays = function_arrays  # tuple of [n_i-1, n_i] per site
        self.id = None
        self.bit_id = None
        self._bit_combos = None
        self._corr = None

    @property
    def multiplicity(self):
        return len(self.clusters)

    @property
    def num_sites(self):
        return len(self.base_cluster)

    def __len__(self):
        return len(self.bit_combos)

    @property
    def bit_combos(self):
        """orbit.py:137-155 -- symmetry classes of site-function labelings."""
        if self._bit_combos is None:
            bits = [range(arr.shape[0]) for arr in self.basis_arrays]
            perms = np.asarray(self.cluster_permutations, dtype=np.int64)
            all_combos, seen = [], set()
            for combo in itertools.product(*bits):
                if combo in seen:
                    continue
                combo_a = np.array(combo, dtype=np.int32)
                new_bits = np.unique(combo_a[perms], axis=0)
                for row in new_bits:
                    seen.add(tuple(int(x) for x in row))
                all_combos.append(new_bits.astype(np.int32))
            self._bit_combos = tuple(all_combos)
        return self._bit_combos

    @property
    def bit_combo_multiplicities(self):
        return [len(c) for c in self.bit_combos]

    @property
    def correlation_tensors(self):
        """orbit.py:217-249."""
        if self._corr is None:
            shape = tuple(arr.shape[1] for arr in self.basis_arrays)
            out = np.zeros((len(self.bit_combos), *shape))
            for i, combos in enumerate(self.bit_combos):
                for bits in combos:
                    out[i] += reduce(lambda a, b: np.tensordot(a, b, axes=0),
                                     (self.basis_arrays[j][b] for j, b in enumerate(bits)))
                out[i] /= len(combos)
            self._corr = out.astype(np.float64)
        return self._corr

    @property
    def flat_correlation_tensors(self):
        """orbit.py:251-266."""
        ct = self.correlation_tensors
        return np.ascontiguousarray(ct.reshape(ct.shape[0], -1)).astype(np.float64)

    @property
    def flat_tensor_indices(self):
        """orbit.py:268-275 -- C-order strides of the species axes."""
        ct = self.correlation_tensors
        ind = np.cumprod(np.append(ct.shape[2:], 1)[::-1])[::-1]
        return np.ascontiguousarray(ind, dtype=np.int32)


@dataclass
class OrbitIndices:
    """Mirror of the reference's OrbitIndices named tuple (arrays only)."""

    arrays: tuple
    container: object = None


class ClusterSubspace:
    """Geometry-built stand-in for ``smol.cofe.ClusterSubspace`` (tables only)."""

    def __init__(self, prim: PrimCell, orbits, basis="sinusoid", orthonormal=True):
        self.prim = prim
        self.orbits = list(orbits)
        self.basis_type = basis
        self.orthonormal = orthonormal
        oid, bid, ncl = 1, 1, 1
        for orb in self.orbits:  # Orbit.assign_ids semantics (clusterspace.py:1295-1310)
            orb.id, orb.bit_id = oid, bid
            oid += 1
            bid += len(orb)
            ncl += orb.multiplicity
        self.num_orbits = oid
        self.num_corr_functions = bid
        self.num_clusters = ncl
        self.external_terms = []   # EwaldTerm instances (clusterspace.py add_external_term)
        self._cache = {}

    def add_external_term(self, term):
        self.external_terms.append(term)

    # ---- reference-named properties ------------------------------------------------
    @property
    def orbit_multiplicities(self):
        """clusterspace.py:384-387."""
        return np.array([1] + [o.multiplicity for o in self.orbits])

    @property
    def function_orbit_ids(self):
        ids = [0]
        for o in self.orbits:
            ids += len(o) * [o.id]
        return np.array(ids)

    @property
    def function_ordering_multiplicities(self):
        return np.array([1] + [m for o in self.orbits for m in o.bit_combo_multiplicities])

    @property
    def function_total_multiplicities(self):
        """clusterspace.py:437-450."""
        return self.orbit_multiplicities[self.function_orbit_ids] * \
            self.function_ordering_multiplicities

    # ---- construction ----------------------------------------------------------------
    @classmethod
    def from_cutoffs(cls, prim: PrimCell, cutoffs: dict, basis="sinusoid", orthonormal=True):
        """Enumerate orbits: ``cutoffs = {2: 6.0, 3: 3.5, 4: 3.0}`` (diameters in A)."""
        ops = prim.space_group()
        L = prim.lattice
        f = prim.frac_coords
        active = [b for b in range(prim.num_sites) if len(prim.site_spaces[b]) > 1]
        farrs = {b: site_function_array(len(prim.site_spaces[b]), basis, orthonormal)
                 for b in active}
        cart0 = f @ L

        def cart(site):
            b, n1, n2, n3 = site
            return cart0[b] + np.array([n1, n2, n3], dtype=float) @ L

        # map (R f_b + t) -> (b', integer shift) for every op once
        op_maps = []
        for R, t in ops:
            m = {}
            for b in range(prim.num_sites):
                img = R @ f[b] + t
                d = img[None, :] - f
                sh = np.round(d)
                hit = np.where(np.all(np.abs(d - sh) < SITE_TOL, axis=1))[0][0]
                m[b] = (int(hit), sh[hit].astype(np.int64))
            op_maps.append((R, m))

        def apply(opm, site):
            R, m = opm
            b, n1, n2, n3 = site
            b2, sh = m[b]
            n = sh + R @ np.array([n1, n2, n3], dtype=np.int64)
            return (b2, int(n[0]), int(n[1]), int(n[2]))

        def shift(cluster, d):
            return [(b, n1 - d[0], n2 - d[1], n3 - d[2]) for (b, n1, n2, n3) in cluster]

        def canon(cluster):
            best = None
            for anchor in cluster:
                c = tuple(sorted(shift(cluster, anchor[1:])))
                if best is None or c < best:
                    best = c
            return best

        maxcut = max([0.0] + [float(v) for k, v in cutoffs.items() if k >= 2])
        # neighbour shell around cell 0
        heights = 1.0 / np.linalg.norm(np.linalg.inv(L), axis=0)
        m = int(math.ceil(maxcut / heights.min())) + 1
        cand = [(b, i, j, k) for b in active for i in range(-m, m + 1)
                for j in range(-m, m + 1) for k in range(-m, m + 1)]
        cand_cart = np.array([cart(s) for s in cand])

        orbits = []
        seen_keys = set()

        def add_orbit(base):
            # all equivalent clusters (mod translation) with carried site order
            keyset, clusters, perms = {}, [], []
            base_key = canon(base)
            for opm in op_maps:
                img = [apply(opm, s) for s in base]
                k = canon(img)
                if k not in keyset:
                    keyset[k] = True
                    clusters.append(img)
                if k == base_key:
                    # permutation: img[i] + tau == base[p[i]]
                    for anchor in base:
                        for a2 in img:
                            tau = tuple(anchor[1 + q] - a2[1 + q] for q in range(3))
                            sh = [(b, n1 + tau[0], n2 + tau[1], n3 + tau[2])
                                  for (b, n1, n2, n3) in img]
                            if sorted(sh) == sorted(base):
                                p = tuple(base.index(s) for s in sh)
                                if p not in perms:
                                    perms.append(p)
            orbit_key = min(keyset)
            if orbit_key in seen_keys:
                return
            seen_keys.add(orbit_key)
            pts = np.array([cart(s) for s in base])
            diam = 0.0 if len(base) == 1 else max(
                np.linalg.norm(pts[i] - pts[j]) for i in range(len(base)) for j in range(i))
            orbits.append(Orbit(list(base), clusters, sorted(perms), diam,
                                tuple(farrs[s[0]] for s in base)))

        for b in active:  # point orbits
            add_orbit([(b, 0, 0, 0)])
        for size in sorted(k for k in cutoffs if k >= 2):
            cut = float(cutoffs[size]) + 1e-8
            for b in active:
                anchor = (b, 0, 0, 0)
                d0 = np.linalg.norm(cand_cart - cart(anchor), axis=1)
                near = [i for i in np.where(d0 <= cut)[0] if cand[i] != anchor]
                for combo in itertools.combinations(near, size - 1):
                    pts = cand_cart[list(combo)]
                    okc = True
                    for i in range(len(combo)):
                        for j in range(i):
                            if np.linalg.norm(pts[i] - pts[j]) > cut:
                                okc = False
                                break
                        if not okc:
                            break
                    if okc:
                        add_orbit([anchor] + [cand[i] for i in combo])
        orbits.sort(key=lambda o: (o.num_sites, round(o.diameter, 6), -o.multiplicity, len(o)))
        return cls(prim, orbits, basis, orthonormal)

    # ---- supercell tables --------------------------------------------------------------
    @staticmethod
    def lattice_points(scmatrix):
        """Integer lattice points inside the supercell (row-vector convention)."""
        S = np.asarray(scmatrix, dtype=np.int64).reshape(3, 3)
        det = int(round(abs(np.linalg.det(S))))
        corners = np.array(list(itertools.product((0, 1), repeat=3))) @ S
        lo, hi = corners.min(0), corners.max(0)
        grid = np.array(list(itertools.product(*[range(lo[i], hi[i] + 1) for i in range(3)])))
        frac = grid @ np.linalg.inv(S)
        inside = np.all((frac > -1e-9) & (frac < 1 - 1e-9), axis=1)
        pts = grid[inside]
        assert len(pts) == det, (len(pts), det)
        return pts

    def supercell_site_index(self, scmatrix):
        """Return ``(pts, lookup)``; site index = b * ncells + cell  (basis-major, as pymatgen)."""
        S = np.asarray(scmatrix, dtype=np.int64).reshape(3, 3)
        pts = self.lattice_points(S)
        Sinv = np.linalg.inv(S)
        table = {tuple(p): i for i, p in enumerate(pts)}

        def lookup(b, n):
            fr = np.asarray(n, dtype=float) @ Sinv
            fr = fr - np.floor(fr + 1e-9)
            p = np.rint(fr @ S).astype(np.int64)
            return b * len(pts) + table[tuple(p)]

        return pts, lookup

    def get_orbit_indices(self, scmatrix) -> OrbitIndices:
        """clusterspace.py:1312-1366: per orbit int32 ``[multiplicity*ncells, size]``."""
        S = np.asarray(scmatrix, dtype=np.int64).reshape(3, 3)
        key = tuple(S.ravel().tolist())
        if key in self._cache:
            return self._cache[key]
        pts, _ = self.supercell_site_index(S)
        ncell = len(pts)
        Sinv = np.linalg.inv(S)
        table = {tuple(p): i for i, p in enumerate(pts)}
        # cell index of (pts[t] + n) for arbitrary n: wrap through fractional coords
        arrays = []
        for orb in self.orbits:
            rows = np.empty((orb.multiplicity * ncell, orb.num_sites), dtype=np.int32)
            for ci, cl in enumerate(orb.clusters):
                for si, (b, n1, n2, n3) in enumerate(cl):
                    n = pts + np.array([n1, n2, n3])
                    fr = n @ Sinv
                    fr = fr - np.floor(fr + 1e-9)
                    wrapped = np.rint(fr @ S).astype(np.int64)
                    cells = np.array([table[tuple(p)] for p in wrapped])
                    rows[ci * ncell:(ci + 1) * ncell, si] = b * ncell + cells
            arrays.append(np.ascontiguousarray(rows, dtype=np.int32))
        out = OrbitIndices(tuple(arrays))
        self._cache[key] = out
        return out

    def supercell_size(self, scmatrix):
        return int(round(abs(np.linalg.det(np.asarray(scmatrix, dtype=float)))))

    def num_supercell_sites(self, scmatrix):
        return self.supercell_size(scmatrix) * self.prim.num_sites

    def allowed_species(self, scmatrix):
        """Per supercell site tuple of species labels (basis-major order)."""
        nc = self.supercell_size(scmatrix)
        return [self.prim.site_spaces[b] for b in range(self.prim.num_sites) for _ in range(nc)]

    def supercell_frac_cart(self, scmatrix):
        """Cartesian coordinates of all supercell sites, basis-major."""
        pts, _ = self.supercell_site_index(scmatrix)
        L = self.prim.lattice
        out = []
        for b in range(self.prim.num_sites):
            out.append((self.prim.frac_coords[b][None, :] + pts) @ L)
        return np.concatenate(out, axis=0)


# --------------------------------------------------------------------------------------
# cluster expansion helpers (cofe/expansion.py)
# --------------------------------------------------------------------------------------
def eci_from_coefs(subspace: ClusterSubspace, coefs) -> np.ndarray:
    """cofe/expansion.py:171-183."""
    return np.asarray(coefs, dtype=np.float64) / subspace.function_total_multiplicities


def cluster_interaction_tensors(subspace: ClusterSubspace, coefs):
    """cofe/expansion.py:185-201 -- ``(coefs[0], I_1, I_2, ...)`` one tensor per orbit."""
    coefs = np.asarray(coefs, dtype=np.float64)
    eci = eci_from_coefs(subspace, coefs)
    return (coefs[0],) + tuple(
        sum(m * eci[orb.bit_id + i] * tensor
            for i, (m, tensor) in enumerate(zip(orb.bit_combo_multiplicities,
                                                orb.correlation_tensors)))
        for orb in subspace.orbits)


class ClusterExpansion:
    """Minimal stand-in for ``smol.cofe.ClusterExpansion`` (cofe/expansion.py): subspace + coefs."""

    def __init__(self, cluster_subspace, coefficients):
        self.cluster_subspace = cluster_subspace
        self.coefs = np.asarray(coefficients, dtype=np.float64)
        n_ext = len(cluster_subspace.external_terms)
        if len(self.coefs) != cluster_subspace.num_corr_functions + n_ext:
            raise AttributeError("Feature matrix shape does not match the number of coefficients.")

    @property
    def eci(self):
        n_ext = len(self.cluster_subspace.external_terms)
        coefs = self.coefs[:-n_ext] if n_ext else self.coefs
        return eci_from_coefs(self.cluster_subspace, coefs)

    @property
    def cluster_interaction_tensors(self):
        n_ext = len(self.cluster_subspace.external_terms)
        coefs = self.coefs[:-n_ext] if n_ext else self.coefs
        return cluster_interaction_tensors(self.cluster_subspace, coefs)


class EwaldTerm:
    """Parameters of the Ewald external term (cofe/extern/ewald.py:25-63)."""

    def __init__(self, eta=None, real_space_cut=None, recip_space_cut=None, use_term="total"):
        if use_term != "total":
            raise NotImplementedError("only the total Ewald matrix is generated here")
        self.eta, self.real_space_cut, self.recip_space_cut = eta, real_space_cut, recip_space_cut
        self.use_term = use_term


# --------------------------------------------------------------------------------------
# Ewald tables
# --------------------------------------------------------------------------------------
def ewald_indices(subspace: ClusterSubspace, scmatrix):
    """Index layout of ``EwaldTerm.get_ewald_structure`` (cofe/extern/ewald.py:64-100).

    Returns ``(inds int32[N, max_species], ewald_site_of_row int[E], charge[E])``; rows are
    numbered consecutively site by site, ``-1`` for vacancies (label ``"Vac"``) / padding.
    """
    spaces = subspace.allowed_species(scmatrix)
    width = max(len(s) for s in spaces)
    inds = np.full((len(spaces), width), -1, dtype=np.int32)
    site_of_row, charge = [], []
    for k, space in enumerate(spaces):
        for i, sp in enumerate(space):
            if sp in ("Vac", "Vacancy", "vacancy"):
                continue
            inds[k, i] = len(site_of_row)
            site_of_row.append(k)
            charge.append(float(subspace.prim.charges[sp]))
    return inds, np.array(site_of_row), np.array(charge)


def _ewald_rows_numpy(cart, rows, gv, coef, tv, eta, real_cut, vol):
    """Pair kernel between the origin sites ``rows`` and every site: reciprocal + real space sums."""
    from scipy.special import erfc
    k_rows = np.zeros((len(rows), len(cart)))
    for bi, r0 in enumerate(rows):
        d = cart - cart[r0]
        for lo in range(0, len(gv), 4096):
            ph = d @ gv[lo:lo + 4096].T
            k_rows[bi] += (2 * math.pi / vol) * (np.cos(ph) @ coef[lo:lo + 4096])
    rse = math.sqrt(eta)
    for bi, r0 in enumerate(rows):
        d = (cart - cart[r0])[:, None, :] + tv[None, :, :]
        r = np.sqrt(np.sum(d * d, axis=2))
        mask = (r > 1e-8) & (r <= real_cut)
        rr = np.where(mask, r, 1.0)
        k_rows[bi] += 0.5 * np.sum(np.where(mask, erfc(rse * rr) / rr, 0.0), axis=1)
    return k_rows


def _ewald_rows_gpu(cart, rows, gv, coef, tv, eta, real_cut, vol):
    """The same sums on the GPU (``lmc_ewald_site_kernel``, one block per site pair)."""
    import torch
    from . import _capi as capi
    lib = capi.load()
    dev = torch.device("cuda", torch.cuda.current_device())

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    cart_d, rows_d = up(cart, np.float64), up(rows, np.int32)
    pad3, pad1 = np.zeros((1, 3)), np.zeros(1)      # (an empty sum still hands the library a valid pointer)
    gv_d, gc_d, tv_d = up(gv if len(gv) else pad3, np.float64), up(coef if len(coef) else pad1, np.float64), \
        up(tv if len(tv) else pad3, np.float64)
    out = torch.empty((len(rows), len(cart)), dtype=torch.float64, device=dev)
    capi.check(lib.lmc_ewald_site_kernel(cart_d.data_ptr(), len(cart), rows_d.data_ptr(), len(rows), gv_d.data_ptr(),
                                         gc_d.data_ptr(), len(gv), tv_d.data_ptr(), len(tv), float(eta),
                                         float(real_cut), float(vol), out.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream))
    return out.cpu().numpy()


def ewald_matrix(subspace: ClusterSubspace, scmatrix, eta=None, real_cut=None, recip_cut=None,
                 acc=12.0, backend="auto", term="total"):
    """Ewald pair matrix M (eV) with ``E = sum(M[occupied][:, occupied])``; ``term`` selects the part as
    ``EwaldTerm.use_term`` does (cofe/extern/ewald.py:168-177): "total" = reciprocal + real + point, or one of them
    ("point" is the diagonal self term).

    Same decomposition as pymatgen's ``EwaldSummation.total_energy_matrix`` used by
    ``cofe/extern/ewald.py:159-177``: reciprocal + real (+ point/self term on the diagonal),
    no charged-cell correction.  Units eV with e^2/(4 pi eps0) = 14.399645 eV*A.
    ``backend``: "gpu" (``lmc_ewald_site_kernel``), "numpy", or "auto" = the GPU when one is present.
    """
    conv = 14.39964547842567
    S = np.asarray(scmatrix, dtype=np.int64).reshape(3, 3)
    L = S @ subspace.prim.lattice
    vol = abs(np.linalg.det(L))
    cart = subspace.supercell_frac_cart(S)
    nsite = len(cart)
    inds, site_of_row, q = ewald_indices(subspace, S)
    if eta is None:
        eta = (nsite * 0.01 / vol) ** (1 / 3) * math.pi  # pymatgen's default heuristic
    if real_cut is None:
        real_cut = math.sqrt(-math.log(10.0 ** -acc) / eta)
    if recip_cut is None:
        recip_cut = 2 * math.sqrt(eta) * math.sqrt(-math.log(10.0 ** -acc))
    # displacement from site 0 of each basis class is enough only for translation-invariant
    # sets; do the general thing on unique displacement vectors instead.
    Linv = np.linalg.inv(L)
    rec = 2 * math.pi * Linv.T  # rows = reciprocal vectors
    pts, _ = subspace.supercell_site_index(S)
    ncell, nb = len(pts), subspace.prim.num_sites
    rows = np.arange(nb) * ncell  # image of every basis site in supercell cell 0
    # reciprocal sum: E_rec = 1/2 sum_ij qi qj (4 pi/V) sum_G coef cos(G.(ri-rj)); k carries 1/2
    heights_r = 1.0 / np.linalg.norm(np.linalg.inv(rec), axis=0)
    mr = np.ceil(recip_cut / heights_r).astype(int) + 1
    gi = np.array(list(itertools.product(*[range(-m_, m_ + 1) for m_ in mr])))
    gv = gi @ rec
    g2 = np.sum(gv * gv, axis=1)
    keep = (g2 > 1e-12) & (g2 <= recip_cut ** 2)
    gv, g2 = gv[keep], g2[keep]
    coef = np.exp(-g2 / (4 * eta)) / g2
    # real-space sum (pairs at zero distance -- the site itself or an overlaid species -- skipped)
    heights = 1.0 / np.linalg.norm(Linv, axis=0)
    mm = np.ceil(real_cut / heights).astype(int) + 1
    ti = np.array(list(itertools.product(*[range(-m_, m_ + 1) for m_ in mm])))
    tv = ti @ L
    if backend == "auto":
        try:
            import torch
            backend = "gpu" if torch.cuda.is_available() else "numpy"
        except ImportError:
            backend = "numpy"
    if backend not in ("gpu", "numpy"):
        raise ValueError("backend must be 'auto', 'gpu' or 'numpy'")
    if term not in ("total", "reciprocal", "real", "point"):
        raise ValueError("term must be one of 'total', 'reciprocal', 'real', 'point'")     # cofe/extern/ewald.py:48-52
    if term in ("real", "point"):
        gv, coef = gv[:0], coef[:0]
    if term in ("reciprocal", "point"):
        tv = tv[:0]
    k_rows = (_ewald_rows_gpu if backend == "gpu" else _ewald_rows_numpy)(cart, rows, gv, coef, tv, eta, real_cut, vol)
    # expand by translation invariance: K[(b,c),(b2,c2)] = k_rows[b][(b2, c2 - c)]
    Sinv = np.linalg.inv(S)
    # [c, c2] -> index of the supercell lattice point pts[c2] - pts[c] (wrapped into the supercell)
    diff = (pts[None, :, :] - pts[:, None, :]).reshape(-1, 3)
    fr = diff @ Sinv
    fr = fr - np.floor(fr + 1e-9)
    wrapped = np.rint(fr @ S).astype(np.int64)
    lo_, span = pts.min(axis=0), pts.max(axis=0) - pts.min(axis=0) + 1

    def key(v):
        v = v - lo_
        return (v[:, 0] * span[1] + v[:, 1]) * span[2] + v[:, 2]
    pkeys = key(pts)
    order = np.argsort(pkeys)
    pos = np.searchsorted(pkeys[order], key(wrapped))
    if (pos >= ncell).any() or (pkeys[order][np.minimum(pos, ncell - 1)] != key(wrapped)).any():
        raise RuntimeError("wrapped lattice point outside the supercell point set")
    cellsub = order[pos].reshape(ncell, ncell)
    k_full = np.empty((nsite, nsite))
    for b in range(nb):
        for b2 in range(nb):
            k_full[b * ncell:(b + 1) * ncell, b2 * ncell:(b2 + 1) * ncell] = \
                k_rows[b][b2 * ncell + cellsub]
    k_rec, k_real = k_full, 0.0
    m = np.outer(q, q) * (k_rec + k_real)[np.ix_(site_of_row, site_of_row)]
    if term in ("total", "point"):
        m[np.arange(len(q)), np.arange(len(q))] += -q * q * math.sqrt(eta / math.pi)
    m = 0.5 * (m + m.T)  # exact symmetry (the delta path reads one triangle only)
    return np.ascontiguousarray(m * conv, dtype=np.float64), inds
