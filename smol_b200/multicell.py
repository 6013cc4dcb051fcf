"""MulticellSampler: sampling over several supercell shapes at once on the GPU.

Mirror of ``Sampler([MulticellMetropolis(mckernels, temperature, ...)], container)`` in the reference
(``smol/moca/kernel/base.py:439-722``, ``smol/moca/kernel/metropolis.py:102-175``) -- the kernel behind smol's
SQS generation (``smol/capp/generate/special/sqs.py:523-540``).  One chain owns ONE occupancy PER supercell shape
("kernel"); most steps are ordinary Metropolis steps of the shape the chain currently sits in, and every
``kernel_hop_period``-th step is a hop attempt: a shape k' is drawn, ONE step is proposed in k' from k's own
stored occupancy, and it is accepted with the enthalpy of the new k' state minus the enthalpy of the current
shape's state (``base.py:612-622, 661-682``).

Device side: one model (``LmcEngine``) and one state array ``occ[k] [W][row]``, ``features[k]``, ``enthalpy[k]``
per shape, all W walkers each.  A launch of shape k advances only the walkers whose byte in
``LmcRunConfig.walker_mask_dev`` is set (those sitting in k); a hop into k' is a one-step launch of k' with
``accept_offset_dev[w] = H_k'[w] - H_current[w]``.  The bookkeeping between launches (current shape, masks,
offsets, samples) is a handful of torch ops on the device; nothing is synchronised with the host inside ``run``.

Randomness: the shape and hop-period choices are state independent and come from
``numpy.random.default_rng(seed).choice(..., p=...)`` exactly as in the reference (one uniform each, in the
reference's order: the initial period in the constructor, then per hop the shape followed by the next period).
Proposals and acceptance uniforms are counter based like everywhere in this engine: Philox keyed by the sub-kernel's
seed, counter = the chain's global step index (the reference draws a hop's uniform from the multicell generator,
data dependent).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _capi as capi
from .container import SampleContainer
from .sampler import _USHERS, kB


class _ChoiceStream:
    """``rng.choice(n, p=p)`` draws of one walker, taken from its generator in blocks.

    ``Generator.choice`` with ``p`` given is ``cdf.searchsorted(rng.random(), side="right")`` with
    ``cdf = p.cumsum(); cdf /= cdf[-1]``, i.e. one double per call -- so a block of ``rng.random(n)`` is the
    same stream as n calls."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.buf = np.empty(0)
        self.pos = 0

    def uniform(self):
        if self.pos >= len(self.buf):
            self.buf = self.rng.random(256)
            self.pos = 0
        u = self.buf[self.pos]
        self.pos += 1
        return u


def _cdf(p):
    c = np.cumsum(np.asarray(p, dtype=np.float64))
    return c / c[-1]


class MulticellSampler:
    def __init__(self, ensembles, temperature, step_type="swap", nwalkers=1, seeds=None, kernel_seeds=None,
                 kernel_temperatures=None, kernel_probabilities=None, kernel_hop_periods=5,
                 kernel_hop_probabilities=None, sublattice_probabilities=None, walker_id_base=0, device=None,
                 kB_=kB, share_visited=True):
        from .engine import LmcEngine
        ensembles = list(ensembles)
        if not ensembles:
            raise ValueError("at least one ensemble is needed")
        if any(e.num_sites != ensembles[0].num_sites for e in ensembles):
            raise ValueError("All ensembles must have the same number of sites.")                 # base.py:485-489
        if any(not np.allclose(e.natural_parameters, ensembles[0].natural_parameters) for e in ensembles):
            raise ValueError("All ensembles must have the same natural parameters.")              # base.py:491-495
        from .processor import DistanceProcessor
        dist = [isinstance(e.processor, DistanceProcessor) for e in ensembles]
        if any(dist) and not all(dist):
            raise ValueError("either all or none of the ensembles wrap a distance processor")
        # distance processors (processor/distance.py; the ensembles of smol's SQS generator): features are the distance
        # vector, each shape keeps the walkers' running correlation vector beside it
        self._dist = [e.processor for e in ensembles] if all(dist) else None
        K = len(ensembles)
        if kernel_probabilities is not None:
            if sum(kernel_probabilities) != 1.0:
                raise ValueError("The kernel_probabilities do not sum to 1.")                       # base.py:498-499
            if len(kernel_probabilities) != K:
                raise ValueError("The length of kernel_probabilities must be equal to the number of mckernels.")
            self._kernel_p = np.array(kernel_probabilities, dtype=np.float64)
        else:
            self._kernel_p = np.array([1.0 / K] * K)
        self._hop_periods = np.array([kernel_hop_periods] if isinstance(kernel_hop_periods, (int, np.integer))
                                     else kernel_hop_periods, dtype=int)
        if (self._hop_periods < 1).any():
            raise ValueError("kernel_hop_periods must be positive")
        if kernel_hop_probabilities is not None:
            if sum(kernel_hop_probabilities) != 1.0:
                raise ValueError("The kernel_hop_probabilities do not sum to 1.")                   # base.py:516-517
            if len(kernel_hop_probabilities) != len(self._hop_periods):
                raise ValueError("The length of kernel_hop_periods and kernel_hop_probabilities does not match.")
            self._hop_p = np.array(kernel_hop_probabilities, dtype=np.float64)
        else:
            self._hop_p = np.array([1.0 / len(self._hop_periods)] * len(self._hop_periods))
        ukey = step_type.lower().replace("-", "").replace("_", "")
        if ukey not in ("flip", "swap"):
            raise NotImplementedError("MulticellSampler proposes flip or swap steps")
        self._usher = _USHERS[ukey]
        self.step_type = step_type
        self.ensembles = ensembles
        self.nwalkers = W = int(nwalkers)
        if seeds is None:
            seeds = [int(x) for x in np.random.SeedSequence().generate_state(W, dtype=np.uint64)]
        if len(seeds) != W:
            raise ValueError("Number of seeds does not match number of kernels!")
        self.seeds = [int(s) for s in seeds]
        if kernel_seeds is None:   # one Philox key per (shape, walker)
            kernel_seeds = [[(s * 0x9E3779B97F4A7C15 + k + 1) & 0xFFFFFFFFFFFFFFFF for s in self.seeds] for k in range(K)]
        self.kernel_seeds = np.array(kernel_seeds, dtype=np.uint64).reshape(K, W)
        self.kB = kB_
        self.temperature = float(temperature)
        self.kernel_temperatures = np.full(K, self.temperature) if kernel_temperatures is None \
            else np.array(kernel_temperatures, dtype=np.float64)
        self.walker_id_base = int(walker_id_base)
        # What the reference does (default): MCKernel.single_step keeps the array it was handed as its trace occupancy
        # without copying (base.py:162) and the sampler hands every kernel the chain's one live row, so a shape that
        # has taken an ordinary step -- or had a hop away from it rejected (base.py:671-674) -- shares the chain's live
        # occupancy from then on: a later hop into it starts from the CURRENT occupancy string re-read in that shape,
        # and only shapes never visited keep the occupancy given at the start.  False: every shape keeps an
        # occupancy of its own (the intent documented at base.py:665-666).  Both are restated in the oracle; the
        # default is pinned against the reference's class (tests/golden/ref_python_steps.npz).
        self.share_visited = bool(share_visited)
        self.engines = [LmcEngine(e.packed_model(sublattice_probabilities=sublattice_probabilities), device=device)
                        for e in ensembles]
        # hop schedule state per walker (base.py:530-533: the first period is drawn in the constructor)
        self._streams = [_ChoiceStream(s) for s in self.seeds]
        self._hop_cdf, self._kernel_cdf = _cdf(self._hop_p), _cdf(self._kernel_p)
        self._period = np.array([self._draw_period(w) for w in range(W)], dtype=np.int64)
        self._counter = np.ones(W, dtype=np.int64)
        self._step_counter = 0
        self._state = None
        N, F = ensembles[0].num_sites, len(ensembles[0].natural_parameters)
        shapes = {"occupancy": ((N,), np.int32), "features": ((F,), np.float64), "enthalpy": ((1,), np.float64),
                  "accepted": ((1,), bool), "n_accepted": ((), np.int32), "temperature": ((1,), np.float64),
                  "kernel_index": ((1,), np.int64)}
        self._container = SampleContainer(ensembles[0], W, shapes, dict(getattr(ensembles[0], "thermo_boundaries", {})))
        self.launches = 0

    # ------------------------------------------------------------------------------------------
    def _draw_period(self, w):
        return int(self._hop_periods[np.searchsorted(self._hop_cdf, self._streams[w].uniform(), side="right")])

    def _draw_kernel(self, w):
        return int(np.searchsorted(self._kernel_cdf, self._streams[w].uniform(), side="right"))

    @property
    def samples(self):
        return self._container

    def clear_samples(self):
        self._container.clear()

    def efficiency(self, discard=0, flat=True):
        return self.samples.sampling_efficiency(discard=discard, flat=flat)

    def current_kernel_indices(self):
        return self._state["cur"].cpu().numpy() if self._state else np.zeros(self.nwalkers, dtype=np.int64)

    def current_occupancies(self):
        """int32 ``[W, K, N]``: every walker's occupancy in every shape."""
        st = self._state
        return np.stack([e.occupancy_to_int32(o, self.nwalkers, e.row_stride).cpu().numpy()
                         for e, o in zip(self.engines, st["occ"])], axis=1)

    # ------------------------------------------------------------------------------------------
    def _launch(self, k, nsteps, step_begin, mask, beta, offset=None):
        st, eng = self._state, self.engines[k]
        cfg = capi.LmcRunConfig()
        cfg.num_walkers, cfg.walker_id_base = self.nwalkers, self.walker_id_base
        cfg.usher, cfg.kernel = self._usher, capi.LMC_KERNEL_METROPOLIS
        cfg.num_samples, cfg.thin_by = 1, int(nsteps)
        cfg.group_size, cfg.block_threads, cfg.spec_mode = 0, 0, 1
        cfg.step_begin = int(step_begin)
        cfg.seeds_dev, cfg.beta_dev = st["kseeds"][k].data_ptr(), beta.data_ptr()
        cfg.occ_dev, cfg.features_dev, cfg.enthalpy_dev = \
            st["occ"][k].data_ptr(), st["feat"][k].data_ptr(), st["enth"][k].data_ptr()
        cfg.trace_accepted_dev, cfg.trace_naccepted_dev = st["acc_tmp"].data_ptr(), st["nacc_tmp"].data_ptr()
        cfg.walker_mask_dev = mask.data_ptr()
        cfg.accept_offset_dev = offset.data_ptr() if offset is not None else None
        if self._dist is not None:
            dt = eng.distance_tables(self._dist[k])
            cfg.dist_mode, cfg.dist_num_groups, cfg.dist_tol = 1, dt["ngrp"], dt["tol"]
            cfg.dist_target_dev, cfg.dist_group_off_dev = dt["target"].data_ptr(), dt["goff"].data_ptr()
            cfg.dist_group_idx_dev, cfg.dist_group_diam_dev = dt["gidx"].data_ptr(), dt["gdiam"].data_ptr()
            cfg.dist_vector_dev = st["vec"][k].data_ptr()
        eng.run(cfg)
        self.launches += 1

    def run(self, nsteps, initial_occupancies=None, thin_by=1, progress=False):
        import torch
        K, W = len(self.engines), self.nwalkers
        e0 = self.engines[0]
        N, F, dev = e0.N, e0.F, e0.device
        if initial_occupancies is None:
            if self._state is None:
                raise RuntimeError("There are no saved samples to obtain the initial occupancies."
                                   "These must be provided.")
        else:
            occ = np.asarray(initial_occupancies)
            if occ.ndim == 2 and W == 1:
                occ = occ[None]
            if occ.shape != (W, K, N):
                raise AttributeError("The given initial occcupancies have incompompatible dimensions. "
                                     f"Shape should be {(W, K, N)}.")
            # (set_aux_state, base.py:694-716: one occupancy per shape; the chain starts in shape 0)
            st = dict(occ=[], feat=[], enth=[], kseeds=[], beta=[], vec=[])
            for k, eng in enumerate(self.engines):
                o = eng.upload_occupancy(np.ascontiguousarray(occ[:, k, :]))
                f, h = eng.full_features(o)
                if self._dist is not None:     # extensive features -> distance vector (in place) + vector per supercell
                    st["vec"].append(eng.distance_init(self._dist[k], f, h))
                st["occ"].append(o); st["feat"].append(f); st["enth"].append(h)
                st["kseeds"].append(torch.from_numpy(self.kernel_seeds[k].view(np.int64).copy()).to(dev))
                st["beta"].append(torch.full((W,), 1.0 / (self.kB * self.kernel_temperatures[k]), dtype=torch.float64, device=dev))
            st["beta_mc"] = torch.full((W,), 1.0 / (self.kB * self.temperature), dtype=torch.float64, device=dev)
            st["cur"] = torch.zeros((W,), dtype=torch.int64, device=dev)
            st["acc_tmp"] = torch.zeros((W,), dtype=torch.uint8, device=dev)
            st["nacc_tmp"] = torch.zeros((W,), dtype=torch.int32, device=dev)
            st["acc_last"] = torch.ones((W,), dtype=torch.uint8, device=dev)
            st["ktemp"] = torch.from_numpy(self.kernel_temperatures.copy()).to(dev)
            st["alias"] = [torch.zeros((W,), dtype=torch.bool, device=dev) for _ in range(K)]
            self._state = st
        st = self._state
        if nsteps % thin_by != 0:
            warnings.warn(f"The number of steps {nsteps} is not a multiple of thin_by  {thin_by}. "
                          f"The last {nsteps % thin_by} will be ignored.", category=RuntimeWarning)
        S = nsteps // thin_by
        ar = torch.arange(W, device=dev)
        tr = dict(occupancy=torch.empty((S, W, N), dtype=torch.int8, device=dev),
                  features=torch.empty((S, W, F), dtype=torch.float64, device=dev),
                  enthalpy=torch.empty((S, W), dtype=torch.float64, device=dev),
                  accepted=torch.empty((S, W), dtype=torch.uint8, device=dev),
                  n_accepted=torch.zeros((S, W), dtype=torch.int32, device=dev),
                  temperature=torch.empty((S, W), dtype=torch.float64, device=dev),
                  kernel_index=torch.empty((S, W), dtype=torch.int64, device=dev))
        nacc = torch.zeros((W,), dtype=torch.int32, device=dev)
        t, total = 0, S * thin_by
        while t < total:
            to_hop = self._period - self._counter          # ordinary steps each walker takes before its next hop
            seg = int(min(to_hop.min(), thin_by - t % thin_by))
            if seg > 0:
                # ordinary steps of every walker in the shape it currently sits in (base.py:683-691)
                for k in range(K):
                    mask = (st["cur"] == k).to(torch.uint8)
                    self._launch(k, seg, self._step_counter + t, mask, st["beta"][k])
                    m = mask.bool()
                    st["acc_last"] = torch.where(m, st["acc_tmp"], st["acc_last"])
                    nacc += torch.where(m, st["nacc_tmp"], torch.zeros_like(nacc))
                    st["alias"][k] = st["alias"][k] | m
                self._counter += seg
                t += seg
            else:
                # hop attempts (base.py:661-682) of the walkers whose counter reached their period; the others
                # take one ordinary step
                hop = to_hop == 0
                target = np.array([self._draw_kernel(w) if hop[w] else -1 for w in range(W)], dtype=np.int64)
                hop_d, target_d = torch.from_numpy(hop).to(dev), torch.from_numpy(target).to(dev)
                cur0 = st["cur"].clone()
                h_cur = torch.stack(st["enth"])[cur0, ar]          # (tracked) enthalpy of the current shape's state
                live = torch.stack(st["occ"])[cur0, ar]            # the chain's live occupancy rows
                if not hop.all():
                    for k in range(K):
                        mask = ((cur0 == k) & ~hop_d).to(torch.uint8)
                        self._launch(k, 1, self._step_counter + t, mask, st["beta"][k])
                        m = mask.bool()
                        st["acc_last"] = torch.where(m, st["acc_tmp"], st["acc_last"])
                        nacc += torch.where(m, st["nacc_tmp"], torch.zeros_like(nacc))
                        st["alias"][k] = st["alias"][k] | m
                for k in sorted(set(int(x) for x in target[hop])):
                    m = target_d == k
                    if self.share_visited:
                        # a visited shape reads the live occupancy: re-read those rows in shape k and evaluate them
                        # in full there (base.py:612-616 computes the new state's features from scratch)
                        r = m & st["alias"][k]
                        st["occ"][k] = torch.where(r[:, None], live, st["occ"][k])
                        f_new, h_new = self.engines[k].full_features(st["occ"][k])
                        if self._dist is not None:
                            v_new = self.engines[k].distance_init(self._dist[k], f_new, h_new)
                            st["vec"][k] = torch.where(r[:, None], v_new, st["vec"][k])
                        st["feat"][k] = torch.where(r[:, None], f_new, st["feat"][k])
                        st["enth"][k] = torch.where(r, h_new, st["enth"][k])
                    mask = m.to(torch.uint8)
                    offset = st["enth"][k] - h_cur
                    self._launch(k, 1, self._step_counter + t, mask, st["beta_mc"], offset)
                    took = m & (st["acc_tmp"] != 0)
                    st["cur"] = torch.where(took, torch.full_like(st["cur"], k), st["cur"])
                    st["acc_last"] = torch.where(m, st["acc_tmp"], st["acc_last"])
                    nacc += torch.where(m, st["nacc_tmp"], torch.zeros_like(nacc))
                    if self.share_visited:
                        stay = m & ~took                           # rejected: the shape hopped FROM now holds the live row
                        for c in range(K):
                            st["alias"][c] = st["alias"][c] | (stay & (cur0 == c))
                for w in np.nonzero(hop)[0]:
                    self._period[w] = self._draw_period(w)     # base.py:678-682
                    self._counter[w] = 1
                self._counter[~hop] += 1
                t += 1
            if t % thin_by == 0:
                s = t // thin_by - 1
                cur = st["cur"]
                tr["occupancy"][s] = torch.stack(st["occ"])[cur, ar, :N]
                tr["features"][s] = torch.stack(st["feat"])[cur, ar]
                tr["enthalpy"][s] = torch.stack(st["enth"])[cur, ar]
                tr["accepted"][s] = st["acc_last"]
                tr["n_accepted"][s] = nacc
                tr["temperature"][s] = st["ktemp"][cur]
                tr["kernel_index"][s] = cur
                nacc.zero_()
        self._step_counter += total
        host = {k: v.cpu().numpy() for k, v in tr.items()}
        traces = {"occupancy": host["occupancy"], "features": host["features"],
                  "enthalpy": host["enthalpy"][:, :, None], "accepted": host["accepted"].astype(bool)[:, :, None],
                  "n_accepted": host["n_accepted"], "temperature": host["temperature"][:, :, None],
                  "kernel_index": host["kernel_index"][:, :, None]}
        self.samples.append(traces, thin_by)
