#!/usr/bin/env python
"""Headline benchmark: attempted MC steps/s of the lattice-MC hot path (BASELINE.json metric).

``--config {2,3,4,5}`` selects the BASELINE.json configuration (one definition each
in ``tests/workloads.py``).  Default = config 2, the configuration the metric is quoted on: binary FCC
8x8x8 (N = 512), cluster set S_fcc, canonical Metropolis swap at T = 1000 K, 4096 walkers per GPU, one
sample per sweep.

A bench "step" = one ``lmc_run`` launch advancing every walker of the GPU by ``samples_per_bench_step``
sampling intervals of ``thin_by`` attempted MC steps, writing the per-interval traces to HBM.

  value     device-resident throughput: CUDA events around the launches, state, tables and traces in HBM
            (``Sampler.run_device``); N > 1: the enthalpy traces of the timed region are all-gathered once behind
            the last launch (the path's only collective) and that time is added to the total
  e2e       the same metric through the public API with HOST buffers: per step the initial occupancies go
            host->device from page-locked int32 memory and every trace comes back to the host;
            ``Sampler.run(..., block=False)`` back to back, each step's result read on the host while the next
            step runs (all copies inside the timed region)
  roofline  bytes per attempted step x steps per launch / launch time vs the measured HBM copy bandwidth
            (MEASURED_PEAKS.json).  ``bytes_per_step`` is SURVEY 8(d)'s figure where the built kernel moves
            those bytes (configs 2 and 4) and the bytes of the BUILT algorithm where it replaces the reference's
            (configs 3 and 5: Ewald potential cache / site-kernel rows), both stated; ``traffic`` and ``limiter``
            quote the committed ncu capture of the same kernel (profiles/)
  cpu_baseline  the reference's own compiled Cython evaluators (oracle/_ref) under the restated smol step
            loop, one process per host core, bounded sample of the same configuration
``--impl reference`` runs only that CPU arm (rank 0) and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "attempted MC steps/sec (whole job)"
DTYPE = "f64 (energies/features), u8 (occupancy), u32 (rng)"


def workload(config_id):
    from tests import workloads as WK
    return WK.get(config_id)


# ------------------------------------------------------------------------------------------
# CPU arm: reference Cython evaluators + restated step loop, one process per host core
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    config_id, kind, walker0, nwalk, seconds = args
    wk = workload(config_id)
    occ = wk.initial_occupancies(nwalk, seed=1000 + walker0).astype(np.int32)
    if kind == "port":
        # C restatement (oracle/lmc_oracle.c): flip / swap, Metropolis / Wang-Landau
        from oracle import c_oracle as CO
        from oracle import lmc_oracle as O
        co = CO.COracle(O.Ensemble(wk.oracle_processor(False), wk.oracle_sublattices(),
                                   chemical_potentials=wk.chemical_potentials()))
        nsteps, total, t0 = 20000, 0, time.perf_counter()
        kw = dict(usher=wk.usher, nthreads=1, record=False, walker_base=walker0)
        if wk.kernel == "WangLandau":
            lo, hi = wk.window()
            kw["wl"] = dict(min=lo, max=hi, bin=wk.bin_size, flatness=wk.flatness, check=wk.check_period)
        else:
            kw["temperature"] = wk.temperature
        while time.perf_counter() - t0 < seconds:
            co.run(occ, nsteps, nsteps, np.arange(walker0, walker0 + nwalk), **kw)
            total += nsteps * nwalk
        return total, time.perf_counter() - t0
    # smol's compiled evaluators under the restated step loop (MCKernel.single_step + the sampler's
    # accumulate-on-accept, sampler.py:195-210)
    kernels = wk.oracle_kernels(np.arange(walker0, walker0 + nwalk), walker0=walker0, use_ref=(kind == "reference"))
    for k, o in zip(kernels, occ):
        k.set_aux_state(o)
    traces = [k.compute_initial_trace(o) for k, o in zip(kernels, occ)]
    feats = [np.array(t.features, dtype=float) for t in traces]
    enth = [float(np.asarray(t.enthalpy).ravel()[0]) for t in traces]
    total, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(32):
            for i, k in enumerate(kernels):
                st = k.single_step(occ[i])
                if st.accepted:
                    feats[i] += st.dfeatures
                    enth[i] += float(np.asarray(st.denthalpy).ravel()[0])
        total += 32 * nwalk
    return total, time.perf_counter() - t0


def cpu_arm(config_id: int, kind: str, target_seconds: float = 12.0):
    """Return (steps/s summed over all cores, cores, kind, sample description)."""
    import multiprocessing as mp
    from oracle import lmc_oracle as O
    cores = os.cpu_count() or 1
    wk = workload(config_id)
    if kind == "reference" and O.load_ref() is None:
        kind = "oracle"
    if kind == "port" and wk.usher not in ("flip", "swap"):
        return None, cores, kind, "the C restatement covers flip / swap steps only"
    if config_id == 5:
        wk.cache_ewald()            # workers map the E x E matrix instead of rebuilding it
    nwalk = 2
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(config_id, kind, c * nwalk, nwalk, target_seconds) for c in range(cores)])
    wall = time.perf_counter() - t0
    rate = sum(r[0] / r[1] for r in res)
    what = {"reference": "smol's compiled Cython evaluators (oracle/_ref) under the restated smol step loop "
                         "(oracle/lmc_oracle.py: single_step + accumulate-on-accept)",
            "oracle": "numpy restatement (oracle/lmc_oracle.py; oracle/_ref not built)",
            "port": "C restatement (oracle/lmc_oracle.c)"}[kind]
    sample = ("%d processes x %d walkers x %.0f s of config %d steps (%d steps in total); %s; pool wall %.1f s"
              % (cores, nwalk, target_seconds, config_id, sum(r[0] for r in res), what, wall))
    return rate, cores, kind, sample


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[1]) for r in self.rows if len(r) > 2 and num(r[1]) is not None]
        mx = [num(r[2]) for r in self.rows if len(r) > 2 and num(r[2]) is not None]
        pw = [num(r[3]) for r in self.rows if len(r) > 3 and num(r[3]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def _profile_summary(config_id):
    """ncu numbers of the kernel this config runs, from the committed capture summaries (profiles/*.json)"""
    for name in ("r02_summary.json", "r01e_summary.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
        c = d.get("config%d" % config_id)
        if c:
            return dict(c, source="profiles/" + name)
        if config_id == 2 and "dram_bytes_per_launch_at_bench_size" in d:
            return {"dram_bytes_per_launch": d["dram_bytes_per_launch_at_bench_size"], "source": "profiles/" + name}
    return {}


# ------------------------------------------------------------------------------------------
def _device_arm(wk, smp, occ_host, W, nsteps, thin, args, world, dist, dev, local_rank, rank, flush):
    """K timed launches of the device-resident path; returns (ms, launches, acceptance, clocks, extra)"""
    import torch
    out = smp.run_device(nsteps, occ_host, thin_by=thin)                 # allocations + initial evaluation
    enth = out["enthalpy"]
    # N > 1: the path's only collective gathers the per-interval observable trace of the WHOLE timed region once,
    # behind the last launch ("once per saved sample (or once per run)"); every launch parks its
    # enthalpy trace in a device buffer (a 256 KB device-to-device copy inside the timed launch interval)
    hist = torch.empty((max(args.steps, 1), *enth.shape), dtype=torch.float64, device=dev) if world > 1 else None
    gathered = torch.empty((world * hist.numel(),), dtype=torch.float64, device=dev) if world > 1 else None

    def step(i):
        smp.run_device(nsteps, None, thin_by=thin, out=out, reuse_state=True)
        if world > 1:
            hist[i % hist.shape[0]].copy_(enth, non_blocking=True)

    for i in range(args.warmup):
        step(i)
    if world > 1:
        dist.all_gather_into_tensor(gathered, hist.view(-1))            # (communicator set up outside the timed region)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = smp.engine.launch_count()
    evs = []
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xff)                      # evict L2 between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(i)
        e1.record()
        evs.append((e0, e1))
    tail_ms = 0.0
    if world > 1:   # the trace gather of the region, timed and added to the total
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        dist.all_gather_into_tensor(gathered, hist.view(-1))
        t1.record()
        torch.cuda.synchronize()
        tail_ms = t0.elapsed_time(t1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = smp.engine.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs) + tail_ms
    acc = float(out["n_accepted"].sum().item()) / (out["n_accepted"].numel() * thin)
    clk = clocks.stop() if rank == 0 else None
    return dev_ms, launches, acc, clk, {"wall_s_timed_region": t_wall, "gather_tail_ms": tail_ms}


def _e2e_arm(wk, smp, occ_dev_state, W, N, nsteps, thin, steps, world, dist):
    """public API with host buffers, back to back (block=False): returns (seconds, h2d bytes, d2h bytes)"""
    import warnings
    import torch
    eng = smp.engine
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # every timed step uploads the same EQUILIBRATED host occupancies (int32, page-locked, as the e2e
        # contract prescribes for the inputs): the walkers' state at the end of the device-resident arm
        occ_host = torch.empty((W, N), dtype=torch.int32, pin_memory=True)
        occ_host.copy_(eng.occupancy_to_int32(occ_dev_state, W, eng.row_stride))
        torch.cuda.synchronize()
        checks = []

        def consume(samples):
            # the step's result, read on the host: mean enthalpy of the last sample + accepted steps
            checks.append((float(samples.get_enthalpies(flat=False)[-1].mean()),
                           int(samples.get_trace_value("n_accepted", flat=False).sum())))
            samples.clear()

        def loop(k):
            prev = None
            for _ in range(k):
                smp.run(nsteps, occ_host, thin_by=thin, block=False)
                cur = smp.detach_samples()
                if prev is not None:
                    consume(prev)             # host read of step i-1 while step i runs
                prev = cur
            consume(prev)
            torch.cuda.synchronize()

        loop(2)                                # warm-up (allocations, page-locked staging)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        loop(steps)
        e2e_s = time.perf_counter() - t0
    h2d = W * N * 4
    d2h = (nsteps // thin) * smp._bytes_per_sample()
    return e2e_s, h2d, d2h, checks[-1]


def run_ours(args):
    # libraries (NCCL's version banner) write to fd 1: keep stdout for the ONE JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wk = workload(args.config)
    t_setup = time.perf_counter()
    ens = wk.product_ensemble()
    N = wk.num_sites
    thin = wk.thin_by
    S_ = args.samples_per_step or wk.samples_per_bench_step
    nsteps = thin * S_

    def one_size(W, steps, with_e2e):
        wbase = rank * W
        seeds = list(range(wbase, wbase + W))
        occ_host = wk.initial_occupancies(W, seed=rank)
        smp = wk.sampler(ens, W, seeds, walker_id_base=wbase)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
        a = argparse.Namespace(steps=steps, warmup=args.warmup)
        dev_ms, launches, acc, clk, extra = _device_arm(wk, smp, occ_host, W, nsteps, thin, a, world, dist, dev,
                                                        local_rank, rank, flush)
        e2e = None
        if with_e2e:
            # enough back-to-back API calls that the fill / drain of the copy pipeline (first upload, last download) is a
            # small part of the region: three times the device arm's steps, 20 .. 100
            e2e_steps = args.e2e_steps or max(20, min(3 * steps, 100))
            e2e_s, h2d, d2h, check = _e2e_arm(wk, smp, smp._occ_dev, W, N, nsteps, thin, e2e_steps, world, dist)
            e2e = (e2e_s, h2d, d2h, e2e_steps, check)
        t = torch.tensor([dev_ms, e2e[0] if e2e else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wl_state = None
        if wk.kernel == "WangLandau":
            st = smp.wang_landau_state
            wl_state = {"bins": int(len(st["levels"])), "min_mod_factor": float(st["mod_factor"].min()),
                        "max_mod_factor": float(st["mod_factor"].max()),
                        "visited_bins_mean": float((st["entropy"] > 0).sum(1).mean())}
        cache = bool(getattr(smp, "ewald_cache_in_use", False))
        del smp, flush
        torch.cuda.empty_cache()
        return dict(W=W, dev_ms=float(t[0]), e2e_s=float(t[1]), launches=launches, acc=acc, clk=clk, extra=extra,
                    e2e=e2e, steps=steps, wl=wl_state, cache=cache)

    W = args.walkers or wk.walkers_per_gpu
    setup_s = time.perf_counter() - t_setup
    main_res = one_size(W, args.steps, True)
    # the rate is a property of the acceptance ratio (rejected steps cost less than accepted ones): the same
    # workload at other temperatures, device-resident, equilibrated for a few launches each (Metropolis configs)
    curve = None
    if args.curve and wk.kernel == "Metropolis" and world == 1:
        curve = []
        for mult in args.curve_temperatures:
            T = wk.temperature * mult
            smp = wk.sampler(ens, W, list(range(W)))
            smp._temperature[:] = T
            out = smp.run_device(nsteps, wk.initial_occupancies(W, seed=7), thin_by=thin)
            for _ in range(4):
                smp.run_device(nsteps, None, thin_by=thin, out=out, reuse_state=True)
            ms = 0.0
            for _ in range(3):
                smp.run_device(nsteps, None, thin_by=thin, out=out, reuse_state=True)
                ms += smp.last_kernel_ms
            acc = float(out["n_accepted"].sum().item()) / (out["n_accepted"].numel() * thin)
            curve.append({"temperature_K": T, "acceptance_ratio": acc, "steps_per_s": 3 * W * nsteps / (ms * 1e-3)})
            del smp, out
            torch.cuda.empty_cache()
    # a sustained region: the same launches back to back for >= args.sustain_seconds (no L2 flush in between)
    sustained = None
    if args.sustain_seconds > 0 and rank == 0 and world == 1:
        smp = wk.sampler(ens, W, list(range(W)))
        out = smp.run_device(nsteps, wk.initial_occupancies(W, seed=0), thin_by=thin)
        for _ in range(3):
            smp.run_device(nsteps, None, thin_by=thin, out=out, reuse_state=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_l, t0 = 0, time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < args.sustain_seconds:
            for _ in range(8):
                smp.run_device(nsteps, None, thin_by=thin, out=out, reuse_state=True)
            n_l += 8
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        sms = e0.elapsed_time(e1)
        sustained = {"seconds": sms / 1e3, "launches": n_l, "value": n_l * W * nsteps / (sms * 1e-3), "unit": "steps/s"}
        del smp, out
        torch.cuda.empty_cache()
    strong = None
    if args.config == 5 and not args.no_strong:
        # north star: 32768 walkers over the GPUs of the box (strong scaling), beside the weak line above
        total_w = args.strong_walkers
        if total_w % world == 0:
            strong = one_size(total_w // world, max(2, args.steps // 4) if world == 1 else args.steps, False)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    steps_per_launch = W * nsteps
    dev_ms = main_res["dev_ms"]
    value = world * steps_per_launch * args.steps / (dev_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_gpu_rate = steps_per_launch * args.steps / (dev_ms * 1e-3)
    built, built_note = wk.built_bytes(main_res["acc"], main_res["cache"])
    primary = wk.algorithmic_bytes if wk.roofline_uses_survey_bytes else built
    achieved = per_gpu_rate * primary / 1e9
    prof = _profile_summary(args.config)
    cpu_rate, cores, kind, sample = (None, 0, "skipped", "")
    port_rate = None
    if not args.no_cpu and world == 1:      # the CPU legs run beside the one-GPU line only
        cpu_rate, cores, kind, sample = cpu_arm(args.config, "reference", target_seconds=args.cpu_seconds)
        port_rate = cpu_arm(args.config, "port", target_seconds=min(6.0, args.cpu_seconds))[0]
    e2e_s, h2d, d2h, e2e_steps, check = main_res["e2e"]
    out = {
        "metric": METRIC, "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": wk.name % W, "baseline_config": args.config,
                   "attempted_steps_per_bench_step": steps_per_launch,
                   "sampling_intervals_per_bench_step": S_, "thin_by": thin, "l2_flush_between_iterations": True,
                   "parallelism": "walkers sharded, %d/GPU" % W, "acceptance_ratio": main_res["acc"],
                   "model_setup_s": setup_s, "ewald_potential_cache": main_res["cache"]},
        "e2e": {"value": world * steps_per_launch * e2e_steps / main_res["e2e_s"], "unit": "steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "smol_b200.Sampler.run(nsteps, initial_occupancies=<page-locked host int32>, thin_by, "
                       "block=False) back to back; each step's samples (detach_samples) are read on the host "
                       "(last-sample mean enthalpy, accepted steps) while the next step runs; occupancies arrive "
                       "on the host as int8 codes and are widened to the reference's int32 on access "
                       "(outside the timed region)",
                "last_result": {"mean_enthalpy": check[0], "accepted_steps": check[1]}},
        "gpu_launches": int(main_res["launches"]),
        "clocks": main_res["clk"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (prof["dram_bytes_per_attempted_step"] * steps_per_launch
                                 if "dram_bytes_per_attempted_step" in prof else prof.get("dram_bytes_per_launch")),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
                     "bytes_per_step": primary,
                     "bytes_model": ("SURVEY 8(d) algorithmic bytes of the reference algorithm"
                                     if wk.roofline_uses_survey_bytes else "bytes the built algorithm moves: " + built_note),
                     "survey_8d_bytes_per_step": wk.algorithmic_bytes,
                     "built_bytes_per_step": built, "built_bytes_model": built_note,
                     "frac_built": per_gpu_rate * built / 1e9 / peak,
                     "limiter": prof.get("limiter"), "profile": prof.get("source")},
        "cpu_baseline": {"value": cpu_rate, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample},
        "cpu_port": {"value": port_rate, "unit": "steps/s", "cores": cores, "kind": "port",
                     "sample": "C restatement oracle/lmc_oracle.c, one process per core"},
    }
    out.update(main_res["extra"])
    if args.config == 2:
        # the reference's OWN Python loop cannot run on the GPU box (it needs the reference tree): its rate measured in
        # the build container is quoted from the committed file for context (scripts/reference_python_rate.py)
        try:
            rp = json.load(open(os.path.join(ROOT, "profiles", "r02_reference_python.json")))
            out["cpu_baseline"]["note"] = ("kind 'reference' = smol's compiled evaluators under the RESTATED step loop; smol's own "
                                           "unmodified Python loop over the same evaluators: %.3g steps/s per process (%s, "
                                           "profiles/r02_reference_python.json)" % (rp["single_process_steps_per_s"], rp["where"]))
        except Exception:
            pass
    if curve:
        out["config"]["acceptance_curve"] = curve
    if sustained:
        out["sustained"] = sustained
    if main_res["wl"]:
        out["config"]["wang_landau"] = main_res["wl"]
    if strong:
        sv = world * strong["W"] * nsteps * strong["steps"] / (strong["dev_ms"] * 1e-3)
        out["strong"] = {"walkers_total": world * strong["W"], "walkers_per_gpu": strong["W"], "value": sv,
                         "unit": "steps/s", "steps": strong["steps"], "ms_per_step": strong["dev_ms"] / strong["steps"],
                         "scaling": "strong", "acceptance_ratio": strong["acc"]}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wk = workload(args.config)
    # every bench step is a bounded sample (--ref-seconds of CPU work on every core): the requested step
    # count is kept up to a budget of ~3 minutes for the whole arm and the clamp is reported
    budget = 180.0
    steps = max(1, min(args.steps, int(budget / max(args.ref_seconds, 0.5)) - 1))
    warmup = min(args.warmup, 1)
    rates = []
    cores = kind = sample = None
    t0 = time.perf_counter()
    for i in range(warmup + steps):
        rate, cores, kind, sample = cpu_arm(args.config, "reference", target_seconds=args.ref_seconds)
        if i >= warmup:
            rates.append(rate)
        per_step = (time.perf_counter() - t0) / (i + 1)      # sample + worker start-up (tables, Ewald matrix)
        if rates and time.perf_counter() - t0 + per_step > budget:
            break
    steps = len(rates)
    value = float(np.mean(rates))
    W = args.walkers or wk.walkers_per_gpu
    out = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps,
        "warmup": warmup,
        "ms_per_step": 1e3 * (time.perf_counter() - t0) / max(1, steps + warmup),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/int32",
        "data": "synthetic",
        "config": {"workload": wk.name % W, "baseline_config": args.config,
                   "requested_steps": args.steps, "requested_warmup": args.warmup,
                   "note": "reference CPU arm: each bench step is a bounded sample of the workload (see "
                           "cpu_baseline.sample); steps / warmup are clamped to keep the arm within minutes; smol "
                           "itself runs walkers serially in one process, here one process per host core"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample,
                         "per_core": value / max(cores or 1, 1)},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configuration (default 2: the one the metric is quoted on)")
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: the config's)")
    ap.add_argument("--samples-per-step", type=int, default=0,
                    help="sampling intervals per bench step (default: the config's; raise for a sustained run)")
    ap.add_argument("--strong-walkers", type=int, default=32768, help="config 5: total walkers of the strong-scaling line")
    ap.add_argument("--no-strong", action="store_true", help="config 5: skip the strong-scaling line")
    ap.add_argument("--e2e-steps", type=int, default=0, help="API calls of the end-to-end arm (default: 3 x steps, 20..100)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-curve", dest="curve", action="store_false",
                    help="skip the acceptance curve (the workload at other temperatures; Metropolis configs, one GPU)")
    ap.add_argument("--curve-temperatures", type=float, nargs="*", default=[2.0, 4.0, 8.0, 16.0],
                    help="temperature multipliers of the acceptance curve")
    ap.add_argument("--sustain-seconds", type=float, default=1.5,
                    help="length of the additional back-to-back region (0 = off)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=4.0)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {2: 30, 3: 20, 4: 10, 5: 8}[args.config]
    if args.warmup is None:
        args.warmup = 5 if args.config == 2 else 3
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    if int(os.environ.get("WORLD_SIZE", "1")) != args.gpus and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
