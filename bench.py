#!/usr/bin/env python
"""Headline benchmark: attempted MC steps/s of the lattice-MC hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): binary FCC 8x8x8 supercell
(N = 512 sites), cluster set S_fcc (point, pairs 1NN-4NN, NN triangle, NN tetrahedron),
cluster-decomposition processor, canonical Metropolis with Swap proposals at T = 1000 K,
4096 walkers per GPU, one sample per sweep (thin_by = 512).

A bench "step" = one ``lmc_run`` launch advancing every walker by SWEEPS_PER_STEP sweeps
(= W * 512 * SWEEPS_PER_STEP attempted MC steps), writing the per-sweep traces to HBM.

  value  device-resident throughput: CUDA events around the launches, state and tables in HBM
  e2e    the same metric through the public API (Sampler.run) with HOST buffers: per step the
         initial occupancies go host->device from pinned memory and all traces come back to host
  roofline   algorithmic bytes (430 B per attempted swap step) / launch time vs the
             measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own compiled Cython evaluators (oracle/_ref) driven by the
             restated smol step loop, all host cores, bounded sample
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

N_CELL = 8
WALKERS_PER_GPU = 4096
TEMPERATURE = 1000.0
SWEEPS_PER_STEP = 8
ALGO_BYTES_PER_STEP = 430.0   # 426 int8 gathers + 2 writes + trace/thin_by
WORKLOAD = ("binary FCC 8x8x8 (512 sites), S_fcc clusters, canonical Metropolis swap, T=1000K, "
            "%d walkers/GPU, thin_by=512" % WALKERS_PER_GPU)


def build_model():
    from smol_b200 import lattice as L
    from tests import models as M
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * N_CELL
    coefs = M.fcc_coefs(sub)
    it = L.cluster_interaction_tensors(sub, coefs)
    return sub, scm, coefs, it


# ------------------------------------------------------------------------------------------
# CPU arm: reference Cython evaluators + restated step loop, one process per host core
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, walker0, nwalk, nsteps = args
    from oracle import lmc_oracle as O
    from tests import models as M
    sub, scm, coefs, it = build_model()
    subl = [O.Sublattice(("A", "B"), np.arange(N_CELL ** 3))]
    if kind == "reference":
        proc = O.ClusterDecompositionProcessor(sub, scm, it, use_ref=True)
        occ0 = M.random_occupancies(sub, scm, nwalk, seed=walker0, balanced=True)
        kernels = [O.Metropolis(O.Ensemble(proc, subl), O.Swap(subl), TEMPERATURE, seed=walker0 + w,
                                walker=walker0 + w) for w in range(nwalk)]
        t0 = time.perf_counter()
        O.run_sampler(kernels, occ0, nsteps, thin_by=max(1, nsteps))
    else:
        from oracle import c_oracle as CO
        proc = O.ClusterDecompositionProcessor(sub, scm, it)
        co = CO.COracle(O.Ensemble(proc, subl))
        occ0 = M.random_occupancies(sub, scm, nwalk, seed=walker0, balanced=True)
        t0 = time.perf_counter()
        co.run(occ0, nsteps, nsteps, np.arange(walker0, walker0 + nwalk), usher="swap",
               temperature=TEMPERATURE, nthreads=1, record=False)
    return nwalk * nsteps, time.perf_counter() - t0


def cpu_arm(kind: str, target_seconds: float = 12.0):
    """Return (steps/s summed over all cores, cores, kind, sample description)."""
    import multiprocessing as mp
    from oracle import lmc_oracle as O
    cores = os.cpu_count() or 1
    if kind == "reference" and O.load_ref() is None:
        kind = "port"
    rate_guess = 1.2e4 if kind == "reference" else 4e5
    nwalk = 2
    nsteps = max(256, int(rate_guess * target_seconds / nwalk))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, 1000 + c * nwalk, nwalk, nsteps) for c in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    what = ("smol's compiled Cython evaluators (oracle/_ref) + restated smol step loop"
            if kind == "reference" else "C restatement (oracle/lmc_oracle.c)")
    sample = ("%d processes x %d walkers x %d swap steps of the same model; %s; slowest worker %.1f s "
              "(pool wall %.1f s)" % (cores, nwalk, nsteps, what, busy, wall))
    return total / busy, cores, kind, sample


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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    # libraries (NCCL's version banner) write to fd 1: keep stdout for the ONE JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import smol_b200 as S
    from smol_b200 import _capi as capi
    from tests import models as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sub, scm, coefs, it = build_model()
    N = N_CELL ** 3
    W = WALKERS_PER_GPU
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it))
    wbase = rank * W                          # weak scaling: every GPU owns 4096 walkers
    seeds = list(range(wbase, wbase + W))
    occ_host = M.random_occupancies(sub, scm, W, seed=rank, balanced=True)
    smp = S.Sampler.from_ensemble(ens, TEMPERATURE, step_type="swap", nwalkers=W, seeds=seeds,
                                  walker_id_base=wbase)
    eng = smp.engine
    steps_per_launch = W * N * SWEEPS_PER_STEP

    # ---- device-resident arm: state + traces live in HBM, one launch per bench step ------------
    occ_dev = eng.upload_occupancy(occ_host)
    feat, enth = eng.full_features(occ_dev)
    S_, F = SWEEPS_PER_STEP, eng.F
    tr_occ = torch.empty((S_, W, N), dtype=torch.int8, device=dev)
    tr_feat = torch.empty((S_, W, F), dtype=torch.float64, device=dev)
    tr_enth = torch.empty((S_, W), dtype=torch.float64, device=dev)
    tr_acc = torch.empty((S_, W), dtype=torch.uint8, device=dev)
    tr_nacc = torch.empty((S_, W), dtype=torch.int32, device=dev)
    seeds_t = torch.from_numpy(np.array(seeds, dtype=np.uint64).view(np.int64)).to(dev)
    beta = torch.full((W,), 1.0 / (smp.kB * TEMPERATURE), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * S_ * W,), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def launch(step_index):
        cfg = capi.LmcRunConfig()
        cfg.num_walkers, cfg.walker_id_base = W, wbase
        cfg.usher, cfg.kernel = capi.LMC_USHER_SWAP, capi.LMC_KERNEL_METROPOLIS
        cfg.num_samples, cfg.thin_by = S_, N
        cfg.step_begin = step_index * S_ * N
        cfg.seeds_dev, cfg.beta_dev = seeds_t.data_ptr(), beta.data_ptr()
        cfg.occ_dev, cfg.features_dev, cfg.enthalpy_dev = occ_dev.data_ptr(), feat.data_ptr(), enth.data_ptr()
        cfg.trace_occ_dev, cfg.trace_features_dev = tr_occ.data_ptr(), tr_feat.data_ptr()
        cfg.trace_enthalpy_dev, cfg.trace_accepted_dev = tr_enth.data_ptr(), tr_acc.data_ptr()
        cfg.trace_naccepted_dev = tr_nacc.data_ptr()
        eng.run(cfg)
        if world > 1:   # the only collective of the path: gather the per-sweep observable trace
            dist.all_gather_into_tensor(gathered, tr_enth.view(-1))

    for i in range(args.warmup):
        launch(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count()
    evs = []
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                      # evict L2 between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(args.warmup + i)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    acc_frac = float(tr_nacc.sum().item()) / (S_ * W * N)

    # ---- end-to-end arm: public API, host buffers, H2D + D2H inside the timed region ------------
    import warnings
    e2e_steps = max(3, min(args.steps, 20))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # every timed step uploads the same EQUILIBRATED host occupancies (int32): the walkers' state at
        # the end of the device-resident arm, so that both arms are timed in the same regime
        # (int32, in page-locked host memory, as the e2e contract prescribes for the inputs)
        occ_host = torch.empty((W, N), dtype=torch.int32, pin_memory=True)
        occ_host.copy_(eng.occupancy_to_int32(occ_dev, W, eng.row_stride))
        torch.cuda.synchronize()
        for _ in range(2):
            smp.run(N * S_, occ_host, thin_by=N)      # warm-up (allocations, pinned staging)
            smp.clear_samples()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            smp.run(N * S_, occ_host, thin_by=N)   # uploads occ_host, returns all traces to host
            smp.clear_samples()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    h2d = W * N * 4
    d2h = S_ * W * (N + 8 * F + 8 + 1 + 4)

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    total_steps = world * steps_per_launch * args.steps
    value = total_steps / (dev_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_gpu_rate = steps_per_launch * args.steps / (dev_ms * 1e-3)
    achieved = per_gpu_rate * ALGO_BYTES_PER_STEP / 1e9
    traffic = None
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel this bench runs, from the
        # committed ncu --set full capture (profiles/r01e_lmc_spec_cfg2.md)
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01e_summary.json")))
        traffic = prof.get("dram_bytes_per_launch_at_bench_size")
    except Exception:
        pass
    cpu_rate, cores, kind, sample = (None, 0, "skipped", "")
    port_rate = None
    if not args.no_cpu:
        cpu_rate, cores, kind, sample = cpu_arm("reference")
        port_rate = cpu_arm("port", target_seconds=6.0)[0]
    out = {
        "metric": "attempted MC steps/sec (whole job)", "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (energies/features), u8 (occupancy), u32 (rng)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "attempted_steps_per_bench_step": steps_per_launch,
                   "sweeps_per_bench_step": SWEEPS_PER_STEP, "l2_flush_between_iterations": True,
                   "parallelism": "walkers sharded, %d/GPU" % W, "acceptance_ratio": acc_frac,
                   "group_size": os.environ.get("LMC_GROUP_SIZE", "auto")},
        "e2e": {"value": world * steps_per_launch * e2e_steps / e2e_s, "unit": "steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "smol_b200.Sampler.run(nsteps, initial_occupancies=<page-locked host int32>, thin_by=512)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650",
                     "algorithmic_bytes_per_attempted_step": ALGO_BYTES_PER_STEP,
                     "note": "sparse integer gather-reduce on an L2/SMEM-resident working set: DRAM traffic is far "
                             "below the algorithmic bytes; the kernel is bound by the L1/shared-memory data pipe "
                             "(l1tex__data_pipe_lsu_wavefronts 88 % of peak, profiles/r01e_lmc_spec_cfg2.md)"},
        "cpu_baseline": {"value": cpu_rate, "unit": "steps/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "cpu_port": {"value": port_rate, "unit": "steps/s", "cores": cores, "kind": "port",
                     "sample": "C restatement oracle/lmc_oracle.c, one process per core"},
        "wall_s_timed_region": t_wall,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates = []
    cores = kind = sample = None
    t0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        rate, cores, kind, sample = cpu_arm("reference", target_seconds=args.ref_seconds)
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    out = {
        "impl": "reference", "metric": "attempted MC steps/sec (whole job)", "value": value,
        "unit": "steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * (time.perf_counter() - t0) / max(1, args.steps + args.warmup),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/int32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU arm: each bench step is a bounded sample "
                   "of the workload (see cpu_baseline.sample); smol itself runs walkers serially in one "
                   "process, here one process per host core"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--ref-seconds", type=float, default=4.0)
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 5)
        args.warmup = min(args.warmup, 1)
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    if int(os.environ.get("WORLD_SIZE", "1")) != args.gpus and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        if args.no_cpu:
            cmd.append("--no-cpu")
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
