"""Quick throughput probe of lmc_run on config 2 (binary FCC 8x8x8 canonical swap)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smol_b200 as S
from smol_b200 import lattice as L
from tests import models as M

def main():
    W = int(os.environ.get("W", 4096)); n = int(os.environ.get("NCELL", 8))
    sweeps = int(os.environ.get("SWEEPS", 20))
    sub = M.fcc_subspace(); scm = np.eye(3, dtype=int) * n
    coefs = M.fcc_coefs(sub)
    it = L.cluster_interaction_tensors(sub, coefs)
    kind = os.environ.get("KIND", "decomposition")
    proc = S.ClusterDecompositionProcessor(sub, scm, it) if kind == "decomposition" else S.ClusterExpansionProcessor(sub, scm, coefs)
    ens = S.Ensemble(proc)
    occ0 = M.random_occupancies(sub, scm, W, seed=0, balanced=True)
    N = occ0.shape[1]
    T = float(os.environ.get("TEMP", 1000.0))
    for G in [int(x) for x in os.environ.get("GS", "4,8,16,32").split(",")]:
        for bt in [int(x) for x in os.environ.get("BTS", "128").split(",")]:
            # GS=-1: speculative-batch kernel, GS=0: auto
            smp = S.Sampler.from_ensemble(ens, T, step_type="swap", nwalkers=W, seeds=list(range(W)),
                                          group_size=max(G, 0), block_threads=bt, spec_mode=2 if G < 0 else (1 if G > 0 else 0))
            smp.run(N * 2, occ0, thin_by=N)
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            smp.clear_samples()
            t0.record(); smp.run(N * sweeps, thin_by=N); t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
            acc = smp.samples.step_efficiency()
            kms = smp.last_kernel_ms
            print(f"G={G:2d} threads={bt:4d} W={W} N={N}: kernel {W*N*sweeps/kms*1e3:.3e} steps/s ({kms:.1f} ms) | run() {W*N*sweeps/ms*1e3:.3e} ({ms:.1f} ms) acc={acc:.3f}", flush=True)

main()
