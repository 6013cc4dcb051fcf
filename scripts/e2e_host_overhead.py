"""Host cost of one Sampler.run(block=False) + detach_samples + read-back: the e2e loop of bench.py on a workload whose
kernel is tiny (few walkers), so that the per-call time is the Python / driver overhead.  python scripts/e2e_host_overhead.py [config]"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads as WK

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wk = WK.get(cid); ens = wk.product_ensemble(); N = wk.num_sites
out = {}
for W in (32, wk.walkers_per_gpu):
    smp = wk.sampler(ens, W, list(range(W)))
    occ = torch.empty((W, N), dtype=torch.int32, pin_memory=True); occ.copy_(torch.from_numpy(wk.initial_occupancies(W)))
    nsteps = wk.thin_by * wk.samples_per_bench_step
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        def loop(k):
            prev = None; t_run = 0.0; t_cons = 0.0
            for _ in range(k):
                t0 = time.perf_counter()
                smp.run(nsteps, occ, thin_by=wk.thin_by, block=False)
                cur = smp.detach_samples()
                t1 = time.perf_counter()
                if prev is not None:
                    _ = float(prev.get_enthalpies(flat=False)[-1].mean()); prev.clear()
                t2 = time.perf_counter()
                t_run += t1 - t0; t_cons += t2 - t1
                prev = cur
            torch.cuda.synchronize()
            return t_run / k, t_cons / k
        loop(3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r, c = loop(40)
        tot = (time.perf_counter() - t0) / 40
    out["W%d" % W] = dict(ms_per_call=1e3 * tot, run_enqueue_ms=1e3 * r, consume_ms=1e3 * c, kernel_ms=smp.last_kernel_ms)
print(json.dumps(out))
