"""Where the end-to-end loop of bench.py loses against the device-resident rate: the same loop with parts removed.
python scripts/e2e_variants.py [config]"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads as WK

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wk = WK.get(cid); ens = wk.product_ensemble(); N = wk.num_sites; W = wk.walkers_per_gpu
nsteps = wk.thin_by * wk.samples_per_bench_step
out = {}
for name, kw, src in (("full", {}, "pinned"), ("no_occupancy_trace", dict(record_occupancy=False), "pinned"), ("device_input", {}, "cuda")):
    smp = wk.sampler(ens, W, list(range(W)), **kw)
    dev = smp.run_device(nsteps * 4, wk.initial_occupancies(W), thin_by=wk.thin_by)      # equilibrate
    occ32 = smp.engine.occupancy_to_int32(smp._occ_dev, W, smp.engine.row_stride)
    if src == "pinned":
        occ = torch.empty((W, N), dtype=torch.int32, pin_memory=True); occ.copy_(occ32)
    else:
        occ = occ32.clone()
    torch.cuda.synchronize()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        def loop(k):
            prev = None
            for _ in range(k):
                smp.run(nsteps, occ, thin_by=wk.thin_by, block=False)
                cur = smp.detach_samples()
                if prev is not None:
                    _ = float(prev.get_enthalpies(flat=False)[-1].mean()); prev.clear()
                prev = cur
            _ = float(prev.get_enthalpies(flat=False)[-1].mean()); prev.clear()
            torch.cuda.synchronize()
        loop(3)
        t0 = time.perf_counter(); loop(40); dt = (time.perf_counter() - t0) / 40
    out[name] = dict(ms_per_call=round(1e3 * dt, 3), kernel_ms=round(smp.last_kernel_ms, 3), steps_per_s="%.3e" % (W * nsteps / dt))
    del smp
print(json.dumps(out))
