"""Where one Sampler.run of a BASELINE config spends its time: CUDA-event timings of the phases
(upload + cast, initial full evaluation, Ewald potential cache, lmc_run launches) and the wall time
of blocking / non-blocking runs.   python scripts/e2e_breakdown.py <config> [samples]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads as WK

cid = int(sys.argv[1]); wk = WK.get(cid)
S_ = int(sys.argv[2]) if len(sys.argv) > 2 else wk.samples_per_bench_step
ens = wk.product_ensemble(); W = wk.walkers_per_gpu; N = wk.num_sites
smp = wk.sampler(ens, W, list(range(W)))
occ = wk.initial_occupancies(W)
pin = torch.empty((W, N), dtype=torch.int32, pin_memory=True); pin.copy_(torch.from_numpy(occ))
eng = smp.engine
nsteps = wk.thin_by * S_
smp.run(nsteps, pin, thin_by=wk.thin_by); smp.clear_samples()


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


res = {}
for rep in range(3):
    torch.cuda.synchronize()
    e0 = ev(); od = eng.upload_occupancy(pin); e1 = ev()
    ff = eng.full_features(od); e2 = ev()
    fld = eng.ewald_field(od) if eng.model_info()[0] else None; e3 = ev()
    torch.cuda.synchronize()
    res = dict(upload_cast_ms=e0.elapsed_time(e1), full_features_ms=e1.elapsed_time(e2), ewald_field_ms=e2.elapsed_time(e3))
for mode in (True, False):
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        smp.run(nsteps, pin, thin_by=wk.thin_by, block=mode)
        s = smp.detach_samples(); _ = s.get_enthalpies(flat=False)[-1].mean(); s.clear()
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    res["run_%s_ms" % ("blocking" if mode else "nonblocking")] = 1e3 * float(np.median(ts))
res["kernel_ms"] = smp.last_kernel_ms
res["d2h_MB"] = S_ * smp._bytes_per_sample() / 1e6; res["h2d_MB"] = W * N * 4 / 1e6
res["config"] = cid; res["samples"] = S_
print(json.dumps(res))
