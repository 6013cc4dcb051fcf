"""One short resident run of a BASELINE config for ncu: python scripts/prof_cfg.py <config> <samples per launch> [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import workloads as WK
cid, S_ = int(sys.argv[1]), int(sys.argv[2])
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wk = WK.get(cid); ens = wk.product_ensemble(); W = int(os.environ.get("PROF_WALKERS", wk.walkers_per_gpu))
smp = wk.sampler(ens, W, list(range(W)))
out = smp.run_device(wk.thin_by * S_, wk.initial_occupancies(W), thin_by=wk.thin_by)
for _ in range(nl - 1):
    smp.run_device(wk.thin_by * S_, None, thin_by=wk.thin_by, out=out, reuse_state=True)
torch.cuda.synchronize()
print("ms", smp.last_kernel_ms, "steps", W * wk.thin_by * S_)
