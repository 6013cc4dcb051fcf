"""BASELINE config 4 run to flat-histogram convergence: every walker's modification factor ln f below 1e-6
(wanglandau.py:253-264: ln f is halved whenever the histogram of the visited bins is flat to 80 %).

Advances all 1024 walkers in chunks, records per walker the first chunk boundary at which ln f <= 1e-1 ... 1e-6,
and checks three walkers against the C restatement of the reference loop (oracle/lmc_oracle.c) run for the same
number of steps: histogram, entropy and modification factor bit for bit.
   python scripts/wl_convergence.py [steps per walker] [chunk] > profiles/r02_cfg4_convergence.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests import workloads as WK

total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 60_000_000
chunk = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 20
wk = WK.get(4)
W = wk.walkers_per_gpu
ens = wk.product_ensemble()
seeds = np.arange(W) * 31 + 5
occ0 = wk.initial_occupancies(W)
smp = wk.sampler(ens, W, seeds)
targets = [10.0 ** -k for k in range(1, 7)]
reached = np.full((len(targets), W), -1, dtype=np.int64)
done, kernel_ms, t0 = 0, 0.0, time.time()
first = True
while done < total:
    n = min(chunk, total - done)
    smp.run(n, occ0 if first else None, thin_by=n)
    first = False
    kernel_ms += smp.last_kernel_ms
    smp.clear_samples()
    done += n
    mf = smp._wl_state["mod_factor"].cpu().numpy()
    for i, t in enumerate(targets):
        hit = (mf <= t) & (reached[i] < 0)
        reached[i, hit] = done
    if (reached[-1] >= 0).all():
        break
st = smp.wang_landau_state
out = {"config": wk.name % W, "steps_per_walker": int(done), "chunk": int(chunk), "kernel_s": kernel_ms / 1e3,
       "wall_s": time.time() - t0, "steps_per_s": W * done / (kernel_ms / 1e3),
       "bins": int(len(st["levels"])), "visited_bins_min": int((st["entropy"] > 0).sum(1).min()),
       "mod_factor_min": float(st["mod_factor"].min()), "mod_factor_max": float(st["mod_factor"].max()),
       "halvings_min": int(np.round(-np.log2(st["mod_factor"].max()))), "halvings_max": int(np.round(-np.log2(st["mod_factor"].min()))),
       "steps_to_ln_f": {("%.0e" % t): {"walkers_reached": int((reached[i] >= 0).sum()),
                                          "min": int(reached[i][reached[i] >= 0].min()) if (reached[i] >= 0).any() else None,
                                          "median": float(np.median(reached[i][reached[i] >= 0])) if (reached[i] >= 0).any() else None,
                                          "max": int(reached[i].max()) if (reached[i] >= 0).all() else None}
                         for i, t in enumerate(targets)}}
# three walkers' final state for the offline check against the C restatement (scripts/wl_convergence_check.py)
pick = [0, 333, 1023]
np.savez(os.environ.get("WL_STATE_OUT", "gpurun_out/r02_wl_state.npz"), pick=np.array(pick), steps=np.int64(done),
         seeds=seeds[pick], occ0=occ0[pick], histogram=st["histogram"][pick], entropy=st["entropy"][pick],
         occurrences=st["occurrences"][pick], mod_factor=st["mod_factor"][pick])
print(json.dumps(out, indent=1))
