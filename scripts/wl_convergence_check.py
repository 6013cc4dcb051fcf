"""Offline half of scripts/wl_convergence.py: the three walkers saved by the GPU run against the C restatement of the
reference loop (oracle/lmc_oracle.c), same seeds / window / number of steps: histogram, entropy, occurrences and
modification factor bit for bit.   python scripts/wl_convergence_check.py gpurun_out/r02_wl_state.npz"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import c_oracle as CO
from oracle import lmc_oracle as O
from tests import workloads as WK

d = np.load(sys.argv[1])
wk = WK.get(4)
lo, hi = wk.window()
co = CO.COracle(O.Ensemble(wk.oracle_processor(), wk.oracle_sublattices()))
steps = int(d["steps"])
res = []
t0 = time.time()
for i, w in enumerate(d["pick"]):
    _, state = co.run(d["occ0"][i:i + 1], steps, steps, d["seeds"][i:i + 1], usher="flip", walker_base=int(w), record=False,
                      wl=dict(min=lo, max=hi, bin=wk.bin_size, flatness=wk.flatness, check=wk.check_period))
    res.append({"walker": int(w), "histogram": bool(np.array_equal(state["histogram"][0], d["histogram"][i])),
                "entropy": bool(np.array_equal(state["entropy"][0], d["entropy"][i])),
                "occurrences": bool(np.array_equal(state["occurrences"][0], d["occurrences"][i])),
                "mod_factor": bool(state["mod_factor"][0] == d["mod_factor"][i]), "ln_f": float(d["mod_factor"][i])})
print(json.dumps({"steps_per_walker": steps, "oracle_s": time.time() - t0, "walkers": res,
                  "all_bit_exact": all(all(v for k, v in r.items() if k not in ("walker", "ln_f")) for r in res)}, indent=1))
