"""Generate golden vectors from the REFERENCE's own compiled Cython evaluators (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python oracle/build_ref.py && python tests/golden/make_golden.py
Inputs are seeded; outputs are what smol's ClusterSpaceEvaluator / delta_ewald_single_flip
return for them (smol/utils/cluster/evaluator.pyx, ewald.pyx).  The vectors pin the oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_golden.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import lmc_oracle as O  # noqa: E402
from smol_b200 import lattice as L  # noqa: E402
from tests import models as M  # noqa: E402

CASES = {
    "fcc2": (lambda: M.fcc_subspace(), 2),
    "fcc3": (lambda: M.fcc_subspace(), 3),
    "rs2": (lambda: M.rocksalt_subspace(anions=("O2-", "F-")), 2),
}


def main():
    ref = O.load_ref()
    assert ref is not None, "build oracle/_ref first"
    out = {}
    for name, (mk, n) in CASES.items():
        sub = mk()
        scm = np.eye(3, dtype=int) * n
        rng = np.random.default_rng(1234)
        coefs = rng.normal(0, 0.05, sub.num_corr_functions)
        it = L.cluster_interaction_tensors(sub, coefs)
        ce = O.ClusterExpansionProcessor(sub, scm, coefs, use_ref=True)
        cd = O.ClusterDecompositionProcessor(sub, scm, it, use_ref=True)
        spaces = sub.allowed_species(scm)
        active = [i for i, s in enumerate(spaces) if len(s) > 1]
        W, k = 6, 3
        occ = M.random_occupancies(sub, scm, W, seed=77)
        sites = rng.choice(active, size=(W, k))
        codes = np.zeros((W, k), dtype=np.int32)
        for w in range(W):
            cur = occ[w].copy()
            for j in range(k):
                s = sites[w, j]
                codes[w, j] = rng.choice([c for c in range(len(spaces[s])) if c != cur[s]])
                cur[s] = codes[w, j]
        out[f"{name}_coefs"] = coefs
        out[f"{name}_occ"] = occ
        out[f"{name}_sites"] = sites.astype(np.int32)
        out[f"{name}_codes"] = codes
        out[f"{name}_full_corr"] = np.array([ce.compute_feature_vector(o) for o in occ])
        out[f"{name}_full_inter"] = np.array([cd.compute_feature_vector(o) for o in occ])
        out[f"{name}_delta_corr"] = np.array(
            [ce.compute_feature_vector_change(occ[w], list(zip(sites[w], codes[w]))) for w in range(W)])
        out[f"{name}_delta_inter"] = np.array(
            [cd.compute_feature_vector_change(occ[w], list(zip(sites[w], codes[w]))) for w in range(W)])
        if name == "rs2":
            ewm, ewi = L.ewald_matrix(sub, scm)
            ew = O.EwaldProcessor(ewm, ewi, 1.0, use_ref=True)
            out["rs2_ewald_matrix"] = ewm
            out["rs2_ewald_inds"] = ewi
            out["rs2_delta_ewald"] = np.array(
                [ew.compute_feature_vector_change(occ[w], list(zip(sites[w], codes[w]))) for w in range(W)])
            out["rs2_full_ewald"] = np.array([ew.compute_feature_vector(o) for o in occ])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
