import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smol_b200 as S
from smol_b200 import lattice as L
from oracle import lmc_oracle as O
from tests import models as M

def run(sub, scm, W=3, nst=300, G=32, label=""):
    rng = np.random.default_rng(3)
    coefs = rng.normal(0, 0.03, sub.num_corr_functions)
    it = L.cluster_interaction_tensors(sub, coefs)
    gp = S.ClusterDecompositionProcessor(sub, scm, it)
    op = O.ClusterDecompositionProcessor(sub, scm, it)
    ens = S.Ensemble(gp)
    occ0 = M.random_occupancies(sub, scm, W, seed=5)
    smp = S.Sampler.from_ensemble(ens, 2000.0, step_type="swap", nwalkers=W, seeds=list(range(W)), group_size=G)
    smp.run(nst, occ0, thin_by=1)
    got = smp.samples.get_occupancies(flat=False)
    print(label, "Rstride", np.diff(gp._tables()["expansion"].site_rec_off).max(), end=": ")
    ok = True
    for w in range(W):
        subl = M.oracle_sublattices(O, ens.sublattices)
        k = O.Metropolis(O.Ensemble(op, subl), O.Swap(subl), 2000.0, seed=w, walker=w)
        occ = occ0[w].copy()
        for s in range(nst):
            st = k.single_step(occ)
            if not np.array_equal(occ, got[s, w]):
                print(f"walker {w} diverged at step {s} step={st.step}", end="; "); ok = False
                break
    print("OK" if ok else "FAIL")

G = int(os.environ.get("G", 32))
run(M.fcc_subspace(), np.eye(3, dtype=int) * 3, G=G, label="fcc3 S_FCC")
big = L.ClusterSubspace.from_cutoffs(L.fcc_prim(), {2: 7.5, 3: 5.0, 4: 4.2})
run(big, np.eye(3, dtype=int) * 4, G=G, label="fcc4 big")
run(M.rocksalt_subspace(anions=("O2-", "F-")), np.eye(3, dtype=int) * 3, G=G, label="rs3 swap")
run(M.rocksalt_subspace(), np.eye(3, dtype=int) * 3, G=G, label="rs3 cation-only swap")
