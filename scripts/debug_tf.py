import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smol_b200 as S
from smol_b200 import lattice as L
from oracle import lmc_oracle as O
from tests import models as M

sub = M.rocksalt_subspace(anions=("O2-", "F-"))
scm = np.eye(3, dtype=int) * 3
rng = np.random.default_rng(21)
coefs = rng.normal(0, 0.03, sub.num_corr_functions)
it = L.cluster_interaction_tensors(sub, coefs)
ewm, ewi = L.ewald_matrix(sub, scm)
use_ew = os.environ.get("EW", "1") == "1"
use_mu = os.environ.get("MU", "1") == "1"
mus = {"Li+": 0.0, "Mn3+": 0.4, "Ti4+": -0.3, "O2-": 0.1, "F-": 0.0} if use_mu else None
if use_ew:
    comp = S.CompositeProcessor(sub, scm)
    comp.add_processor(S.ClusterDecompositionProcessor(sub, scm, it))
    comp.add_processor(S.EwaldProcessor(sub, scm, coefficient=0.05, ewald_matrix=ewm, ewald_inds=ewi))
    ora_p = O.CompositeProcessor([O.ClusterDecompositionProcessor(sub, scm, it), O.EwaldProcessor(ewm, ewi, 0.05)])
else:
    comp = S.ClusterDecompositionProcessor(sub, scm, it)
    ora_p = O.ClusterDecompositionProcessor(sub, scm, it)
ens_g = S.Ensemble(comp, chemical_potentials=mus)
table = [[-1, 1, 0, 2, -2], [0, -1, 1, 1, -1]]
W, ncell = 3, 27
occ0 = np.zeros((W, 2 * ncell), dtype=np.int32)
for w in range(W):
    occ0[w, :ncell] = rng.permutation(np.array([0] * 20 + [1] * 4 + [2] * 3))
    occ0[w, ncell:] = rng.permutation(np.array([0] * 17 + [1] * 10))
seeds = np.arange(900, 900 + W)
G = int(os.environ.get("G", 32))
smp = S.Sampler.from_ensemble(ens_g, 2000.0, step_type="table_flip", nwalkers=W, seeds=list(seeds),
                              flip_table=table, swap_weight=0.2, group_size=G)
nst = 400
smp.run(nst, occ0, thin_by=1)
got = smp.samples.get_occupancies(flat=False)
gacc = smp.samples.get_trace_value("accepted", flat=False)[:, :, 0]
genth = smp.samples.get_enthalpies(flat=False)[:, :, 0]
for w in range(W):
    subl = M.oracle_sublattices(O, ens_g.sublattices)
    ens_o = O.Ensemble(ora_p, subl, chemical_potentials=mus)
    k = O.Metropolis(ens_o, O.TableFlip(subl, table, swap_weight=0.2), 2000.0, seed=int(seeds[w]), walker=w)
    occ = occ0[w].copy()
    enth = k.compute_initial_trace(occ).enthalpy[0]
    for s in range(nst):
        before = occ.copy()
        st = k.single_step(occ)
        if st.accepted:
            enth += st.denthalpy
        if not np.array_equal(occ, got[s, w]) or bool(gacc[s, w]) != bool(st.accepted):
            print(f"walker {w} step {s}: oracle step={st.step} accepted={st.accepted} dH={st.denthalpy:.12g} exponent={st.exponent:.12g}")
            print("   gpu accepted", gacc[s, w], "gpu enth", genth[s, w], "oracle enth", enth, "prev gpu enth", genth[s-1, w] if s else None)
            d = np.where(got[s, w] != before)[0]
            print("   gpu changed sites", d, "->", got[s, w][d], "from", before[d])
            break
    else:
        print("walker", w, "OK")
