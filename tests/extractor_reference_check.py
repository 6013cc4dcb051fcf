"""smol_b200.interop.from_smol_ensemble fed with LIVE objects of the reference's own classes (smol/moca/ensemble.py,
processor/{expansion,ewald,composite}.py, unmodified, imported from /root/reference behind package shells): the packed
device tables must equal those of the natively built ensemble.  Run in a fresh process (tests/test_host_logic.py)."""
import sys, os, json, importlib.util, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
spec = importlib.util.spec_from_file_location("gen", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_python_golden.py"))
gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
objs = gen.reference_ensemble_objects()
import smol_b200 as S
from smol_b200 import interop, _capi as capi
res = {}
for name, (ref_ens, (sub, scm, it, ew, mus)) in objs.items():
    assert type(ref_ens).__module__ == "smol.moca.ensemble", type(ref_ens).__module__
    got = interop.from_smol_ensemble(ref_ens)
    ce = S.ClusterDecompositionProcessor(sub, scm, it)
    if ew is None:
        proc = ce
    else:
        proc = S.CompositeProcessor(sub, scm); proc.add_processor(ce)
        proc.add_processor(S.EwaldProcessor(sub, scm, coefficient=ew[2], ewald_matrix=ew[0], ewald_inds=ew[1]))
    native = S.Ensemble(proc, chemical_potentials=mus)
    a, b = got.packed_model(), native.packed_model()
    ok = len(a.keep) == len(b.keep) and all(np.array_equal(x, y) for x, y in zip(a.keep, b.keep))
    ok = ok and np.array_equal(got.natural_parameters, np.asarray(ref_ens.natural_parameters))
    ok = ok and all(getattr(a.desc, f) == getattr(b.desc, f) for f, ct in capi.LmcModelDesc._fields_ if ct is not ctypes.c_void_p)
    res[name] = bool(ok)
print(json.dumps(res))
