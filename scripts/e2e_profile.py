import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, warnings
import bench
import smol_b200 as S
from tests import models as M
sub, scm, coefs, it = bench.build_model()
W, N = 4096, 512
ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it))
occ = torch.from_numpy(M.random_occupancies(sub, scm, W, seed=0, balanced=True).astype(np.int32)).pin_memory()
smp = S.Sampler.from_ensemble(ens, 1000.0, step_type="swap", nwalkers=W, seeds=list(range(W)))
warnings.simplefilter("ignore")
for _ in range(3):
    smp.run(N * 8, occ, thin_by=N); smp.clear_samples()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in range(10):
    smp.run(N * 8, occ, thin_by=N); smp.clear_samples()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print("per run ms", dt / 10 * 1e3, "kernel ms", smp.last_kernel_ms)
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
