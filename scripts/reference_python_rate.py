"""The reference's OWN Python path timed in the build container (it needs /root/reference, so it cannot run on the
GPU box): smol's unmodified Sampler.run -> Metropolis.single_step -> Swap.propose_step (numpy PCG64, its real RNG use)
-> Ensemble / ClusterDecompositionProcessor.compute_feature_vector_change on the compiled Cython evaluators
(oracle/_ref), for BASELINE config 2 (binary FCC 8x8x8, canonical swap, T = 1000 K).  One process per core.
   python scripts/reference_python_rate.py [seconds] > profiles/r02_reference_python.json"""
import importlib.util, json, multiprocessing as mp, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(args):
    idx, seconds = args
    import numpy as np
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tests", "golden", "make_reference_python_golden.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    gen.import_reference_kernels()
    Sampler = gen.import_reference_sampler()
    CE, CD, RefSubspace = gen.import_reference_processors()
    cm = __import__("types").ModuleType("pymatgen.core.composition"); cm.ChemicalPotential = dict
    sys.modules["pymatgen.core.composition"] = cm
    pp = sys.modules["smol.moca.processor"]
    pp.CompositeProcessor = type("CompositeProcessor", (), {}); pp.EwaldProcessor = type("EwaldProcessor", (), {})
    RefEnsemble = importlib.import_module("smol.moca.ensemble").Ensemble
    sys.modules["smol.moca"].Ensemble = RefEnsemble
    from oracle import lmc_oracle as O
    from tests import workloads as WK
    wk = WK.get(2)
    sub, scm = wk.subspace(), wk.supercell()
    subl = wk.oracle_sublattices()
    for sl in subl:
        sl.site_space = gen._SiteSpace({spc: 1.0 / len(sl.species) for spc in sl.species})
    proc = CD(RefSubspace(sub), scm, wk.interaction_tensors())
    ens = RefEnsemble(proc, sublattices=subl)
    assert type(ens).__module__ == "smol.moca.ensemble" and type(proc).__module__ == "smol.moca.processor.expansion"
    smp = Sampler.from_ensemble(ens, temperature=wk.temperature, step_type="swap", kernel_type="Metropolis",
                                seeds=[1000 + idx], nwalkers=1)
    assert type(smp.mckernels[0]).__module__ == "smol.moca.kernel.metropolis"
    occ = wk.initial_occupancies(1, seed=idx)
    smp.run(2000, occ, thin_by=100, progress=False)          # warm-up
    n, total, t0 = 5000, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        smp.run(n, thin_by=100, progress=False)
        total += n
    return total, time.perf_counter() - t0, float(smp.efficiency())


if __name__ == "__main__":
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    one = worker((0, seconds))
    with ctx.Pool(cores) as pool:
        res = pool.map(worker, [(i, seconds) for i in range(cores)])
    import platform
    print(json.dumps({
        "what": "smol's own Sampler.run / Metropolis.single_step / Swap.propose_step / Ensemble / ClusterDecompositionProcessor "
                "(unmodified Python from /root/reference, numpy PCG64) over its compiled Cython evaluators; BASELINE config 2 "
                "(binary FCC 8x8x8, canonical swap, T=1000K), one walker per process",
        "where": "build container (no GPU): %s, %d logical cores" % (platform.processor() or platform.machine(), cores),
        "single_process_steps_per_s": one[0] / one[1],
        "all_cores_steps_per_s": sum(r[0] / r[1] for r in res), "cores": cores,
        "per_core_steps_per_s_all_cores_busy": sum(r[0] / r[1] for r in res) / cores,
        "acceptance_flag_ratio": one[2], "seconds_per_process": seconds}, indent=1))
