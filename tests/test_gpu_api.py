"""GPU tests of the round-2 host API: device-resident entry, non-blocking back-to-back runs, streaming to a
file backend, device selection, 2-rank sharding on real GPUs, and the engine limits that only
``lmc_model_create`` can see."""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

from tests import models as M
from tests import workloads as WK

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fcc_sampler(W, n=4, **kw):
    import smol_b200 as S
    from smol_b200 import lattice as L
    sub = M.fcc_subspace()
    scm = np.eye(3, dtype=int) * n
    it = L.cluster_interaction_tensors(sub, M.fcc_coefs(sub))
    ens = S.Ensemble(S.ClusterDecompositionProcessor(sub, scm, it))
    occ0 = M.random_occupancies(sub, scm, W, seed=1, balanced=True)
    smp = S.Sampler.from_ensemble(ens, 1000.0, step_type="swap", nwalkers=W, seeds=list(range(10, 10 + W)), **kw)
    return smp, occ0


def _traces(samples):
    return {k: samples.get_trace_value(k, flat=False).copy() for k in ("occupancy", "features", "enthalpy", "accepted", "n_accepted")}


def test_nonblocking_back_to_back_runs_equal_blocking_runs(cuda_device):
    """run(block=False) + detach_samples: same chains, same traces, whatever the interleaving of reads"""
    W, n, thin = 64, 64 * 10, 64
    a, occ0 = _fcc_sampler(W)
    a.run(n, occ0, thin_by=thin)
    a.run(n, thin_by=thin)
    want = _traces(a.samples)
    b, _ = _fcc_sampler(W)
    b.run(n, occ0, thin_by=thin, block=False)
    first = b.detach_samples()
    assert first.num_samples == 10 and first.has_deferred
    b.run(n, thin_by=thin, block=False)          # resolves nothing of `first`: it was detached
    second = b.detach_samples()
    got2, got1 = _traces(second), _traces(first)  # read in reverse order
    for k in want:
        np.testing.assert_array_equal(np.concatenate([got1[k], got2[k]]), want[k])
    # without detaching, the deferred tail is appended in order
    c, _ = _fcc_sampler(W)
    c.run(n, occ0, thin_by=thin, block=False)
    c.run(n, thin_by=thin, block=False)
    assert c.samples.num_samples == 20
    got = _traces(c.samples)
    for k in want:
        np.testing.assert_array_equal(got[k], want[k])
    first.clear(), second.clear()


def test_device_resident_entry_equals_host_entry(cuda_device):
    """CUDA-tensor occupancies in, device traces out (run_device) == the host path"""
    import torch
    W, n, thin = 32, 64 * 6, 64
    a, occ0 = _fcc_sampler(W)
    a.run(n, occ0, thin_by=thin)
    want = _traces(a.samples)
    b, _ = _fcc_sampler(W)
    occ_dev = torch.from_numpy(occ0).to(cuda_device)
    out = b.run_device(n // 2, occ_dev, thin_by=thin)
    first = {k: v.clone() for k, v in out.items()}
    out2 = b.run_device(n // 2, None, thin_by=thin, out=out, reuse_state=True)      # continues, no re-evaluation
    assert out2 is out and all(v.is_cuda for v in out.values())
    for k, name in (("occupancy", "occupancy"), ("n_accepted", "n_accepted")):
        got = torch.cat([first[k], out[k]]).cpu().numpy()
        np.testing.assert_array_equal(got.astype(want[name].dtype), want[name])
    enth = torch.cat([first["enthalpy"], out["enthalpy"]]).cpu().numpy()
    np.testing.assert_allclose(enth, want["enthalpy"][:, :, 0], rtol=1e-12, atol=1e-10)
    assert b.samples.num_samples == 0                                      # nothing was copied to the host
    # int64 CUDA input is converted on the device; run() accepts CUDA tensors as well
    c, _ = _fcc_sampler(W)
    c.run(n, occ_dev.to(torch.int64), thin_by=thin)
    np.testing.assert_array_equal(c.samples.get_occupancies(flat=False), want["occupancy"])
    assert np.array_equal(occ_dev.cpu().numpy(), occ0)                       # the input is never modified


def test_streaming_run_writes_the_backend_in_chunks(cuda_device, tmp_path):
    """Sampler.run(stream_chunk, stream_file, keep_last_chunk), sampler.py:271-297"""
    from smol_b200.container import SampleContainer
    W, thin, nsteps = 16, 32, 32 * 24
    a, occ0 = _fcc_sampler(W)
    a.run(nsteps, occ0, thin_by=thin)
    want = _traces(a.samples)
    b, _ = _fcc_sampler(W)
    path = str(tmp_path / "stream.lmc")
    b.run(nsteps, occ0, thin_by=thin, stream_chunk=32 * 8, stream_file=path)
    assert b.samples.num_samples == 0                                        # keep_last_chunk=False clears
    loaded = SampleContainer.from_hdf5(path, ensemble=b.ensemble)
    assert loaded.num_samples == 24 and loaded.total_mc_steps == nsteps
    np.testing.assert_array_equal(loaded.get_occupancies(flat=False), want["occupancy"])
    np.testing.assert_array_equal(loaded.get_feature_vectors(flat=False), want["features"])
    np.testing.assert_array_equal(loaded.get_trace_value("accepted", flat=False), want["accepted"])
    # a second run appends to the same file; keep_last_chunk leaves the last chunk readable in memory
    b.run(nsteps, thin_by=thin, stream_chunk=32 * 8, stream_file=path, keep_last_chunk=True)
    assert np.array_equal(b.samples.get_occupancies(flat=False).shape, (8, W, 64))
    assert SampleContainer.from_hdf5(path).num_samples == 48
    with pytest.raises(ValueError, match="divisor"):
        b.run(100, thin_by=10, stream_chunk=30, stream_file=path)
    with pytest.raises(ValueError, match="multiple"):
        b.run(120, thin_by=7, stream_chunk=30, stream_file=path)


def test_sampler_on_a_device_that_is_not_current(cuda_device):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    a, occ0 = _fcc_sampler(8)
    a.run(256, occ0, thin_by=64)
    torch.cuda.set_device(0)
    b, _ = _fcc_sampler(8, device="cuda:1")
    b.run(256, occ0, thin_by=64)                  # device 0 stays current
    assert torch.cuda.current_device() == 0
    np.testing.assert_array_equal(b.samples.get_occupancies(flat=False), a.samples.get_occupancies(flat=False))


_SHARD_SCRIPT = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from tests.test_gpu_api import _fcc_sampler
from smol_b200.dist import ShardedSampler
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
W = 24
one, occ0 = _fcc_sampler(W)
sh = ShardedSampler(one.ensemble, W, list(range(10, 10 + W)), 1000.0, step_type="swap")
sh.run(512, occ0, thin_by=64)
occ = sh.gather("occupancy"); enth = sh.gather("enthalpy")
assert occ.is_cuda and occ.shape == (8, W, 64), occ.shape
if rank == 0:
    one.run(512, occ0, thin_by=64)
    np.testing.assert_array_equal(occ.cpu().numpy(), one.samples.get_occupancies(flat=False))
    np.testing.assert_allclose(enth.cpu().numpy(), one.samples.get_enthalpies(flat=False), rtol=1e-12, atol=1e-10)
    print("SHARD_OK")
dist.barrier(); dist.destroy_process_group()
"""


def test_sharded_sampler_on_two_gpus_equals_one_rank(cuda_device, tmp_path):
    """walker_id_base invariance on hardware: 2 ranks x 12 walkers == 1 rank x 24 walkers, traces gathered
    over NCCL from device tensors"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    script = tmp_path / "shard.py"
    script.write_text(_SHARD_SCRIPT % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_config5_full_cell_table_flips_bit_exact_vs_oracle(cuda_device):
    """12x12x12 five-species rocksalt + Ewald (N = 3456, E = 8640), charge-neutral table flips: 2 walkers x 600
    steps against the oracle's TableFlip (kernel/mcusher.py:553-711), occupancies bit-exact; both Ewald paths of the
    classic kernel and the speculative table-flip kernel (the default while few steps are accepted)"""
    from oracle import lmc_oracle as O
    wk = WK.get(5)
    ens = wk.product_ensemble()
    W = 2
    occ0 = wk.initial_occupancies(W, seed=77)
    seeds = [4242, 4243]
    kernels = wk.oracle_kernels(seeds, walker0=0, use_ref=O.load_ref() is not None)
    ref = O.run_sampler(kernels, occ0, 600, 100)
    assert ref["n_accepted"].sum() > 100
    scale = np.abs(ref["features"]).max()
    for field, spec_mode in ((False, 1), (True, 1), (True, 2), ("auto", 0)):
        smp = wk.sampler(ens, W, seeds, ewald_field=field, spec_mode=spec_mode)
        smp.run(600, occ0, thin_by=100)
        s = smp.samples
        np.testing.assert_array_equal(s.get_occupancies(flat=False), ref["occupancy"])
        np.testing.assert_array_equal(s.get_trace_value("n_accepted", flat=False), ref["n_accepted"])
        np.testing.assert_allclose(s.get_feature_vectors(flat=False), ref["features"], rtol=1e-10, atol=1e-10 * scale)
        np.testing.assert_allclose(s.get_enthalpies(flat=False), ref["enthalpy"], rtol=1e-10,
                                   atol=1e-10 * np.abs(ref["enthalpy"]).max())


def test_config3_at_4096_walkers(cuda_device):
    """BASELINE config 3 at its full walker count: walkers picked across the range bit-exact vs the C oracle, all
    walkers: anion sublattice untouched, running features == full re-evaluation"""
    from oracle import c_oracle as CO
    from oracle import lmc_oracle as O
    wk = WK.get(3)
    ens = wk.product_ensemble()
    W = 4096
    occ0 = wk.initial_occupancies(W)
    seeds = np.arange(W) + 1000
    smp = wk.sampler(ens, W, seeds)
    smp.run(1024, occ0, thin_by=512)
    s = smp.samples
    occ_tr = s.get_occupancies(flat=False)
    assert np.all(occ_tr[:, :, 512:] == 0)
    feats = s.get_feature_vectors(flat=False)
    full_last = ens.compute_feature_vector_batch(occ_tr[-1])
    scale = np.abs(full_last).max()
    np.testing.assert_allclose(feats[-1], full_last, rtol=1e-10, atol=1e-10 * scale)
    co = CO.COracle(O.Ensemble(wk.oracle_processor(), wk.oracle_sublattices(), chemical_potentials=wk.chemical_potentials()))
    for w in (0, 1234, 4095):
        out, _ = co.run(occ0[w:w + 1], 1024, 512, seeds[w:w + 1], usher="flip", temperature=wk.temperature, walker_base=w)
        np.testing.assert_array_equal(out["occupancy"][:, 0], occ_tr[:, w])
        np.testing.assert_allclose(out["features"][:, 0], feats[:, w], rtol=1e-10, atol=1e-10 * scale)


def test_config4_window_constants(cuda_device):
    wk = WK.get(4)
    ens = wk.product_ensemble()
    e0 = ens.compute_feature_vector_batch(wk.initial_occupancies(64)) @ ens.natural_parameters
    lo, hi = wk.window_from_enthalpies(e0, wk.bin_size)
    assert lo == wk.window()[0] and abs(hi - wk.window()[1]) < 1e-9      # (mean / std rounding differs by ulps)
    assert len(np.arange(lo, hi, wk.bin_size)) == len(np.arange(*wk.window(), wk.bin_size)) == 298


def test_engine_limits_fail_loudly(cuda_device):
    """limits of the engine that the reference does not have raise at model creation"""
    import smol_b200 as S
    from smol_b200 import lattice as L
    # flat tensor stride above 255 (u8 strides for dp4a): 7 species on a 4-site cluster -> stride 343
    species = tuple("A%d" % i for i in range(7))
    sub = L.ClusterSubspace.from_cutoffs(L.fcc_prim(species=species), {2: 3.0, 3: 3.0, 4: 3.0})
    scm = np.eye(3, dtype=int) * 2
    coefs = np.zeros(sub.num_corr_functions)
    with pytest.raises((RuntimeError, ValueError), match="stride"):
        S.Ensemble(S.ClusterExpansionProcessor(sub, scm, coefs)).compute_feature_vector(np.zeros(8, dtype=np.int32))


def test_anneal_streams_every_temperature_to_one_file(cuda_device, tmp_path):
    """Sampler.anneal(..., stream_chunk, stream_file) (sampler.py:303-384): one file, all temperatures, the container
    cleared at the end; progress=True only draws a bar"""
    from smol_b200.container import SampleContainer
    W, thin = 8, 32
    a, occ0 = _fcc_sampler(W)
    a.anneal([1500.0, 1000.0, 500.0], 32 * 6, occ0, thin_by=thin)
    want = _traces(a.samples)
    temps = a.samples.get_trace_value("temperature", flat=False)
    b, _ = _fcc_sampler(W)
    path = str(tmp_path / "anneal.lmc")
    b.anneal([1500.0, 1000.0, 500.0], 32 * 6, occ0, thin_by=thin, stream_chunk=32 * 3, stream_file=path, progress=True)
    assert b.samples.num_samples == 0
    loaded = SampleContainer.from_hdf5(path, ensemble=b.ensemble)
    assert loaded.num_samples == 18 and loaded.total_mc_steps == 3 * 32 * 6
    np.testing.assert_array_equal(loaded.get_occupancies(flat=False), want["occupancy"])
    np.testing.assert_array_equal(loaded.get_trace_value("temperature", flat=False), temps)
    assert sorted(set(temps[:, 0, 0].tolist())) == [500.0, 1000.0, 1500.0]
