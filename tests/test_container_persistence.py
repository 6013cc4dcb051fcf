"""SampleContainer persistence and streaming (smol/moca/sampler/container.py:420-692,
sampler.py:271-297): flush / get_backend / to_hdf5 / from_hdf5 over the directory backend (h5py is not
in this image; with h5py the same calls write HDF5), as_dict / from_dict with the reference's keys, and
the deferred-chunk bookkeeping behind ``Sampler.run(block=False)``.  Host logic only (no GPU)."""
import json

import numpy as np
import pytest

from smol_b200.container import DirectoryBackend, SampleContainer
from smol_b200.sublattice import Sublattice


class _Ens:
    sublattices = [Sublattice(("A", "B"), np.arange(4))]
    natural_parameters = np.array([1.0, 2.0])
    num_energy_coefs = 2
    num_sites = 4
    chemical_potentials = None
    thermo_boundaries = {"temperature": 300.0}


SHAPES = {"occupancy": ((4,), np.int32), "features": ((2,), np.float64), "enthalpy": ((1,), np.float64),
          "accepted": ((1,), bool), "temperature": ((1,), np.float64), "n_accepted": ((), np.int32)}


def _chunk(n, seed, W=3):
    rng = np.random.default_rng(seed)
    return dict(occupancy=rng.integers(0, 2, (n, W, 4)).astype(np.int8), features=rng.normal(size=(n, W, 2)),
                enthalpy=rng.normal(size=(n, W, 1)), accepted=rng.integers(0, 2, (n, W, 1)).astype(bool),
                temperature=np.full((n, W, 1), 300.0), n_accepted=rng.integers(0, 5, (n, W)).astype(np.int32))


def test_flush_to_backend_appends_and_rewinds(tmp_path):
    c = SampleContainer(_Ens(), 3, SHAPES, {"kernel": "Metropolis", "seeds": [1, 2, 3]})
    a, b = _chunk(5, 0), _chunk(5, 1)
    path = str(tmp_path / "run.lmc")
    backend = c.get_backend(path, alloc_nsamples=10)
    assert backend["trace"].attrs["nsamples"] == 0 and len(backend["trace"]["occupancy"]) == 10
    assert backend["trace"]["occupancy"].dtype == np.int32 and backend["trace"]["accepted"].dtype == np.bool_
    c.append(a, thinned_by=4)
    c.flush_to_backend(backend)
    # container.py:435-437: the write position is rewound, the flushed samples stay readable (keep_last_chunk)
    assert len(c) == 0 and c.total_mc_steps == 0
    assert np.array_equal(c.get_occupancies(flat=False), a["occupancy"])
    assert backend["trace"].attrs["nsamples"] == 5 and backend["trace"].attrs["total_mc_steps"] == 20
    c.append(b, thinned_by=4)
    assert len(c) == 5 and np.array_equal(c.get_enthalpies(flat=False), b["enthalpy"])     # replaced, not appended
    c.flush_to_backend(backend)
    backend.close()
    # a second reader sees everything (SWMR: attributes are replaced atomically after the data is flushed)
    r = DirectoryBackend(path, "r")
    assert r["trace"].attrs["nsamples"] == 10 and r["trace"].attrs["total_mc_steps"] == 40
    assert np.array_equal(r["trace"]["occupancy"][:10], np.concatenate([a["occupancy"], b["occupancy"]]).astype(np.int32))
    assert np.array_equal(r["trace"]["features"][5:10], b["features"])
    meta = json.loads(r["metadata"]["sampling_metadata"])
    assert meta["kernel"] == "Metropolis" and "n_accepted" not in r["trace"].keys()
    r.close()
    loaded = SampleContainer.from_hdf5(path)
    assert loaded.num_samples == 10 and loaded.total_mc_steps == 40
    assert np.array_equal(loaded.get_occupancies(flat=False)[:5], a["occupancy"])
    assert loaded.get_occupancies().dtype == np.int32
    assert np.allclose(loaded.get_sublattice_compositions(loaded.sublattices[0]).sum(axis=-1), 1.0)
    assert np.array_equal(loaded.natural_parameters, _Ens.natural_parameters)


def test_to_hdf5_keeps_samples_and_appends_to_existing_file(tmp_path):
    c = SampleContainer(_Ens(), 3, SHAPES)
    c.append(_chunk(4, 2), thinned_by=2)
    path = str(tmp_path / "samples.lmc")
    c.to_hdf5(path)
    assert len(c) == 4 and c.total_mc_steps == 8                       # container.py:622-628
    c.to_hdf5(path)                                                    # the file grows (get_backend -> _grow_backend)
    loaded = SampleContainer.from_hdf5(path, ensemble=_Ens())
    assert loaded.num_samples == 8 and loaded.total_mc_steps == 16
    # incompatible dimensions (container.py:481-488)
    other = SampleContainer(_Ens(), 2, SHAPES)
    with pytest.raises(RuntimeError, match="incompatible dimensions"):
        other.get_backend(path)
    # unfinished run: allocated but unwritten rows are not loaded (container.py:648-656)
    c2 = SampleContainer(_Ens(), 3, SHAPES)
    c2.append(_chunk(2, 3), thinned_by=1)
    p2 = str(tmp_path / "unfinished.lmc")
    b = c2.get_backend(p2, alloc_nsamples=6)
    c2.flush_to_backend(b)
    b.close()
    with pytest.warns(UserWarning, match="unifinished"):
        assert SampleContainer.from_hdf5(p2).num_samples == 2


def test_as_dict_round_trip_uses_reference_keys():
    c = SampleContainer(_Ens(), 3, SHAPES, {"kernel": "Metropolis"})
    tr = _chunk(3, 4)
    c.append(tr, thinned_by=5)
    d = c.as_dict()
    assert set(d) == {"@module", "@class", "ensemble", "metadata", "total_mc_steps", "nsamples", "trace",
                      "aux_checkpoint"}                                  # container.py:531-540
    assert d["nsamples"] == 3 and d["total_mc_steps"] == 15 and d["@class"] == "SampleContainer"
    d = json.loads(json.dumps(d))                                        # JSON serialisable
    back = SampleContainer.from_dict(d)
    assert back.num_samples == 3 and back.total_mc_steps == 15
    assert np.array_equal(back.get_occupancies(flat=False), tr["occupancy"])
    assert np.array_equal(back.get_trace_value("accepted", flat=False), tr["accepted"])
    assert np.allclose(back.get_feature_vectors(flat=False), tr["features"])
    assert back.sublattices[0].species == ("A", "B") and back.metadata["kernel"] == "Metropolis"
    with pytest.raises(ValueError, match="do not match"):
        class Two(_Ens):
            sublattices = _Ens.sublattices * 2
        SampleContainer.from_dict(d, ensemble=Two())


def test_deferred_chunks_count_and_resolve_on_access():
    c = SampleContainer(_Ens(), 3, SHAPES)
    a, b = _chunk(2, 5), _chunk(3, 6)
    calls = []

    def resolve():
        calls.append(1)
        return b, 7, None
    c.append(a, thinned_by=7)
    c.defer(3, 7, resolve)
    assert c.num_samples == 5 and c.has_deferred and not calls            # counted, not yet resolved
    assert c.get_enthalpies(flat=False).shape == (5, 3, 1) and calls == [1] and not c.has_deferred
    assert c.total_mc_steps == 35
    c.defer(3, 7, resolve)
    c.clear()                                                            # waits for the chunk, then drops it
    assert calls == [1, 1] and c.num_samples == 0
