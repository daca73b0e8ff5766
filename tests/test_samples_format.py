"""Samples store: the reference's .npy + .pkl layout (hmclab/Samples.py:146-171, 324-333),
batched (chain-concatenated) writes, early-stop compaction and error behaviour."""
import os
import pickle

import numpy as np
import pytest

from hmclab_b200.Samples import Samples, combine_samples


def _write(path, chains=3, per_chain=5, dims=4, rows_written=None, overwrite=True):
    s = Samples(path, mode="w", overwrite=overwrite)
    s.allocate(chains, per_chain, dims)
    rng = np.random.default_rng(0)
    data = rng.normal(size=(per_chain, chains, dims + 1))
    n = per_chain if rows_written is None else rows_written
    if n >= 2:
        s.write_block(data[:2])
        s.write_block(data[2:n])
    elif n == 1:
        s.write_block(data[:1])
    s.write_attribute("proposals", per_chain)
    s.close()
    return data


def test_roundtrip_layout_is_the_references(tmp_path):
    path = str(tmp_path / "run.npy")
    data = _write(path)
    raw = np.load(path)                       # on disk: (n, d+1), chains one after the other
    assert raw.shape == (15, 5)
    assert np.array_equal(raw.reshape(3, 5, 5), data.transpose(1, 0, 2))
    with open(path + ".pkl", "rb") as f:
        attrs = pickle.load(f)
    assert attrs["write_index"] == 15 and attrs["last_written_sample"] == 14
    assert attrs["chains"] == 3 and attrs["samples_per_chain"] == 5
    with Samples(path) as s:                  # reader view: (d+1, n)
        assert s.numpy.shape == (5, 15)
        assert np.array_equal(s.samples, raw.T[:-1])
        assert np.array_equal(s.misfits, raw.T[-1])
        assert np.array_equal(s.chain(1), data[:, 1, :].T)
        assert np.array_equal(s[:, 3], raw[3])
    with Samples(path, burn_in=4) as s:
        assert s.numpy.shape == (5, 11)
    assert combine_samples([path, path]).shape == (5, 30)


def test_interrupted_run_is_compacted(tmp_path):
    path = str(tmp_path / "part.npy")
    data = _write(path, rows_written=3)
    raw = np.load(path)
    assert raw.shape == (9, 5)
    assert np.array_equal(raw.reshape(3, 3, 5), data[:3].transpose(1, 0, 2))
    with Samples(path) as s:
        assert s.read_attribute("write_index") == 9
        assert s.read_attribute("samples_per_chain") == 3


def test_error_behaviour(tmp_path):
    path = str(tmp_path / "x.npy")
    _write(path)
    with pytest.raises(FileExistsError, match="already existing file"):
        Samples(path, mode="w")
    import gc

    gc.collect()  # the refused writer must not touch the existing files when it is collected
    with Samples(path) as s:
        assert s.read_attribute("write_index") == 15
    Samples(path, mode="w", overwrite=True).close()
    with pytest.raises(FileNotFoundError):
        Samples(str(tmp_path / "missing.npy"))
    with pytest.raises(AttributeError, match="extension"):
        Samples(str(tmp_path / "x.csv"), mode="w")
    with pytest.raises(NotADirectoryError):
        Samples(str(tmp_path / "nodir" / "x.npy"), mode="w")
    with pytest.raises(AttributeError):
        Samples(path, mode="w", burn_in=3)
    _write(path, overwrite=True)
    with pytest.raises(ValueError, match="burn-in"):
        Samples(path, burn_in=15)


def test_hdf5_needs_h5py_and_says_so(tmp_path):
    from hmclab_b200.Samples import _have_h5py

    if _have_h5py():
        pytest.skip("h5py present")
    with pytest.raises(ImportError, match="h5py"):
        Samples(str(tmp_path / "run.h5"), mode="w")


@pytest.mark.skipif(not os.path.isdir("/root/reference/hmclab"), reason="reference not mounted")
def test_reference_reader_opens_our_files(tmp_path):
    from _reference_shim import import_reference

    hmclab = import_reference()
    path = str(tmp_path / "ours.npy")
    data = _write(path, chains=1, per_chain=6, dims=3)
    for key, val in dict(sampler="Hamiltonian Monte Carlo", acceptance_rate=0.5, online_thinning=1,
                         start_time="a", end_time="b", runtime="c", runtime_seconds=1.0).items():
        with open(path + ".pkl", "rb") as f:
            attrs = pickle.load(f)
        attrs[key] = val
        with open(path + ".pkl", "wb") as f:
            pickle.dump(attrs, f)
    ref = hmclab.Samples(path, burn_in=1)
    assert ref.numpy.shape == (4, 5)
    assert np.array_equal(np.asarray(ref.numpy), data[1:, 0, :].T)
    assert ref.read_attribute("write_index") == 6
    ref.close()
