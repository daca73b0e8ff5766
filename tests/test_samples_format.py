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
    with Samples(path, burn_in=4) as s:       # burn-in is per chain: 3 chains x (5 - 4) rows stay
        assert s.numpy.shape == (5, 3)
        assert np.array_equal(s.numpy, data[4:, :, :].transpose(2, 1, 0).reshape(5, 3))
        assert np.array_equal(s.chain(2), data[4:, 2, :].T)
    with pytest.raises(ValueError, match="burn-in"):
        Samples(path, burn_in=5)              # as long as one chain: nothing would be left
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


def _libhdf5_file():
    """A file written by libhdf5 itself: SciPy ships a MATLAB v7.3 (= HDF5 behind a 512-byte user
    block) test file."""
    import scipy.io

    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data",
                        "testhdf5_7.4_GLNX86.mat")
    if not os.path.isfile(path):
        pytest.skip("SciPy's HDF5 test file is not installed")
    return path


def test_native_hdf5_reader_parses_a_file_written_by_libhdf5():
    """Pins the reader (and with it the writer's view of the format, which the next test reads
    back through the same reader) to real libhdf5 output: superblock, root symbol table, group
    B-tree, local heap, version 1 object header, datatype / dataspace / attribute messages."""
    from hmclab_b200 import _hdf5

    arr, attrs = _hdf5.open_dataset(_libhdf5_file(), "testdouble")
    assert arr.shape == (9, 1) and arr.dtype == np.dtype("<f8")
    assert np.allclose(np.asarray(arr)[:, 0], np.linspace(0, 2 * np.pi, 9))
    assert attrs == {"MATLAB_class": "double"}


def test_native_hdf5_messages_are_byte_identical_to_libhdf5s():
    """The float64 datatype, the rank-2 dataspace, a scalar string attribute and the default
    fill-value message encode to exactly the bytes libhdf5 wrote into that file."""
    from hmclab_b200 import _hdf5

    f = _hdf5._File(_libhdf5_file())
    msgs = dict((t, d) for t, d in f.messages(f.links(f.root_header)["testdouble"]))
    assert _hdf5._dtype_message("<f8") == msgs[0x0003][:20]
    assert _hdf5._dataspace_message((9, 1)) == msgs[0x0001]
    assert _hdf5._message(0x0005, b"\x01\x02\x02\x01\0\0\0\0")[8:] == msgs[0x0005]
    mine = _hdf5._attribute_message("MATLAB_class", b"double")[8:]
    # libhdf5 stored a null-terminated ASCII string of 6 bytes; ours is null padded UTF-8 of 7:
    # same layout, the two datatype flag/size bytes and the padding differ
    theirs = msgs[0x000C]
    assert mine[:24] == theirs[:24] and mine[32:40] == theirs[32:40] and mine[40:46] == theirs[40:46]


def test_hdf5_samples_file_without_h5py(tmp_path):
    """`.h5` (the reference's default format, Samples.py:127-144): dataset "samples" (d+1, n)
    float64 with the attributes on the dataset, chains one after the other."""
    from hmclab_b200 import _hdf5
    from hmclab_b200.Samples import _have_h5py

    if _have_h5py():
        pytest.skip("h5py present: the h5py backend is used")
    path = str(tmp_path / "run.h5")
    s = Samples(path, mode="w")
    s.allocate(3, 5, 4)
    rng = np.random.default_rng(0)
    data = rng.normal(size=(5, 3, 5))
    s.write_block(data[:2])
    s.write_block(data[2:])
    s.write_attribute("proposals", 5)
    s.write_attribute("acceptance_rate", 0.8125)
    s.write_attribute("sampler", "HMC")
    s.write_attribute("start_time", "2026-10-17 12:00:00.000001")
    s.close()
    with open(path, "rb") as f:
        assert f.read(8) == _hdf5.SIGNATURE
    with Samples(path) as r:
        assert r.numpy.shape == (5, 15)
        assert np.array_equal(r.chain(1), data[:, 1, :].T)
        assert np.array_equal(r.misfits[:, 0], np.concatenate([data[:, c, -1] for c in range(3)]))
        assert r.read_attribute("write_index") == 15 and r.read_attribute("last_written_sample") == 14
        assert r.read_attribute("acceptance_rate") == 0.8125
        assert r.read_attribute("sampler") == "HMC"
        assert r.read_attribute("start_time") == "2026-10-17 12:00:00.000001"
    with Samples(path, burn_in=2) as r:      # per chain: 3 chains x (5 - 2) rows stay
        assert r.numpy.shape == (5, 9)
        assert np.array_equal(r.samples, data[2:, :, :-1].transpose(2, 1, 0).reshape(4, 9))
        assert np.array_equal(r.misfits[:, 0], data[2:, :, -1].T.reshape(9))
    with pytest.raises(FileExistsError):
        Samples(path, mode="w")
    assert combine_samples([path, path]).shape == (5, 30)
    # a run that stopped early is compacted to the written samples
    part = str(tmp_path / "part")          # no extension = HDF5, as in the reference
    s = Samples(part, mode="w")
    s.allocate(3, 5, 4)
    s.write_block(data[:2])
    s.close()
    with Samples(part) as r:
        assert r.numpy.shape == (5, 6) and r.read_attribute("samples_per_chain") == 2
        assert np.array_equal(r.chain(2), data[:2, 2, :].T)
    arr, attrs = _hdf5.open_dataset(part + ".h5")
    assert os.path.getsize(part + ".h5") < _hdf5.DATA_OFFSET + 8 * 75 + 4096 and attrs["chains"] == 3


def test_native_hdf5_attributes_above_64_kb(tmp_path):
    """An autotuned run stores per-proposal `stepsizes` / `acceptance_rates` (Samplers.py:1340-1341,
    443-451 of the host mirror): above 64 KB they cannot be header messages; they become companion
    datasets and come back as attributes, and the file stays readable."""
    from hmclab_b200 import _hdf5
    from hmclab_b200.Samples import _have_h5py

    if _have_h5py():
        pytest.skip("h5py present: the h5py backend is used")
    path = str(tmp_path / "tuned.h5")
    s = Samples(path, mode="w")
    s.allocate(2, 3, 4)
    s.write_block(np.arange(30, dtype=float).reshape(3, 2, 5))
    steps = np.random.default_rng(1).uniform(size=(10000, 2))
    rates = np.random.default_rng(2).uniform(size=(10000, 2))
    s.write_attribute("stepsizes", steps)
    s.write_attribute("acceptance_rates", rates)
    s.write_attribute("final_stepsizes", np.ones(20000))
    s.write_attribute("small", np.ones((7, 1)))
    s.close()
    with Samples(path) as r:
        assert r.read_attribute("write_index") == 6
        assert np.array_equal(r.read_attribute("stepsizes"), steps)
        assert np.array_equal(r.read_attribute("acceptance_rates"), rates)
        assert r.read_attribute("final_stepsizes").shape == (20000,)
        assert r.read_attribute("small").shape == (7, 1)
        assert r.numpy.shape == (5, 6)
    f = _hdf5._File(path)
    assert sorted(f.links(f.root_header)) == ["samples", "samples.acceptance_rates",
                                              "samples.final_stepsizes", "samples.stepsizes"]


def test_native_hdf5_attribute_kinds_and_edge_cases(tmp_path):
    """Scalars of every kind the samplers store, arrays, non-ASCII text, long names, an empty
    dataset, a file behind a user block, and refusal of things the reader does not cover."""
    from hmclab_b200 import _hdf5

    path = str(tmp_path / "attrs.h5")
    w = _hdf5.Writer(path, (2, 3))
    w.data[:] = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]
    attrs = {"i": 7, "neg": -3, "np_i": np.int32(9), "f": 0.125, "np_f": np.float32(1.5), "flag": True,
             "text": "Hamiltonian Monte Carlo", "unicode": "étape ε=0.1", "empty": "", "none": None,
             "vector": np.array([1.0, 2.0, 3.0]), "ints": np.arange(4), "names": np.array(["lf", "4s"]),
             "a_rather_long_attribute_name_that_needs_padding_to_eight_bytes": 1}
    w.close(attrs)
    arr, got = _hdf5.open_dataset(path)
    assert np.array_equal(arr, [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    assert got["i"] == 7 and got["neg"] == -3 and got["np_i"] == 9 and got["flag"] == 1
    assert got["f"] == 0.125 and got["np_f"] == 1.5
    assert got["text"] == "Hamiltonian Monte Carlo" and got["unicode"] == "étape ε=0.1"
    assert got["empty"] == "" and got["none"] == "None"
    assert np.array_equal(got["vector"], [1.0, 2.0, 3.0]) and np.array_equal(got["ints"], np.arange(4))
    assert list(got["names"]) == ["lf", "4s"]
    assert got["a_rather_long_attribute_name_that_needs_padding_to_eight_bytes"] == 1
    # every structure sits on an 8-byte boundary and the end-of-file address is the file size
    with open(path, "rb") as f:
        raw = f.read()
    eof, root = np.frombuffer(raw[40:48], "<u8")[0], np.frombuffer(raw[64:72], "<u8")[0]
    assert eof == len(raw) and root % 8 == 0 and raw[root] == 1
    # empty dataset
    empty = str(tmp_path / "empty.h5")
    _hdf5.Writer(empty, (5, 0)).close({"write_index": 0})
    arr, got = _hdf5.open_dataset(empty)
    assert arr.shape == (5, 0) and got["write_index"] == 0
    # the same file behind a 512-byte user block (as MATLAB writes them) is still found
    shifted = str(tmp_path / "shifted.h5")
    with open(shifted, "wb") as f:
        f.write(b"\0" * 512)
        sb = bytearray(raw)
        sb[24:32] = np.array([512], "<u8").tobytes()       # base address
        f.write(bytes(sb))
    arr, got = _hdf5.open_dataset(shifted)
    assert np.array_equal(arr, [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]) and got["i"] == 7
    with pytest.raises(KeyError):
        _hdf5.open_dataset(path, "missing")
    notes = str(tmp_path / "notes.h5")
    with open(notes, "wb") as f:
        f.write(b"not an hdf5 file" * 10)
    with pytest.raises(ValueError):
        _hdf5.open_dataset(notes)
    with pytest.raises(FileExistsError):
        _hdf5.Writer(path, (1, 1))
