"""Sample store compatible with ``hmclab.Samples`` (hmclab/Samples.py:16-469).

On-disk format is the reference's:

* ``*.npy``  -- a plain NumPy file holding an ``(n, d+1)`` float64 array (one row per stored
  proposal: model, then misfit; the reader transposes, Samples.py:160-161) next to a
  ``<file>.pkl`` pickle of the attribute dictionary (Samples.py:146-151, 163-171);
* ``*.h5`` / no extension -- HDF5 dataset ``"samples"`` of shape ``(d+1, n)`` with the
  attributes on the dataset (Samples.py:127-144, 305-322).  Written and read with ``h5py`` when
  it is installed, otherwise by the native writer/reader of ``hmclab_b200._hdf5`` (classic
  HDF5 layout: version 0 superblock, contiguous dataset, fixed-length string attributes).

Files written here open with the reference's ``hmclab.Samples`` and vice versa.  A batched
run stores its chains one after the other (all rows of chain 0, then chain 1, ...), which
is the layout ``hmclab.Samples.combine_samples`` produces from per-chain files; the
attributes ``chains`` and ``samples_per_chain`` say how to split it again.
"""
from __future__ import annotations

import os as _os
import pickle as _pickle
from typing import List as _List, Union as _Union

import numpy as _numpy


def _have_h5py():
    try:
        import h5py

        return isinstance(getattr(h5py, "__version__", None), str)   # not an import stub
    except Exception:
        return False


class Samples:
    filetype = None
    mode = None
    filename = None
    burn_in = 0

    def __init__(self, filename, burn_in=None, mode="r", overwrite=None):
        stripped, ext = _os.path.splitext(filename)
        if ext in ("", ".h5"):
            self.filetype = "HDF5"
            filename = stripped + ".h5"
        elif ext == ".npy":
            self.filetype = "NPY"
        else:
            raise AttributeError(f"Unkown extension `{ext}` for samples file.")
        self.mode, self.filename = mode, filename
        self._closed = True  # nothing to flush until the constructor has succeeded
        self._attributes = {}
        self._memmap = None
        self._h5 = None
        self._native = None   # hmclab_b200._hdf5.Writer when h5py is not installed

        if mode == "r":
            if overwrite is not None:
                raise AttributeError("Overwrite is not relevant when writing samples.")
            if not _os.path.isfile(filename):
                raise FileNotFoundError(
                    f"Trying to read samples file `{filename}` which does not exist.")
            if self.filetype == "NPY":
                self._array = _numpy.load(filename, mmap_mode="r").T
                with open(f"{filename}.pkl", "rb") as f:
                    self._attributes = _pickle.load(f)
            elif _have_h5py():
                import h5py

                try:
                    self._h5 = h5py.File(filename, "r")
                    self._dataset = self._h5["samples"]
                except Exception as e:
                    raise ValueError(f"Was not able to open the samples file. Exception: {e}")
            else:
                from . import _hdf5

                try:
                    self._dataset, self._attributes = _hdf5.open_dataset(filename, "samples")
                except Exception as e:
                    raise ValueError(f"Was not able to open the samples file. Exception: {e}")
            self.burn_in = 0 if burn_in is None else burn_in
            self.last_sample = self.read_attribute("write_index")
            # a batched run stores chain after chain: burn-in is per chain, so it has to be
            # shorter than ONE chain (a reference file holds one chain: same rule as there)
            if self.last_sample // self._chain_count() <= self.burn_in:
                self._closed = False
                self.close()
                raise ValueError(
                    f"The burn-in phase is longer than the chain itself. "
                    f"Total samples before burn in: {self.last_sample}")
        elif mode == "w":
            if burn_in is not None:
                raise AttributeError("Burn in is not relevant when writing samples.")
            directory = _os.path.dirname(filename)
            if directory != "" and not _os.path.isdir(directory):
                raise NotADirectoryError(
                    f"Trying to write a samples file to a non-existent directory `{directory}`.")
            self.overwrite = bool(overwrite)
            exists = _os.path.isfile(filename) or (
                self.filetype == "NPY" and _os.path.isfile(f"{filename}.pkl"))
            if not self.overwrite and exists:
                shown = filename
                if _os.path.isfile(f"{filename}.pkl"):
                    shown += f"` or attributes file `{filename}.pkl"
                raise FileExistsError(
                    f"Trying to write samples to an already existing file `{shown}`.")
            self._rows_written = 0
            self.write_attribute("write_index", 0)
            self.write_attribute("last_written_sample", -1)
        else:
            raise AttributeError(f"Unkown file mode `{mode}` for samples file.")
        self._closed = False

    @staticmethod
    def _require_h5py():
        if not _have_h5py():
            raise ImportError(
                "Writing/reading HDF5 samples needs the `h5py` package, which is not installed. "
                "Use a `.npy` samples filename (the reference's NumPy variant) instead.")

    # -- batched writer (used by hmclab_b200.Samplers.HMC) ---------------------------------
    def allocate(self, chains: int, samples_per_chain: int, dims: int):
        """Reserve space for ``chains`` x ``samples_per_chain`` rows of ``dims + 1`` values."""
        assert self.mode == "w"
        self._chains, self._per_chain, self._width = int(chains), int(samples_per_chain), int(dims) + 1
        total = self._chains * self._per_chain
        if self.filetype == "NPY":
            if _os.path.isfile(self.filename):
                _os.remove(self.filename)
            self._memmap = _numpy.lib.format.open_memmap(
                self.filename, mode="w+", dtype=_numpy.float64, shape=(total, self._width))
        elif not _have_h5py():
            from . import _hdf5

            self._native = _hdf5.Writer(self.filename, (self._width, total), "samples", overwrite=self.overwrite)
            self._dataset = self._native.data
        else:
            import h5py

            self._h5 = h5py.File(self.filename, "w" if self.overwrite else "w-", libver="latest")
            self._dataset = self._h5.create_dataset(
                "samples", (self._width, total), maxshape=(None, None), dtype="f8", chunks=True)
            for key, value in self._attributes.items():
                self._dataset.attrs[key] = value
        self._allocated = True
        self.write_attribute("chains", self._chains)
        self.write_attribute("samples_per_chain", self._per_chain)

    def write_block(self, block: _numpy.ndarray):
        """``block`` [rows, chains, dims+1]: the next ``rows`` stored proposals of every chain."""
        assert self.mode == "w" and block.ndim == 3 and block.shape[1:] == (self._chains, self._width)
        r0, r1 = self._rows_written, self._rows_written + block.shape[0]
        assert r1 <= self._per_chain, "more rows than allocated"
        if self.filetype == "NPY":
            view = self._memmap.reshape(self._chains, self._per_chain, self._width)
            view[:, r0:r1, :] = block.transpose(1, 0, 2)
        elif self._native is not None:   # memmap of the (d+1, chains * per_chain) dataset: one strided copy
            view = self._dataset.reshape(self._width, self._chains, self._per_chain)
            view[:, :, r0:r1] = block.transpose(2, 1, 0)
        else:
            for c in range(self._chains):
                lo = c * self._per_chain
                self._dataset[:, lo + r0: lo + r1] = block[:, c, :].T
        self._rows_written = r1
        self._attributes["write_index"] = r1 * self._chains
        self._attributes["last_written_sample"] = r1 * self._chains - 1

    def _compact(self):
        """A run that stopped early leaves unwritten rows; drop them so that every stored
        row is a sample (the reference's files only ever hold written samples)."""
        if self._memmap is None and self._h5 is None and self._native is None:
            return
        if self._rows_written == self._per_chain:
            return
        keep = self._rows_written
        if self._native is not None:
            self._native.resize_columns(keep, self._per_chain)
            self._dataset = self._native.data
            self._per_chain = keep
            self._attributes["samples_per_chain"] = keep
            return
        if self.filetype == "NPY":
            data = _numpy.array(
                self._memmap.reshape(self._chains, self._per_chain, self._width)[:, :keep, :])
            del self._memmap
            self._memmap = None
            _numpy.save(self.filename, data.reshape(self._chains * keep, self._width))
        else:
            data = self._dataset[...].reshape(self._width, self._chains, self._per_chain)[:, :, :keep]
            self._dataset.resize((self._width, self._chains * keep))
            self._dataset[...] = data.reshape(self._width, self._chains * keep)
        self._per_chain = keep
        self._attributes["samples_per_chain"] = keep

    # -- attributes ----------------------------------------------------------------------
    def write_attribute(self, name, value):
        assert self.mode == "w"
        self._attributes[name] = value
        if self.filetype == "HDF5" and self._h5 is not None:
            self._dataset.attrs[name] = value

    def read_attribute(self, name):
        if self.filetype == "HDF5" and self._h5 is not None and self.mode == "r":
            return self._dataset.attrs[name]
        return self._attributes[name]

    def _flush_attributes(self):
        if self.filetype == "NPY":
            with open(f"{self.filename}.pkl", "wb") as f:
                _pickle.dump(self._attributes, f)
        elif self._native is not None:
            self._native.close(self._attributes)
            self._native = None
        elif self._h5 is not None:
            for key, value in self._attributes.items():
                self._dataset.attrs[key] = value

    # -- reading -------------------------------------------------------------------------
    def _chain_count(self):
        try:
            return max(1, int(self.read_attribute("chains")))
        except (KeyError, AttributeError):
            return 1       # a file written by the reference: one chain

    def _after_burn_in(self, data):
        """Columns with the first ``burn_in`` samples of EVERY chain removed (a batched run
        stores all rows of chain 0, then chain 1, ...; Samples.py:280-322 holds one chain)."""
        chains = self._chain_count()
        if chains == 1 or self.burn_in == 0:
            return data[:, self.burn_in:]
        per = data.shape[1] // chains
        block = _numpy.asarray(data[:, : chains * per]).reshape(data.shape[0], chains, per)
        return block[:, :, self.burn_in:].reshape(data.shape[0], chains * (per - self.burn_in))

    @property
    def numpy(self):
        if self.filetype == "HDF5":
            return self._after_burn_in(self._dataset)
        return self._after_burn_in(self._array)

    @property
    def samples(self):
        return self.numpy[:-1, :]

    @property
    def misfits(self):
        if self.filetype == "HDF5":
            return self.numpy[-1, :][:, None]
        return self.numpy[-1, :]

    def __getitem__(self, key):
        return self.numpy[key]

    def chain(self, index: int) -> _numpy.ndarray:
        """``(d+1, samples_per_chain)`` block of one chain of a batched run."""
        n = int(self.read_attribute("samples_per_chain"))
        data = self._dataset if self.filetype == "HDF5" else self._array
        return data[:, index * n + self.burn_in: (index + 1) * n]

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_closed", True):
            return
        self._closed = True
        if self.mode == "w":
            self._compact()
            if self._memmap is not None:
                self._memmap.flush()
                self._memmap = None
            if not getattr(self, "_allocated", False):
                return  # nothing was ever written: leave existing files and attributes alone
            self._flush_attributes()
        if self._h5 is not None:
            self._h5.close()
            self._h5 = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, value, traceback):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def combine_samples(samples_list: _Union[_List[Samples], _List[str]], output_filename=None,
                    cull_nan=True):
    """In-memory concatenation of sample collections (Samples.py:415-455)."""
    assert type(samples_list) == list, "Passed sample files/objects are not in list format."
    close_files = False
    if all(isinstance(n, str) for n in samples_list):
        close_files = True
        samples_list = [Samples(item) for item in samples_list]
    elif not all(isinstance(n, Samples) for n in samples_list):
        raise ValueError("Passed neither only strings to a sample files nor only sample "
                         "collections. Can't combine samples. ")
    if output_filename is not None:
        raise NotImplementedError
    out = _numpy.hstack([item.numpy for item in samples_list])
    if cull_nan:
        out = out[:, _numpy.logical_not(_numpy.isnan(_numpy.sum(out, axis=0)))]
    if close_files:
        for item in samples_list:
            item.close()
    return out
