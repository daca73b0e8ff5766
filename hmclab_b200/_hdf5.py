"""Minimal native HDF5 writer/reader for the samples file (no libhdf5 / h5py needed).

The reference stores samples in an HDF5 dataset ``"samples"`` of shape ``(d+1, n)`` float64 with
scalar attributes on the dataset (hmclab/Samples.py:127-144, 305-322, 369-385).  h5py is not part
of this image, so this module writes that file directly in the classic on-disk format every
libhdf5 release reads (HDF5 File Format Specification, version 0 superblock):

    superblock v0 -> root group (object header v1 + symbol-table message -> B-tree v1 node ->
    symbol-table node "SNOD" + local heap with the link names) -> dataset object header v1 with
    dataspace, datatype, fill-value, contiguous-layout and attribute messages

The raw data of the dataset is one contiguous block right after the superblock, so a writer can
``numpy.memmap`` it and stream blocks of samples into place; the metadata is (re)written behind
the data by :meth:`Writer.commit` and the superblock patched to point at it.

The reader understands what the writer produces plus the same structures as libhdf5 writes them
(checked in tests against a MATLAB v7.3 file written by libhdf5 that ships with SciPy): version
0/1 superblocks, version 1 object headers with continuation blocks, group B-trees, contiguous and
compact layouts, fixed-point / floating-point / fixed-length string types, version 1-3 attribute
messages.  Files written with ``libver="latest"`` (version 2 object headers, chunked layout
version 4) need h5py.
"""
from __future__ import annotations

import os
import struct
from typing import Any, Dict, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
DATA_OFFSET = 2048            # raw data starts here (superblock is 96 bytes)
_GROUP_LEAF_K, _GROUP_INTERNAL_K = 4, 16
_HEAP_FREE_NULL = 1           # libhdf5's H5HL_FREE_NULL: end of a local heap's free list


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------ encode ---
def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        size = dt.itemsize
        exp_bits, man_bits, bias = (11, 52, 1023) if size == 8 else (8, 23, 127)
        # class 1 (floating point) version 1; little endian, mantissa normalisation "implied msb",
        # sign bit position in the second flag byte
        head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)
        return head + struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu":
        flags = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, flags, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        # class 3 (string) version 1; null padded (1), character set UTF-8 (1 << 4)
        return struct.pack("<BBBBI", 0x13, 0x11, 0, 0, dt.itemsize)
    raise TypeError(f"unsupported attribute/dataset type {dt}")


def _dataspace_message(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


_MAX_MESSAGE = 0xFFF0       # a version 1 header message stores its size in 16 bits


def _as_array(value: Any) -> np.ndarray:
    """The array an attribute value is stored as (little-endian f8 / i8 / fixed-length string)."""
    if isinstance(value, (str, bytes)):
        raw = value.encode("utf-8") if isinstance(value, str) else value
        return np.array(raw + b"\0", dtype=f"S{len(raw) + 1}")
    if isinstance(value, (bool, np.bool_)):
        return np.array(int(value), dtype="<i8")
    arr = np.asarray(value)
    if arr.dtype.kind == "f":
        return arr.astype("<f8")
    if arr.dtype.kind in "iu":
        return arr.astype("<i8")
    if arr.dtype.kind == "b":
        return arr.astype("<i8")
    if arr.dtype.kind == "U":
        enc = np.char.encode(arr, "utf-8")
        return enc.astype(f"S{enc.dtype.itemsize + 1}")
    if arr.dtype.kind == "S":
        return arr
    raw = str(value).encode("utf-8")
    return np.array(raw + b"\0", dtype=f"S{len(raw) + 1}")


def _attribute_body(name: str, value: Any) -> bytes:
    arr = _as_array(value)
    nm = name.encode("utf-8") + b"\0"
    dtm, dsm = _dtype_message(arr.dtype), _dataspace_message(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(dsm))
    return body + _pad8(nm) + _pad8(dtm) + _pad8(dsm) + arr.tobytes()


def _attribute_message(name: str, value: Any) -> bytes:
    return _message(0x000C, _attribute_body(name, value))


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class Writer:
    """One contiguous float64 dataset, preallocated, writable through ``self.data``."""

    def __init__(self, filename: str, shape: Tuple[int, int], name: str = "samples", overwrite: bool = False):
        if os.path.exists(filename) and not overwrite:
            raise FileExistsError(filename)
        self.filename, self.name = filename, name
        self.shape = (int(shape[0]), int(shape[1]))
        self._nbytes_allocated = 8 * self.shape[0] * self.shape[1]
        with open(filename, "wb") as f:
            f.truncate(DATA_OFFSET + max(self._nbytes_allocated, 8))
        self.data = np.memmap(filename, dtype="<f8", mode="r+", offset=DATA_OFFSET,
                              shape=self.shape) if self._nbytes_allocated else np.zeros(self.shape)
        self.commit({})

    def resize_columns(self, keep_columns: int, columns_per_group: int):
        """Keep the first ``keep_columns`` of every group of ``columns_per_group`` columns
        (a run that stopped early), compacting the rows in place."""
        rows, total = self.shape
        groups = total // columns_per_group if columns_per_group else 0
        old = np.array(self.data.reshape(rows, groups, columns_per_group)[:, :, :keep_columns])
        del self.data
        self.shape = (rows, groups * keep_columns)
        flat = np.memmap(self.filename, dtype="<f8", mode="r+", offset=DATA_OFFSET,
                         shape=(max(rows * groups * keep_columns, 1),))
        flat[: rows * groups * keep_columns] = old.reshape(-1)
        flat.flush()
        del flat
        self._nbytes_allocated = 8 * rows * groups * keep_columns   # commit() truncates the file behind it
        self.data = np.memmap(self.filename, dtype="<f8", mode="r+", offset=DATA_OFFSET,
                              shape=self.shape) if self.shape[0] * self.shape[1] else np.zeros(self.shape)

    def commit(self, attributes: Dict[str, Any]):
        """(Re)write the metadata behind the data block and point the superblock at it.

        A version 1 object-header message carries a 16-bit size, so an attribute above 64 KB
        cannot live in the dataset's header (libhdf5 has the same limit for compact attribute
        storage).  Such attributes -- the per-proposal ``stepsizes`` / ``acceptance_rates`` of an
        autotuned run -- are stored as root-level datasets ``<name>.<attribute>`` next to the
        samples; :func:`open_dataset` folds them back into the attribute dictionary."""
        if isinstance(self.data, np.memmap):
            self.data.flush()
        meta0 = DATA_OFFSET + max(self._nbytes_allocated, 8)
        meta0 += -meta0 % 8
        small, big = {}, {}
        for k, v in attributes.items():
            msg = _attribute_body(k, v)
            if len(msg) + 8 > _MAX_MESSAGE:
                big[f"{self.name}.{k}"] = _as_array(v)
            else:
                small[k] = msg
        if len(big) > 2 * _GROUP_LEAF_K - 1:
            raise ValueError("too many oversize attributes for one symbol-table node")
        names = sorted([self.name] + list(big), key=lambda n: n.encode("utf-8"))
        # local heap data segment: "" at 0 (the root's own name), the link names, one free block
        heap_names, name_off, off = b"", {}, 8
        for n in names:
            enc = _pad8(n.encode("utf-8") + b"\0")
            name_off[n] = off
            heap_names += enc
            off += len(enc)
        heap_data_size = off + 16
        free_off = off
        heap_data = b"\0" * 8 + heap_names + struct.pack("<QQ", _HEAP_FREE_NULL, 16)
        btree_size = 24 + (2 * _GROUP_INTERNAL_K + 1) * 8 + 2 * _GROUP_INTERNAL_K * 8
        snod_size = 8 + 2 * _GROUP_LEAF_K * 40
        a_root = meta0
        a_heap = a_root + 16 + 24 + 8                      # root header: prefix + stab message + NIL
        a_heap_data = a_heap + 32
        a_btree = a_heap_data + heap_data_size
        a_snod = a_btree + btree_size
        a_dset = a_snod + snod_size
        root = _object_header([_message(0x0011, struct.pack("<QQ", a_btree, a_heap), flags=1),
                               _message(0x0000, b"")])
        heap = b"HEAP" + struct.pack("<B3xQQQ", 0, heap_data_size, free_off, a_heap_data)

        def header(shape, dt, data_addr, nbytes, attr_msgs=()):
            msgs = [
                _message(0x0001, _dataspace_message(shape)),
                _message(0x0003, _dtype_message(dt), flags=1),
                # fill value message exactly as libhdf5 writes its default (version 1, allocation
                # time late, write time "if set", defined, size 0)
                _message(0x0005, struct.pack("<BBBBI", 1, 2, 2, 1, 0)),
                _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, nbytes)),
            ]
            return _object_header(msgs + [_message(0x000C, m) for m in attr_msgs])

        main = header(self.shape, np.dtype("<f8"), DATA_OFFSET, 8 * self.shape[0] * self.shape[1],
                      small.values())
        # oversize attributes: header + raw data, one after the other behind the main header
        addr = {self.name: a_dset}
        cursor = a_dset + len(main)
        extras = b""
        for n in names:
            if n == self.name:
                continue
            arr = big[n]
            raw = arr.tobytes()
            probe = header(arr.shape, arr.dtype, 0, len(raw))
            addr[n] = cursor
            extras += header(arr.shape, arr.dtype, cursor + len(probe), len(raw)) + _pad8(raw)
            cursor += len(probe) + len(_pad8(raw))
        btree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF)
        btree += struct.pack("<QQQ", 0, a_snod, name_off[names[-1]])
        btree += b"\0" * (btree_size - len(btree))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:                                     # entries sorted by link name
            snod += struct.pack("<QQII16x", name_off[n], addr[n], 0, 0)
        snod += b"\0" * (snod_size - len(snod))
        blob = root + heap + heap_data + btree + snod + main + extras
        assert len(root) == a_heap - a_root and a_dset - meta0 == len(blob) - len(main) - len(extras)
        eof = meta0 + len(blob)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _GROUP_LEAF_K, _GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)
        with open(self.filename, "r+b") as f:
            f.seek(meta0)
            f.write(blob)
            f.truncate(eof)
            f.seek(0)
            f.write(sb)

    def close(self, attributes: Dict[str, Any]):
        self.commit(attributes)
        if isinstance(self.data, np.memmap):
            del self.data
        self.data = None


# ------------------------------------------------------------------------------ decode ---
class _File:
    def __init__(self, filename):
        self.filename = filename
        with open(filename, "rb") as f:
            self.buf = f.read(1 << 20)          # metadata of small files; larger reads go to disk
        self.size = os.path.getsize(filename)
        off = 0
        while off < self.size and self._read(off, 8) != SIGNATURE:
            off = 512 if off == 0 else off * 2
        if off >= self.size:
            raise ValueError(f"{filename}: not an HDF5 file")
        sb = self._read(off, 128)
        version = sb[8]
        if version not in (0, 1):
            raise NotImplementedError(
                f"{filename}: HDF5 superblock version {version} (written with libver='latest'?) needs h5py")
        if sb[13] != 8 or sb[14] != 8:
            raise NotImplementedError("only 8-byte offsets and lengths are supported")
        p = 24 if version == 0 else 28
        self.base = struct.unpack_from("<Q", sb, p)[0]
        self.root_header = struct.unpack_from("<Q", sb, p + 32 + 8)[0]

    def _read(self, addr, n):
        if addr + n <= len(self.buf):
            return self.buf[addr: addr + n]
        with open(self.filename, "rb") as f:
            f.seek(addr)
            return f.read(n)

    def read(self, addr, n):
        return self._read(self.base + addr, n)

    # object header v1 -> list of (type, data)
    def messages(self, addr):
        head = self.read(addr, 16)
        if head[:4] == b"OHDR":
            raise NotImplementedError("version 2 object headers (libver='latest') need h5py")
        version, _, nmsgs, _, size = struct.unpack_from("<BBHII", head)
        if version != 1:
            raise ValueError("bad object header")
        out, blocks = [], [(addr + 16, size)]
        while blocks and len(out) < nmsgs:
            a, n = blocks.pop(0)
            chunk, p = self.read(a, n), 0
            while p + 8 <= n and len(out) < nmsgs:
                mtype, msize, _flags = struct.unpack_from("<HHB", chunk, p)
                data = chunk[p + 8: p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:
                    blocks.append(struct.unpack("<QQ", data[:16]))
                out.append((mtype, data))
        return out

    def links(self, group_addr):
        stab = [d for t, d in self.messages(group_addr) if t == 0x0011]
        if not stab:
            raise NotImplementedError("groups without a symbol table (new-style links) need h5py")
        btree, heap = struct.unpack("<QQ", stab[0][:16])
        h = self.read(heap, 32)
        assert h[:4] == b"HEAP"
        hsize, _free, haddr = struct.unpack_from("<QQQ", h, 8)
        names = self.read(haddr, hsize)
        out = {}

        def walk(node):
            t = self.read(node, 24)
            assert t[:4] == b"TREE" and t[4] == 0
            level, used = t[5], struct.unpack_from("<H", t, 6)[0]
            body = self.read(node + 24, (2 * used + 1) * 8)
            for i in range(used):
                child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
                if level:
                    walk(child)
                    continue
                s = self.read(child, 8)
                assert s[:4] == b"SNOD"
                nsym = struct.unpack_from("<H", s, 6)[0]
                ents = self.read(child + 8, 40 * nsym)
                for j in range(nsym):
                    noff, oaddr = struct.unpack_from("<QQ", ents, 40 * j)
                    out[names[noff: names.index(b"\0", noff)].decode("utf-8")] = oaddr

        walk(btree)
        return out

    @staticmethod
    def _dtype(data):
        cls, version = data[0] & 0x0F, data[0] >> 4
        b0, b1 = data[1], data[2]
        size = struct.unpack_from("<I", data, 4)[0]
        order = ">" if (b0 & 1) else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if b0 & 0x08 else 'u'}{size}"), 8 + 4
        if cls == 1:
            return np.dtype(f"{order}f{size}"), 8 + 12
        if cls == 3:
            return np.dtype(f"S{size}"), 8
        raise NotImplementedError(f"HDF5 datatype class {cls} (version {version}) is not supported")

    @staticmethod
    def _dataspace(data):
        version, rank, flags = data[0], data[1], data[2]
        p = 8 if version == 1 else 4
        return tuple(struct.unpack_from("<Q", data, p + 8 * i)[0] for i in range(rank))

    def attribute(self, data):
        version = data[0]
        nlen, tlen, slen = struct.unpack_from("<HHH", data, 2)
        p = 8 + (1 if version == 3 else 0)
        al = (lambda n: n + (-n % 8)) if version == 1 else (lambda n: n)
        name = data[p: p + nlen].split(b"\0")[0].decode("utf-8")
        p += al(nlen)
        dt, _ = self._dtype(data[p: p + tlen])
        p += al(tlen)
        shape = self._dataspace(data[p: p + slen])
        p += al(slen)
        count = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(data[p: p + count * dt.itemsize], dtype=dt).reshape(shape)
        if dt.kind == "S":
            arr = np.char.decode(np.char.rstrip(arr, b"\0"), "utf-8") if arr.shape else \
                arr.tobytes().split(b"\0")[0].decode("utf-8")
            return name, arr
        return name, (arr[()] if not shape else arr.copy())

    def dataset(self, addr):
        """-> (dtype, shape, ('contiguous', file offset) | ('compact', bytes), attributes)."""
        dt = shape = where = None
        attrs = {}
        for mtype, data in self.messages(addr):
            if mtype == 0x0001:
                shape = self._dataspace(data)
            elif mtype == 0x0003:
                dt, _ = self._dtype(data)
            elif mtype == 0x0008:
                version, cls = data[0], data[1]
                if version in (1, 2):     # version, rank + 1, class, 5 reserved, address, 4-byte dims
                    if data[2] != 1:
                        raise NotImplementedError("only contiguous version 1/2 layouts are supported")
                    where = ("contiguous", self.base + struct.unpack_from("<Q", data, 8)[0])
                    continue
                if version != 3:
                    raise NotImplementedError(f"data layout message version {version} is not supported")
                if cls == 1:
                    where = ("contiguous", self.base + struct.unpack_from("<Q", data, 2)[0])
                elif cls == 0:
                    n = struct.unpack_from("<H", data, 2)[0]
                    where = ("compact", data[4: 4 + n])
                else:
                    raise NotImplementedError("chunked datasets need h5py")
            elif mtype == 0x000C:
                try:
                    k, v = self.attribute(data)
                    attrs[k] = v
                except NotImplementedError:
                    pass
        if dt is None or shape is None or where is None:
            raise ValueError("not a dataset")
        return dt, shape, where, attrs


def open_dataset(filename: str, name: str = "samples"):
    """-> (array (numpy.memmap for contiguous data), attributes dict) of a root-level dataset."""
    f = _File(filename)
    links = f.links(f.root_header)
    if name not in links:
        raise KeyError(f"{filename}: no dataset `{name}` (found {sorted(links)})")
    def load(link, mapped):
        dt, shape, where, attrs = f.dataset(links[link])
        count = int(np.prod(shape)) if shape else 1
        if where[0] == "compact":
            arr = np.frombuffer(where[1][: count * dt.itemsize], dtype=dt).reshape(shape).copy()
        elif count == 0 or where[1] == UNDEF:
            arr = np.zeros(shape, dtype=dt)
        elif mapped:
            arr = np.memmap(filename, dtype=dt, mode="r", offset=where[1], shape=shape)
        else:
            arr = np.fromfile(filename, dtype=dt, count=count, offset=where[1]).reshape(shape)
        return arr, attrs

    arr, attrs = load(name, True)
    for link in links:          # oversize attributes stored as companion datasets (Writer.commit)
        if link.startswith(name + "."):
            attrs[link[len(name) + 1:]] = load(link, False)[0]
    return arr, attrs
