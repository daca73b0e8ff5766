"""Host logic of the shared-memory staged CSR SpMM (no GPU): hmcb_debug_spmm_tables builds the strip
tables exactly as hmcb_finalize does, checks their format invariants and evaluates Y = A B by
walking them the way the kernel does.  Compared with scipy on awkward matrices, for every thread
mapping the library is built with, both nonzero encodings and strip limits that force the
builder's fallback."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

SHAPES = [(31, 8, 2), (27, 8, 2), (23, 12, 2), (20, 12, 2), (19, 16, 2), (4, 3, 1)]


@pytest.fixture(scope="module")
def lib():
    from hmclab_b200 import _build, _engine

    return _engine.load_library(_build.build())


def _tables(lib, A, B, shape, kb, emax, allow_compact=True):
    A = sp.csr_matrix(A)
    rows, cols = A.shape
    chains = B.shape[1]
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    Bc = np.ascontiguousarray(B, dtype=np.float64)
    Y = np.empty((rows, chains))
    info = (C.c_int64 * 4)()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))   # noqa: E731
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    status = lib.hmcb_debug_spmm_tables(rows, cols, data.size, ip(indptr), ip(indices), dp(data), *shape, kb, emax,
                                        int(allow_compact), chains, dp(Bc), dp(Y), info)
    assert status == 0, lib.hmcb_last_error().decode()
    return Y, dict(T=info[0], groups=info[1], compact=info[2], nbytes=info[3])


def _awkward(rows, cols, seed, f32):
    rng = np.random.default_rng(seed)
    A = sp.random(rows, cols, density=0.04, random_state=np.random.RandomState(seed), format="lil")
    A[3, :] = 0.0
    A[min(40, rows - 1):min(70, rows), :] = 0.0      # a run of empty rows
    A[1, :] = rng.normal(size=cols)                   # full row
    A[:, 5] = rng.normal(size=(rows, 1))              # full column
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    if f32:
        A.data = A.data.astype(np.float32).astype(np.float64)
    return A


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("f32", [True, False])
def test_tables_reproduce_the_product(lib, shape, f32):
    A = _awkward(700, 333, 2, f32)
    B = np.random.default_rng(1).normal(size=(333, 5))
    Y, info = _tables(lib, A, B, shape, kb=120, emax=640)
    assert info["compact"] == int(f32)
    assert info["T"] >= -(-333 // 120)
    np.testing.assert_allclose(Y, A @ B, rtol=0, atol=1e-12 * np.abs(A @ B).max())
    # the transposed product goes through the same builder (more columns than rows per chunk)
    Yt, _ = _tables(lib, A.T, np.random.default_rng(2).normal(size=(700, 3)), shape, kb=64, emax=300)
    np.testing.assert_allclose(Yt, A.T @ np.random.default_rng(2).normal(size=(700, 3)), rtol=0, atol=1e-11)


def test_compact_form_halves_the_tables_and_can_be_refused(lib):
    A = _awkward(500, 260, 4, True)
    B = np.random.default_rng(3).normal(size=(260, 4))
    Y8, i8 = _tables(lib, A, B, (31, 8, 2), 192, 896)
    Y16, i16 = _tables(lib, A, B, (31, 8, 2), 192, 896, allow_compact=False)
    assert i8["compact"] == 1 and i16["compact"] == 0 and i8["nbytes"] < 0.62 * i16["nbytes"]
    # same values; the strip count (and with it the order of summation) may differ
    np.testing.assert_allclose(Y8, Y16, rtol=0, atol=1e-12 * np.abs(Y16).max())


def test_tight_limits_force_more_narrower_strips(lib):
    A = _awkward(300, 200, 6, False)
    B = np.random.default_rng(5).normal(size=(200, 2))
    wide, iw = _tables(lib, A, B, (4, 3, 1), kb=50, emax=4096)
    tight, it = _tables(lib, A, B, (4, 3, 1), kb=50, emax=1)     # floor: one full column per group
    assert it["T"] > iw["T"] >= 4
    np.testing.assert_allclose(tight, A @ B, rtol=0, atol=1e-12)
    np.testing.assert_allclose(wide, tight, rtol=0, atol=1e-12)


def test_unsorted_rows_and_split_entries(lib):
    rng = np.random.default_rng(8)
    A = _awkward(150, 90, 7, True)
    ip, ix, dv = [0], [], []
    for i in range(A.shape[0]):
        cols, vals = list(A.indices[A.indptr[i]:A.indptr[i + 1]]), list(A.data[A.indptr[i]:A.indptr[i + 1]])
        if cols:
            cols.append(cols[0]); vals.append(vals[0] / 2); vals[0] /= 2
        perm = rng.permutation(len(cols))
        ix += [cols[p] for p in perm]; dv += [vals[p] for p in perm]
        ip.append(len(ix))
    raw = sp.csr_matrix((np.array(dv), np.array(ix, dtype=np.int32), np.array(ip, dtype=np.int32)), shape=A.shape)
    assert not raw.has_sorted_indices or True
    B = rng.normal(size=(90, 3))
    Y, _ = _tables(lib, raw, B, (20, 12, 2), 40, 640)
    np.testing.assert_allclose(Y, A @ B, rtol=0, atol=1e-12)


def test_bad_csr_is_refused(lib):
    info = (C.c_int64 * 4)()
    indptr = np.array([0, 2, 1], dtype=np.int32)     # not monotone
    indices = np.array([0, 1], dtype=np.int32)
    data = np.ones(2)
    B, Y = np.ones((2, 1)), np.ones((2, 1))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))   # noqa: E731
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    assert lib.hmcb_debug_spmm_tables(2, 2, 2, ip(indptr), ip(indices), dp(data), 4, 3, 1, 8, 64, 1, 1,
                                      dp(B), dp(Y), info) != 0
    assert len(lib.hmcb_last_error()) > 0


# ---- row-blocked tensor-core tables (csr_spmm_block_kernel) ---------------------------------

def _block_tables(lib, A, B, warps, gw=2, nb=1, cap16=7168, allow_compact=True):
    A = sp.csr_matrix(A)
    rows, cols = A.shape
    chains = B.shape[1]
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    Bc = np.ascontiguousarray(B, dtype=np.float64)
    Y = np.empty((rows, chains))
    info = (C.c_int64 * 8)()
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))   # noqa: E731
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    status = lib.hmcb_debug_spmm_block_tables(rows, cols, data.size, ip(indptr), ip(indices), dp(data), warps, gw, nb,
                                              cap16, int(allow_compact), chains, dp(Bc), dp(Y), info)
    assert status == 0, lib.hmcb_last_error().decode()
    return Y, dict(T=info[0], groups=info[1], blocks=info[2], nbytes=info[3], nnz=info[4], ktiles=info[5],
                   full_tiles=info[6], compact=info[7])


@pytest.mark.parametrize("shape", [(31, 4, 1), (31, 2, 2), (15, 8, 1), (3, 2, 1)])
@pytest.mark.parametrize("f32", [True, False])
def test_block_tables_reproduce_the_product(lib, shape, f32):
    A = _awkward(700, 333, 2, f32)
    B = np.random.default_rng(1).normal(size=(333, 5))
    Y, info = _block_tables(lib, A, B, *shape)
    assert info["nnz"] == A.nnz and 0 < info["blocks"] <= A.nnz and info["compact"] == int(f32)
    assert info["ktiles"] * 4 >= info["blocks"]
    np.testing.assert_allclose(Y, A @ B, rtol=0, atol=1e-12 * np.abs(A @ B).max())
    Bt = np.random.default_rng(2).normal(size=(700, 3))
    Yt, info_t = _block_tables(lib, A.T, Bt, *shape, cap16=1024)     # small stages: many strips
    assert info_t["T"] > 1
    np.testing.assert_allclose(Yt, A.T @ Bt, rtol=0, atol=1e-11)
    _, info_f = _block_tables(lib, A, B, *shape, allow_compact=False)
    assert info_f["compact"] == 0 and info_f["nbytes"] > info["nbytes"] * (1.25 if f32 else 0.99)


def test_block_tables_unsorted_rows_duplicates_and_empty_matrix_parts(lib):
    rng = np.random.default_rng(9)
    rows, cols = 90, 41
    dense = np.where(rng.uniform(size=(rows, cols)) < 0.2, rng.normal(size=(rows, cols)), 0.0)
    dense[10:30] = 0.0                                   # a run of empty rows
    A = sp.csr_matrix(dense)
    # split every entry into two slots and shuffle each row (scipy accepts both)
    indptr, indices, data = [0], [], []
    for i in range(rows):
        cols_i = A.indices[A.indptr[i]:A.indptr[i + 1]]
        vals_i = A.data[A.indptr[i]:A.indptr[i + 1]]
        ci = np.concatenate([cols_i, cols_i]); vi = np.concatenate([0.25 * vals_i, 0.75 * vals_i])
        perm = rng.permutation(ci.size)
        indices += list(ci[perm]); data += list(vi[perm]); indptr.append(len(indices))
    messy = sp.csr_matrix((np.array(data), np.array(indices), np.array(indptr)), shape=(rows, cols))
    B = rng.normal(size=(cols, 4))
    Y, info = _block_tables(lib, messy, B, 3, cap16=256)
    assert info["nnz"] == A.nnz
    np.testing.assert_allclose(Y, dense @ B, rtol=0, atol=1e-13)
    assert np.all(Y[10:30] == 0.0)


def test_row_clustering_finds_the_reuse_of_straight_rays(lib):
    """Rays of a tomography operator that run side by side cross the same cells: after clustering
    one gathered row of the operand feeds several rows of a group (the reason the blocked kernel
    exists); a matrix without such structure gives almost none and keeps the plain strip kernel."""
    from hmclab_b200.workloads import straight_ray_matrix

    G = straight_ray_matrix(40, 40, 6000, seed=3)
    B = np.random.default_rng(0).normal(size=(1600, 2))
    Y, info = _block_tables(lib, G, B, 31, 4, 1)
    np.testing.assert_allclose(Y, G @ B, rtol=0, atol=1e-11)
    assert info["nnz"] / info["blocks"] > 2.5
    assert info["full_tiles"] > 0.5 * info["ktiles"]     # most k-tiles carry 4 columns
    Bt = np.random.default_rng(1).normal(size=(6000, 2))
    Yt, info_t = _block_tables(lib, G.T, Bt, 31, 4, 1)
    np.testing.assert_allclose(Yt, G.T @ Bt, rtol=0, atol=1e-10)
    assert info_t["nnz"] / info_t["blocks"] > 2.0
    R = sp.random(600, 500, density=0.01, random_state=np.random.RandomState(1), format="csr")
    _, info_r = _block_tables(lib, R, np.ones((500, 1)), 31, 4, 1)
    assert info_r["nnz"] / info_r["blocks"] < 1.5
