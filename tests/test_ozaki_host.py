"""Host side of the Ozaki slicing (csrc/launch_ozaki.cu: oz_slice_rows_host, the digits hmcb_finalize uploads
for the model matrix): balanced radix-256 digits reproduce the matrix, stay inside int8, and a float32
matrix (the reference's default rounding, LinearMatrix.py:148-153) needs 5 digit planes."""
import ctypes as C

import numpy as np

from hmclab_b200._engine import load_library


def _slice(A, S):
    lib = load_library()
    A = np.ascontiguousarray(A, dtype=np.float64)
    rows, cols = A.shape
    sl = np.zeros((S, rows, cols), dtype=np.int8)
    ea = np.zeros(rows, dtype=np.int32)
    err = lib.hmcb_debug_oz_slice_rows(A.ctypes.data_as(C.POINTER(C.c_double)), rows, cols, S,
                                       sl.ctypes.data_as(C.c_void_p), ea.ctypes.data_as(C.c_void_p))
    return sl, ea, err


def _rebuild(sl, ea):
    S = sl.shape[0]
    X = np.zeros(sl.shape[1:], dtype=object)
    for s in range(S):
        X = X * 256 + sl[s].astype(object)          # exact integers
    return np.array([[float(x) for x in row] for row in X]) * np.ldexp(1.0, ea - 8 * S)[:, None]


def test_digits_rebuild_the_matrix_within_half_a_unit_of_the_last_digit():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(37, 53)) * np.exp(rng.normal(size=(37, 1)) * 8.0)
    A[3] = 0.0
    A[5, :] = [1.0] * 53                      # maximum exactly a power of two
    A[6, 0] = np.nextafter(2.0, 0.0)         # maximum just below a power of two: needs the extra exponent step
    for S in (1, 3, 6, 7):
        sl, ea, err = _slice(A, S)
        assert sl.min() >= -128 and sl.max() <= 127
        back = _rebuild(sl, ea)
        unit = np.ldexp(1.0, ea - 8 * S)[:, None]
        assert np.all(np.abs(back - A) <= 0.5 * unit)
        rowsum = np.abs(A).sum(axis=1)
        worst = max(np.abs(back - A).sum(axis=1)[rowsum > 0] / rowsum[rowsum > 0])
        assert err == worst or abs(err - worst) <= 1e-3 * worst
        # scaling leaves the largest entry below 0.494 of the range and above an eighth of it
        top = np.abs(A).max(axis=1)
        nz = top > 0
        assert np.all(top[nz] * np.ldexp(1.0, -ea[nz]) < 0.494) and np.all(top[nz] * np.ldexp(1.0, -ea[nz]) >= 0.123)
    assert np.all(_slice(A, 4)[0][:, 3] == 0) and _slice(A, 4)[1][3] == 0


def test_float32_matrix_needs_five_planes_and_small_integers_are_exact():
    rng = np.random.default_rng(1)
    G = (rng.normal(size=(64, 2000)) / 100.0).astype(np.float32).astype(np.float64)
    assert _slice(G, 4)[2] > 2.0 ** -45 and _slice(G, 5)[2] < 2.0 ** -45
    ints = rng.integers(-100, 101, size=(8, 16)).astype(np.float64)
    sl, ea, err = _slice(ints, 2)
    assert err == 0.0 and np.array_equal(_rebuild(sl, ea), ints)
    G64 = rng.normal(size=(16, 500))
    assert _slice(G64, 5)[2] > 2.0 ** -45 and _slice(G64, 6)[2] < 2.0 ** -45
