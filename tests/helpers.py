"""Shared helpers for the parity tests."""
import os

import numpy as np

import cases  # tests/golden/cases.py

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    data = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    inp = {k[3:]: data[k] for k in data.files if k.startswith("in_")}
    ref = {k[4:]: data[k] for k in data.files if k.startswith("ref_")}
    return inp, ref


# Cases whose reference arithmetic contains a float32 BLAS product that the reference re-forms at every
# call (LinearMatrix.py:288, `Gt @ invcov` of the dense-covariance direct form): its rounding belongs to
# the sgemm kernel of the host that ran the reference, so on another CPU the stored outputs are
# reproduced to float32 level only.  The 1e-10 check for these cases is against the oracle evaluated on
# the SAME host (same numpy expression, same BLAS) -- see test_gpu_parity.py.
FLOAT32_BLAS_CASES = {"dense_fullcov_direct": 3e-5}


def tol_for(name, tight):
    return FLOAT32_BLAS_CASES.get(name, tight)


def build_mirror(name, inp):
    import hmclab_b200

    return cases.build(name, inp, hmclab_b200)


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over finite entries; non-finite entries must agree in
    class (nan vs nan, +inf vs +inf)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "finite/non-finite pattern differs"
    if not fin.all():
        assert np.array_equal(np.isnan(a), np.isnan(b)), "nan pattern differs"
        assert np.array_equal(a[np.isinf(b)], b[np.isinf(b)]), "inf sign differs"
    if not fin.any():
        return 0.0
    scale = max(np.max(np.abs(b[fin])), 1e-300)
    return float(np.max(np.abs(a[fin] - b[fin])) / scale)
