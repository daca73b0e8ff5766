"""CPU replica of the engine's on-device random streams (TEST INFRASTRUCTURE ONLY).

Mirrors hmclab_b200/csrc/common.cuh: Philox4x32-10 keyed by the 64-bit seed with counter
(chain, proposal, pair, stream); 53-bit uniforms; Box-Muller normals.  The reference draws
from numpy's Generator instead (hmclab/Samplers.py:316-319), so there is nothing in the
reference to compare these streams with: this replica pins their *definition* (a change of
the device code that alters a stream fails tests/test_gpu_rng.py) and checks the accuracy of
the device's specialised log / sqrt / sincos against numpy's libm.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_NORMAL, STREAM_UNIFORM = 0, 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u53(hi, lo):
    w = (hi << np.uint64(32)) | lo
    return (w >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def normals(seed, chains, proposal, dims, chain_offset=0):
    """z [chains, dims] of one proposal (coordinates 2m, 2m+1 come from pair m)."""
    pairs = (dims + 1) // 2
    c = (np.arange(chains, dtype=np.uint64) + np.uint64(chain_offset))[:, None]
    m = np.arange(pairs, dtype=np.uint64)[None, :]
    x, y, z, w = philox4x32_10(c, np.uint64(proposal), m, np.uint64(STREAM_NORMAL),
                               seed & 0xFFFFFFFF, seed >> 32)
    u1 = 1.0 - u53(x, y)
    u2 = u53(z, w)
    rad = np.sqrt(-2.0 * np.log(u1))
    out = np.empty((chains, 2 * pairs))
    out[:, 0::2] = rad * np.cos(2.0 * np.pi * u2)
    out[:, 1::2] = rad * np.sin(2.0 * np.pi * u2)
    return out[:, :dims]


def uniforms(seed, chains, proposal, chain_offset=0):
    """(step-size factor in [0.5, 1.5), acceptance uniform in [0, 1)) per chain."""
    c = np.arange(chains, dtype=np.uint64) + np.uint64(chain_offset)
    x, y, z, w = philox4x32_10(c, np.uint64(proposal), np.uint64(0), np.uint64(STREAM_UNIFORM),
                               seed & 0xFFFFFFFF, seed >> 32)
    return 0.5 + u53(x, y), u53(z, w)
