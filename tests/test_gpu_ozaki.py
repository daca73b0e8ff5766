"""The tcgen05 building block of the Ozaki-sliced dense products (csrc/ozaki.cuh): exact int8 slice
products accumulated per order in tensor memory, against integer arithmetic done with torch."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _i8_gemm(A, B, orders):
    import torch

    from hmclab_b200._engine import load_library

    lib = load_library()
    SA, M, K = A.shape
    SB, N, _ = B.shape
    out = torch.full((orders, M, N), -7, dtype=torch.int32, device="cuda")
    rc = lib.hmcb_debug_i8_gemm(torch.cuda.current_device(), M, N, K, SA, SB, orders, A.data_ptr(), B.data_ptr(),
                                out.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K,SA,SB,orders", [
    (128, 256, 128, 1, 1, 1), (256, 512, 384, 2, 3, 4), (384, 384, 1152, 3, 2, 4),
    (256, 1280, 640, 5, 6, 6),      # the default scheme: digits 5 x 6, orders 0..5 in three groups of two
    (128, 2304, 256, 6, 6, 6), (256, 256, 2048, 7, 7, 7), (128, 128, 256, 5, 6, 3), (2560, 256, 128, 4, 7, 5)])
@pytest.mark.parametrize("pair", ["1", "0"])
def test_i8_slice_products_are_exact(monkeypatch, pair, M, N, K, SA, SB, orders):
    """Every order plane equals the integer sum of its slice products, over the full int8 range; the
    grouped kernel (two orders per CTA sharing operand tiles) with odd and even order counts, column
    panels that do not fill (N / 256 not a multiple of 8) and a half-empty last column tile.  Both
    launch modes: CTA pairs (tcgen05 cta_group::2, the default; an odd number of row tiles leaves the last
    pair's second CTA without rows) and one CTA per tile (HMCB_OZAKI_PAIR=0)."""
    import torch

    monkeypatch.setenv("HMCB_OZAKI_PAIR", pair)

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randint(-128, 128, (SA, M, K), generator=g, device="cuda", dtype=torch.int8)
    B = torch.randint(-128, 128, (SB, N, K), generator=g, device="cuda", dtype=torch.int8)
    got = _i8_gemm(A, B, orders)
    ref = torch.zeros(orders, M, N, dtype=torch.float64, device="cuda")
    for s in range(SA):
        for t in range(SB):
            if s + t < orders:
                ref[s + t] += A[s].double() @ B[t].double().T
    assert torch.equal(got.double(), ref)


@pytest.mark.parametrize("crt", ["0", "1"])
@pytest.mark.parametrize("name", ["dense_direct_4s", "dense_direct_f64", "dense_full_mass_3s"])
def test_ozaki_path_reproduces_the_golden_trajectories(monkeypatch, name, crt):
    """The dense direct products on tcgen05 (int8 slices, exact int32 accumulation, fp64 recombination)
    forced on for the small golden cases: same 1e-10 / identical-decision bar as the DMMA path.  With
    HMCB_OZAKI_CRT=1 the second product G^T r runs as 13 modular products + Chinese-remainder reconstruction."""
    import torch

    import cases
    from helpers import build_mirror, load_golden, rel_err
    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten

    monkeypatch.setenv("HMCB_OZAKI", "1")
    monkeypatch.setenv("HMCB_OZAKI_CRT", crt)
    inp, ref = load_golden(name)
    s = cases.SETTINGS[name]
    K, C_, d = inp["z"].shape
    post, mass = build_mirror(name, inp)
    eng = Engine(flatten(describe(post)), describe_mass(mass), C_, integrator=s["integrator"], amount_of_steps=s["steps"])
    assert (eng.tcgen05_slice_pairs in (20 + 13, 21 + 13)) == (crt == "1") and eng.tcgen05_slice_pairs >= 33
    G = eng.grads_per_proposal
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()   # noqa: E731
    q = dev(inp["q0"])
    x = eng.misfit(q)
    out = dict(out_accept=torch.zeros(K, C_, dtype=torch.uint8, device="cuda"),
               out_h0=torch.zeros(K, C_, dtype=torch.float64, device="cuda"),
               out_h1=torch.zeros(K, C_, dtype=torch.float64, device="cuda"),
               out_q_prop=torch.zeros(K, C_, d, dtype=torch.float64, device="cuda"),
               out_p_prop=torch.zeros(K, C_, d, dtype=torch.float64, device="cuda"),
               trace_q=torch.zeros(K, G, C_, d, dtype=torch.float64, device="cuda"),
               trace_g=torch.zeros(K, G, C_, d, dtype=torch.float64, device="cuda"))
    eng.run_block(q, x, K, stepsize=s["stepsize"], randomize_stepsize=s["randomize"], z=dev(inp["z"]),
                  u_step=dev(inp["u_step"]), u_accept=dev(inp["u_acc"]), **out)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    assert np.array_equal(got["out_accept"].astype(bool), ref["accept"])
    for key, rk in (("trace_q", "trace_q"), ("trace_g", "trace_g"), ("out_q_prop", "q_prop"),
                    ("out_p_prop", "p_prop"), ("out_h0", "H0"), ("out_h1", "H1")):
        assert rel_err(got[key], ref[rk]) < 1e-10, key
    # the gradient entry point goes through the same path; a chain with a non-finite coordinate comes back NaN
    qq = dev(inp["q0"])
    g0 = eng.gradient(qq).cpu().numpy()
    qq[0, 0] = float("inf")
    g1 = eng.gradient(qq).cpu().numpy()
    assert np.all(np.isnan(g1[0]) | np.isinf(g1[0])) and np.array_equal(g1[1:], g0[1:])


def test_ozaki_gradient_follows_chain_scales_over_many_magnitudes():
    """Per-chain scaling: chains whose coordinates differ by hundreds of binades (and an all-zero chain)
    share one batch; each chain's gradient agrees with numpy to fp64-BLAS accuracy relative to its own
    natural scale sum_k |G^T|_jk |r_k|."""
    import os

    import torch

    import hmclab_b200.Distributions as D
    import hmclab_b200.MassMatrices as Mm
    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten

    os.environ["HMCB_OZAKI"] = "1"
    try:
        rng = np.random.default_rng(5)
        N, d, C_ = 300, 200, 64
        G = rng.normal(size=(N, d)) * np.exp(rng.normal(size=(N, 1)) * 3.0)
        dat = rng.normal(size=(N, 1))
        var = rng.uniform(0.5, 2.0, size=(N, 1))
        lik = D.LinearMatrix(G, dat, var, premultiplication=False)
        eng = Engine(flatten(describe(lik)), describe_mass(Mm.Unit(d)), C_, integrator="lf", amount_of_steps=1)
        assert eng.tcgen05_slice_pairs > 0
        q = rng.normal(size=(C_, d)) * np.exp2(rng.integers(-300, 300, size=(C_, 1)).astype(np.float64))
        q[7] = 0.0
        got = eng.gradient(torch.as_tensor(q).cuda()).cpu().numpy()
    finally:
        os.environ.pop("HMCB_OZAKI", None)
    inner = lik.Distribution
    G32, Gt = np.asarray(inner.G, dtype=np.float64), np.asarray(inner.Gt, dtype=np.float64)
    d32, v32 = np.asarray(inner.d, dtype=np.float64), np.asarray(inner.data_variance, dtype=np.float64)
    r = (G32 @ q.T - d32) / v32
    ref = (Gt @ r).T
    scale = (np.abs(Gt) @ np.abs(r)).T + (np.abs(Gt) @ ((np.abs(G32) @ np.abs(q.T)) / v32)).T
    assert np.all(np.abs(got - ref) <= 1e-12 * scale)


@pytest.mark.parametrize("N,kblocks,SA,SB,orders", [(256, (1,), 1, 1, 1), (384, (2, 1, 3), 3, 4, 5),
                                                     (1280, (3, 3, 2, 1), 5, 6, 6)])
def test_gathered_slice_products_are_exact(N, kblocks, SA, SB, orders):
    """The block-sparse building block (csrc/ozaki_sparse.cuh): every bundle's dense int8 tile times the B rows
    its list names, gathered by the producer warps into the swizzled operand tile -- against integer
    arithmetic; bundles of different lengths, repeated list entries, a half-empty last column tile."""
    import torch

    from hmclab_b200._engine import load_library

    lib = load_library()
    g = torch.Generator(device="cuda").manual_seed(N + sum(kblocks))
    nb, rows_b = len(kblocks), 700
    Ktot = 128 * sum(kblocks)
    koff = np.concatenate([[0], np.cumsum(kblocks)[:-1]]) * 128
    bundles = torch.as_tensor(np.stack([koff, kblocks], axis=1).astype(np.int32)).cuda().contiguous()
    lst = torch.randint(0, rows_b, (Ktot,), generator=g, device="cuda", dtype=torch.int32)
    A = torch.randint(-128, 128, (SA, 128, Ktot), generator=g, device="cuda", dtype=torch.int8)
    B = torch.randint(-128, 128, (SB, rows_b, N), generator=g, device="cuda", dtype=torch.int8)
    # the digit planes hold the chains of every 128-chain block permuted: chain h + 8 m sits at byte 16 h + m
    n = np.arange(N)
    pos = torch.as_tensor((n & ~127) + 16 * (n & 7) + ((n & 127) >> 3)).cuda()
    planes = torch.empty_like(B)
    planes[:, :, pos] = B
    out = torch.full((orders, nb * 128, N), -7, dtype=torch.int32, device="cuda")
    rc = lib.hmcb_debug_i8_gather_gemm(torch.cuda.current_device(), nb, N, Ktot, rows_b, SA, SB, orders, A.data_ptr(),
                                       bundles.data_ptr(), lst.data_ptr(), planes.data_ptr(), out.data_ptr(),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    torch.cuda.synchronize()
    ref = torch.zeros(orders, nb * 128, N, dtype=torch.float64, device="cuda")
    for b in range(nb):
        sl = slice(int(koff[b]), int(koff[b]) + 128 * kblocks[b])
        rows = lst[sl].long()
        for s in range(SA):
            for t in range(SB):
                if s + t < orders:
                    ref[s + t, b * 128:(b + 1) * 128] += A[s][:, sl].double() @ B[t][rows].double()
    assert torch.equal(out.double(), ref)


@pytest.mark.parametrize("premult,crt", [(False, "0"), (False, "1"), (True, "0")])
def test_ozaki_path_at_config3_shape_vs_oracle(monkeypatch, premult, crt):
    """A scaled config-3 problem (333 parameters x 700 data, 130 chains, 4-stage) with the tcgen05 path
    forced on, both forms of the dense LinearMatrix: the direct products G q / G^T r and the premultiplied
    GtG q (an fp64-valued operator: six digits), checked against the oracle like the full-size tests."""
    from hmclab_b200 import workloads
    from test_gpu_fullsize import _compare_with_oracle

    monkeypatch.setenv("HMCB_OZAKI", "1")
    monkeypatch.setenv("HMCB_OZAKI_CRT", crt)
    w = workloads.dense_large(dims=333, data=700, chains=130, premultiplication=premult)
    eng, _ = _compare_with_oracle(w, K=2)
    assert eng.path == "staged" and eng.tcgen05_slice_pairs > 0
    assert (eng.tcgen05_slice_pairs <= 21) == premult       # one product premultiplied, two in the direct form
    assert (eng.tcgen05_slice_pairs == 20 + 13) == (crt == "1")


@pytest.mark.parametrize("pair", ["1", "0"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (384, 640, 1152), (256, 256, 10112)])
def test_modular_products_rebuild_the_exact_integer_product(monkeypatch, pair, M, N, K):
    """The modular variant (13 int8 products modulo coprime moduli + Chinese-remainder reconstruction): for
    integer operands the result is the integer product to fp64 rounding; for real operands it is within 2^-42
    of sum |a| |x| (operands cut at 44 bits below the row / chain maximum)."""
    import torch

    from hmclab_b200._engine import load_library

    monkeypatch.setenv("HMCB_OZAKI_PAIR", pair)
    lib = load_library()
    rng = np.random.default_rng(M + N + K)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def product(A, X):
        Xd = torch.as_tensor(X).cuda().contiguous()
        Y = torch.full((M, N), np.nan, dtype=torch.float64, device="cuda")
        rc = lib.hmcb_debug_crt_product(torch.cuda.current_device(), M, N, K,
                                        np.ascontiguousarray(A).ctypes.data_as(C.POINTER(C.c_double)), Xd.data_ptr(),
                                        Y.data_ptr(), st)
        assert rc == 0, rc
        return Y.cpu().numpy()

    Ai = rng.integers(-2**15, 2**15, size=(M, K)).astype(np.float64)
    Xi = rng.integers(-2**15, 2**15, size=(K, N)).astype(np.float64)
    Ai[3] = 0.0                       # an all-zero row and an all-zero chain
    Xi[:, 5] = 0.0
    exact = (Ai.astype(np.int64) @ Xi.astype(np.int64)).astype(np.float64)
    got = product(Ai, Xi)            # the integer product is rebuilt exactly; the conversion to fp64 rounds once or twice
    assert np.max(np.abs(got - exact)) <= 4e-16 * np.max(np.abs(exact)) and np.all(got[3] == 0) and np.all(got[:, 5] == 0)
    A = rng.normal(size=(M, K)) * np.exp(rng.normal(size=(M, 1)) * 3)      # row scales over orders of magnitude
    X = rng.normal(size=(K, N)) * np.exp(rng.normal(size=(1, N)) * 3)
    got = product(A, X)
    bound = np.abs(A) @ np.abs(X)
    assert np.max(np.abs(got - A @ X) / bound) < 2.0**-42
