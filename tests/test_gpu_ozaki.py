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


@pytest.mark.parametrize("M,N,K,SA,SB", [(128, 256, 128, 1, 1), (256, 512, 384, 2, 3), (384, 384, 1152, 3, 2)])
def test_i8_slice_products_are_exact(M, N, K, SA, SB):
    import torch

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randint(-64, 64, (SA, M, K), generator=g, device="cuda", dtype=torch.int8)
    B = torch.randint(-64, 64, (SB, N, K), generator=g, device="cuda", dtype=torch.int8)
    orders = SA + SB - 1
    got = _i8_gemm(A, B, orders)
    ref = torch.zeros(orders, M, N, dtype=torch.float64, device="cuda")
    for s in range(SA):
        for t in range(SB):
            ref[s + t] += A[s].double() @ B[t].double().T
    assert torch.equal(got.double(), ref)


@pytest.mark.parametrize("name", ["dense_direct_4s", "dense_direct_f64", "dense_full_mass_3s"])
def test_ozaki_path_reproduces_the_golden_trajectories(monkeypatch, name):
    """The dense direct products on tcgen05 (int8 slices, exact int32 accumulation, fp64 recombination)
    forced on for the small golden cases: same 1e-10 / identical-decision bar as the DMMA path."""
    import torch

    import cases
    from helpers import build_mirror, load_golden, rel_err
    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten

    monkeypatch.setenv("HMCB_OZAKI", "1")
    inp, ref = load_golden(name)
    s = cases.SETTINGS[name]
    K, C_, d = inp["z"].shape
    post, mass = build_mirror(name, inp)
    eng = Engine(flatten(describe(post)), describe_mass(mass), C_, integrator=s["integrator"], amount_of_steps=s["steps"])
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
