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
