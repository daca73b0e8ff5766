"""Throughput of the tcgen05 int8 slice-product kernel at the config-3 shapes (both products of one
gradient evaluation: digits 5 x 6 and 6 x 6, orders 0..5 = 20 / 21 slice pairs)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmclab_b200._engine import load_library
lib = load_library()
out = {}
for name, (M, N, K, SA) in {"G q (10112 x 8192 x 2048)": (10112, 8192, 2048, 5),
                            "G^T r (2048 x 8192 x 10112)": (2048, 8192, 10112, 6)}.items():
    SB, orders = 6, 6
    A = torch.randint(-64, 64, (SA, M, K), device="cuda", dtype=torch.int8)
    B = torch.randint(-64, 64, (SB, N, K), device="cuda", dtype=torch.int8)
    Cc = torch.empty(orders, M, N, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        assert lib.hmcb_debug_i8_gemm(0, M, N, K, SA, SB, orders, A.data_ptr(), B.data_ptr(), Cc.data_ptr(), st) == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lib.hmcb_debug_i8_gemm(0, M, N, K, SA, SB, orders, A.data_ptr(), B.data_ptr(), Cc.data_ptr(), st); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    pairs = sum(1 for s in range(SA) for t in range(SB) if s + t < orders)
    ops = 2.0 * M * N * K * pairs
    out[name] = {"ms": min(ts), "pairs": pairs, "int8_tops": ops / min(ts) / 1e9}
print(json.dumps(out))
