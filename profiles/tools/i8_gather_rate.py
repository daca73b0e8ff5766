"""Throughput of the gathered (block-sparse) tcgen05 slice-product kernel at config-4-like shapes: G q
(391 bundles of 128 rays x ~1150 listed cells) and G^T r (79 bundles of 128 cells x ~5760 listed rays),
8192 chains, digits 5 x 6 / 6 x 6, orders 0..5."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmclab_b200._engine import load_library
lib = load_library()
out = {}
N = 8192
for name, (nb, kb, rows_b, SA) in {"G q": (391, 9, 10000, 5), "G^T r": (79, 45, 50048, 6)}.items():
    SB, orders = 6, 6
    Ktot = nb * kb * 128
    bundles = torch.as_tensor(np.stack([np.arange(nb) * kb * 128, np.full(nb, kb)], axis=1).astype(np.int32)).cuda()
    # a bundle's list: a band of neighbouring rows (what clustered rays give), shuffled
    rng = np.random.default_rng(0)
    lst = np.concatenate([(rng.integers(0, rows_b) + rng.permutation(kb * 128 * 2)[: kb * 128]) % rows_b
                          for _ in range(nb)]).astype(np.int32)
    lst = torch.as_tensor(lst).cuda()
    A = torch.randint(-64, 64, (SA, 128, Ktot), device="cuda", dtype=torch.int8)
    B = torch.randint(-64, 64, (SB, rows_b, N), device="cuda", dtype=torch.int8)
    Cc = torch.empty(orders, nb * 128, N, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: lib.hmcb_debug_i8_gather_gemm(0, nb, N, Ktot, rows_b, SA, SB, orders, A.data_ptr(), bundles.data_ptr(),
                                                 lst.data_ptr(), B.data_ptr(), Cc.data_ptr(), st)
    for _ in range(2):
        assert call() == 0
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    pairs = sum(1 for s in range(SA) for t in range(SB) if s + t < orders)
    ops = 2.0 * nb * 128 * N * kb * 128 * pairs
    out[name] = {"ms": min(ts), "pairs": pairs, "int8_tops": ops / min(ts) / 1e9, "bundles": nb, "kblocks": kb}
    del A, B, Cc
print(json.dumps(out))
