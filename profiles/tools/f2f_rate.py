"""fp32 -> fp64 conversion rate next to fp64 adds (hmcb_debug_fp64_peak kind 2) vs the DFMA rate."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmclab_b200._engine import load_library
lib = load_library()
out = {}
for kind, key in ((0, "dfma_ginst"), (2, "f2f_plus_dadd_gconv")):
    ms, n = ctypes.c_double(), ctypes.c_double()
    assert lib.hmcb_debug_fp64_peak(0, kind, 20000, 3, ctypes.byref(ms), ctypes.byref(n)) == 0
    out[key] = n.value / (2.0 if kind == 0 else 1.0) / ms.value / 1e6
print(json.dumps(out))
