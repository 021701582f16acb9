"""Cycle counts of the phases of the diagonal-block (leaf) kernel and timing of the factorisation chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi
S = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
for nl in (64, 128):
    S.time_kernel(9, nl, 1)
for n in [int(a) for a in sys.argv[1:]] or [64, 128, 256, 1501, 2000]:
    ms, work = S.time_kernel(2, n, 10)
    ms3, work3 = S.time_kernel(3, n, 10)
    print(f"n={n}: potrf+inverse {ms * 1e3:.1f} us ({work / ms * 1e-9:.2f} TF)   potrf {ms3 * 1e3:.1f} us ({work3 / ms3 * 1e-9:.2f} TF)", flush=True)
if os.environ.get("PROFILE_CLASSES"):
    for n in (1501, 2000):
        S.set_profiling(True)
        S.time_kernel(2, n, 1)
        prof = S.get_profile()
        S.set_profiling(False)
        print(f"potrf+inverse n={n} by class:", {k: (v["launches"], round(v["ms"] * 1e3, 1)) for k, v in prof.items() if v["launches"]}, "(4 runs)", flush=True)
