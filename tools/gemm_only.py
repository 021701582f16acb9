"""Profiling target: the standalone 2000^3 and 4096^3 DGEMM of the library (device resident), for ncu --set full captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi
S = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
for n in (2000, 4096):
    ms, fl = S.time_kernel(0, n, 2)
    print(n, ms, fl / ms / 1e9)
