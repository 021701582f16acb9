"""One resident solve of a synthetic shape (for ncu launch lists): python tools/solve_once.py cls|tt500|mkp120|maxcut2000"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi, generators
name = sys.argv[1]
make = {"tt500": lambda: generators.truss(6, 6, 500, seed=1001), "cls": lambda: generators.cls(199, 99, 10, seed=2002),
        "mkp120": lambda: generators.mkp(120, seed=3003), "maxcut2000": lambda: generators.maxcut(2000, 0.01, seed=4004)}
gpu = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
fp = make[name]().flatten()[0]
r = gpu.solve(fp, fetch=False, gaptol=1e-5, feastol=1e-5)
print(name, r["phase_name"], r["iterations"], r["device_ms"], r["launches"])
