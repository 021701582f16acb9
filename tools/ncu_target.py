"""Profiling target: one max-cut relaxation (n from argv, default 2000) solved twice (upload + resident) for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi, generators  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
fp, _ = generators.maxcut(n, min(0.5, 20.0 / n), seed=4004).flatten()
S = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
r = S.solve(fp, gaptol=1e-5, feastol=1e-5, maxiter=maxiter, fetch=False)
print(r["phase_name"], r["iterations"], r["launches"], r["device_ms"])
