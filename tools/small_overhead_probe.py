"""Where the time of ONE tiny relaxation goes on the way through the boundary: resident re-solve (kernel + result), full solve of the
library (upload + kernel), SCIPsdpiSolverLoadAndSolve of the binding (marshalling, hash, solve, getters, post-check)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scip_sdp_b200 import abi, misdp, sdpisolver_host
G = os.path.join(ROOT, "tests", "golden")
K = 300
lib = abi.Lib(abi.PRODUCT_LIB)
for name in ("example_small.dat-s", "example_TT.dat-s.gz", "example_MkP.dat-s.gz"):
    M = misdp.read_sdpa(os.path.join(G, name)).rows_to_bounds()
    fp = M.flatten()[0]
    g = abi.Solver(lib, 0)
    r = g.solve(fp, fetch=False, gaptol=1e-5, feastol=1e-5)
    par = lib.default_params(gaptol=1e-5, feastol=1e-5)
    t = time.perf_counter()
    for _ in range(K): g.solve_resident(params=par)
    t_res = (time.perf_counter() - t) / K
    t = time.perf_counter()
    for _ in range(K): g.solve(fp, params=par, fetch=False)
    t_full = (time.perf_counter() - t) / K
    g.close()
    bp = sdpisolver_host.BoundaryProblem(M)
    s = sdpisolver_host.SdpiSolver(gaptol=1e-5, feastol=1e-5)
    s.load_and_solve(bp)
    t = time.perf_counter()
    for _ in range(K): s.load_and_solve(bp)
    t_bind = (time.perf_counter() - t) / K
    t = time.perf_counter()
    for _ in range(K): s.load_and_solve(bp); s.dual_sol(); s.flag("IsOptimal")
    t_bind2 = (time.perf_counter() - t) / K
    s.close()
    print(f"{name:22s} device {r['device_ms']:.3f} ms, {r['iterations']} iterations | resident re-solve {1e3 * t_res:.3f} ms | library solve {1e3 * t_full:.3f} ms | "
          f"LoadAndSolve {1e3 * t_bind:.3f} ms | + GetDualSol, IsOptimal {1e3 * t_bind2:.3f} ms", flush=True)
