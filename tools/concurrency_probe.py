"""Concurrent node relaxations on one GPU: T host threads, each with its own solver handle/stream (what SCIP's concurrent
solver threads do through SCIPsdpiSolverCreate), repeatedly solving a resident small relaxation.  Aggregate relaxations/s."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scip_sdp_b200 import abi, misdp
G = os.path.join(ROOT, "tests", "golden")
name = sys.argv[1] if len(sys.argv) > 1 else "example_TT.dat-s.gz"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
fp, _ = misdp.read_sdpa(os.path.join(G, name)).rows_to_bounds().flatten()
lib = abi.Lib(abi.PRODUCT_LIB)
for T in (1, 2, 4, 8, 16, 32, 64, 128):
    solvers = [abi.Solver(lib, 0) for _ in range(T)]
    for s in solvers:
        s.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False)
    par = lib.default_params(gaptol=1e-5, feastol=1e-5)
    def work(s):
        for _ in range(K):
            s.solve_resident(params=par)
    th = [threading.Thread(target=work, args=(s,)) for s in solvers]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t
    print(f"{name} threads {T:4d}: {T * K / dt:9.1f} relaxations/s  ({1e3 * dt / K:.2f} ms per relaxation per thread)", flush=True)
    for s in solvers: s.close()
