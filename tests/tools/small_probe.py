"""Timing of the single-launch path on the shipped instances: device time of the kernel vs wall time of sdpcuda_solve."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from scip_sdp_b200 import abi, misdp
G = os.path.join(ROOT, "tests", "golden")
gpu = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
for name in ["example_small.dat-s", "example_TT.dat-s.gz", "example_CLS.dat-s.gz", "example_MkP.dat-s.gz"]:
    fp, _ = misdp.read_sdpa(os.path.join(G, name)).rows_to_bounds().flatten()
    for path in ("s", "m"):
        os.environ["SDPCUDA_PATH"] = path
        gpu.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False)
        t = time.perf_counter(); r = gpu.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False); w = time.perf_counter() - t
        t = time.perf_counter(); r2 = gpu.solve_resident(gaptol=1e-5, feastol=1e-5); w2 = time.perf_counter() - t
        print(f"{name:22s} path {path}: iters {r['iterations']:3d} launches {r['launches']:5d} device {r['device_ms']:.3f} ms  solve wall {1e3*w:.3f} ms  resident wall {1e3*w2:.3f} ms (device {r2['device_ms']:.3f})")
    t = time.perf_counter(); rc = cpu.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False); wc = time.perf_counter() - t
    print(f"{name:22s} cpu oracle: iters {rc['iterations']:3d} wall {1e3*wc:.3f} ms")
