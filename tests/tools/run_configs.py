"""Checker-side comparison (lives under tests/ because it loads the CPU oracle): solves the synthetic BASELINE.json shapes on the
GPU and (optionally) on the CPU oracle; prints objective, iterations, time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from scip_sdp_b200 import abi, generators  # noqa: E402

which = sys.argv[1:] or ["tt500", "cls", "mkp120"]
with_oracle = os.environ.get("ORACLE", "1") == "1"
gpu = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB)) if with_oracle else None
make = {"tt500": lambda: generators.truss(6, 6, 500, seed=1001), "cls": lambda: generators.cls(199, 99, 10, seed=2002),
        "mkp120": lambda: generators.mkp(120, seed=3003), "mkp60": lambda: generators.mkp(60, seed=3003),
        "maxcut2000": lambda: generators.maxcut(2000, 0.01, seed=4004)}
out = {}
for name in which:
    t = time.time()
    fp, _ = make[name]().flatten()
    tg = time.time() - t
    kw = dict(gaptol=1e-5, feastol=1e-5)
    r = gpu.solve(fp, fetch=False, **kw)
    r2 = gpu.solve_resident(**kw)
    rec = dict(m=fp.m, blocks=[int(b) for b in fp.blocksizes], nlp=fp.nlp, gen_s=round(tg, 2), gpu_phase=r["phase_name"], gpu_iters=r["iterations"],
               gpu_obj=r["dobj"], gpu_ms_first=round(1e3 * r["seconds"], 1), gpu_ms_resident=round(r2["device_ms"], 1), launches=r2["launches"])
    if cpu is not None:
        t = time.time()
        rc = cpu.solve(fp, fetch=False, **kw)
        rec.update(cpu_phase=rc["phase_name"], cpu_iters=rc["iterations"], cpu_obj=rc["dobj"], cpu_s=round(time.time() - t, 2),
                   relerr=abs(rc["dobj"] - r["dobj"]) / max(1.0, abs(rc["dobj"])))
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
