"""Per-iteration phase timing (verbose = 2) of the resident solve on the synthetic BASELINE shapes; potrf rate at Schur sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi, generators  # noqa: E402

which = sys.argv[1:] or ["tt500", "cls", "mkp120"]
gpu = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
make = {"tt500": lambda: generators.truss(6, 6, 500, seed=1001), "cls": lambda: generators.cls(199, 99, 10, seed=2002),
        "mkp120": lambda: generators.mkp(120, seed=3003), "mkp60": lambda: generators.mkp(60, seed=3003),
        "maxcut2000": lambda: generators.maxcut(2000, 0.01, seed=4004)}
for name in which:
    if name.startswith("potrf"):
        n = int(name[5:])
        for kind, label in ((3, "potrf"), (2, "potrf+inverse")):
            ms, work = gpu.time_kernel(kind, n, 5)
            print(f"{label} n={n}: {ms:.3f} ms  {work / ms * 1e-9:.2f} TFLOP/s", flush=True)
        continue
    fp = generators.dense_sdp_flat(600, 300) if name == "dense600" else make[name]().flatten()[0]
    kw = dict(gaptol=1e-5, feastol=1e-5)
    gpu.solve(fp, fetch=False, **kw)
    print("====", name, "m", fp.m, "blocks", list(fp.blocksizes), "nlp", fp.nlp, flush=True)
    r = gpu.solve_resident(verbose=2, **kw)
    print(name, r["phase_name"], r["iterations"], r["device_ms"], r["launches"], flush=True)
