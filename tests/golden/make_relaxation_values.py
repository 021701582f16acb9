"""Pins the root-relaxation objectives of the synthetic BASELINE.json shapes with the CPU oracle (relaxation-level values are not
pinned anywhere in the reference).  Run in the build container (about 2 minutes on 8 cores):
    python tests/golden/make_relaxation_values.py
writes tests/golden/relaxation_values.json: {name: {m, blocks, nlp, dobj, pobj, iterations, gaptol}}."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from scip_sdp_b200 import abi, generators  # noqa: E402

SHAPES = {"tt500": lambda: generators.truss(6, 6, 500, seed=1001), "cls": lambda: generators.cls(199, 99, 10, seed=2002),
          "mkp60": lambda: generators.mkp(60, seed=3003), "mkp120": lambda: generators.mkp(120, seed=3003),
          "maxcut2000": lambda: generators.maxcut(2000, 0.01, seed=4004), "dense600x300": lambda: generators.dense_sdp_flat(600, 300, seed=5005)}

if __name__ == "__main__":
    cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    out = {}
    for name, make in SHAPES.items():
        M = make()
        fp = M if isinstance(M, abi.FlatProblem) else M.flatten()[0]          # dense_sdp_flat is in solver form already
        t = time.time()
        r = cpu.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False)
        assert r["phase_name"] == "pdOPT", (name, r["phase_name"])
        out[name] = dict(m=int(fp.m), blocks=[int(b) for b in fp.blocksizes], nlp=int(fp.nlp), dobj=r["dobj"], pobj=r["pobj"],
                         iterations=int(r["iterations"]), gaptol=1e-5)
        print(name, out[name], f"{time.time() - t:.1f} s", flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "relaxation_values.json"), "w") as f:
        json.dump(out, f, indent=1)
