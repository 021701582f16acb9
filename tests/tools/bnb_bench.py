"""B&B nodes/sec on the shipped BASELINE instances through the reference's sdpi.c + our binding: CUDA library vs CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("SHIM_QUIET", "1")
from harness import bnb, sdpi_ref  # noqa: E402
from scip_sdp_b200 import misdp  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
want = sys.argv[1:] or ["cuda", "oracle"]
libs = {}
if "cuda" in want:
    libs["cuda"] = sdpi_ref.SdpiLib(sdpi_ref.LIB_CUDA)
if "oracle" in want:
    libs["oracle"] = sdpi_ref.SdpiLib(sdpi_ref.LIB_ORACLE)
for name in ["example_small.dat-s", "example_TT.dat-s.gz", "example_CLS.dat-s.gz", "example_MkP.dat-s.gz"]:
    M = misdp.read_sdpa(os.path.join(G, name))
    for tag, lib in libs.items():
        r = bnb.solve_misdp(lib, M, timelimit=900)
        print(f"{name:24s} {tag:7s} status {r['status']:10s} obj {r['objval']:.6f} nodes {r['nodes']:5d} sdpcalls {r['sdpcalls']:5d} "
              f"iters {r['iterations']:6d} time {r['seconds']:.2f}s  nodes/s {r['nodes'] / r['seconds']:.1f}  ms/relaxation {1e3 * r['seconds'] / max(1, r['sdpcalls']):.2f}", flush=True)
