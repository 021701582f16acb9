import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scip_sdp_b200 import abi, misdp
G = os.path.join(ROOT, "tests", "golden")
gpu = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
os.environ["SDPCUDA_PATH"] = "s"
for name in ["example_TT.dat-s.gz", "example_CLS.dat-s.gz", "example_MkP.dat-s.gz"]:
    fp, _ = misdp.read_sdpa(os.path.join(G, name)).rows_to_bounds().flatten()
    print(name, flush=True)
    r = gpu.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False, verbose=2)
    print(r["iterations"], r["device_ms"], flush=True)
