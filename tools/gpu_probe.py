"""First-contact probe on the GPU box: kernel timings (device-resident, CUDA events) and one solve per shape."""
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi, generators  # noqa: E402

L = abi.Lib(abi.PRODUCT_LIB)
S = abi.Solver(L, 0)
out = {}
ms, fl = S.time_kernel(4, 0, 5)
out["dmma_peak_tflops"] = fl / ms / 1e9
ms, w = S.time_kernel(2, 64, 20)
print("diag64 potrf+inv (memset + one diag kernel) ms", ms, flush=True)
for n in (512, 1024, 2000, 4096):
    for kind, name in ((0, "gemm_nn"), (1, "gemm_nt"), (5, "syrk_nt_lower"), (2, "potrf_inv"), (3, "potrf"), (6, "copy")):
        ms, w = S.time_kernel(kind, n, 5)
        key = f"{name}_{n}"
        out[key] = dict(ms=ms, rate=(w / ms / 1e9 if kind != 6 else w / ms / 1e6), unit="TFLOP/s" if kind != 6 else "GB/s")
        print(key, out[key], flush=True)
print(json.dumps(out))
sizes = [int(a) for a in sys.argv[1:]] or [200, 500, 1000]
for n in sizes:
    fp, _ = generators.maxcut(n, min(0.5, 20.0 / n), seed=4004).flatten()
    t = time.time()
    r = S.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False, verbose=int(os.environ.get("VERBOSE", "0")))
    print("maxcut", n, r["phase_name"], r["stop_name"], "iters", r["iterations"], "launches", r["launches"], "obj", r["dobj"],
          "dev_ms %.1f wall %.3f" % (r["device_ms"], time.time() - t), flush=True)
