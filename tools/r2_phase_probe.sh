for tiny in 1 0; do
SDPCUDA_BATCH_TINY=$tiny timeout 300 python - <<'P' 2>&1 | grep -E "cycles|iterations|TINY"
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
print("TINY", os.environ["SDPCUDA_BATCH_TINY"])
for name in ("example_TT", "example_MkP", "example_CLS"):
    M = nodesets.WORKLOADS[name][0]()
    codes, want = nodesets.frontier_of_rank(name, 0, table=t)
    lbs, ubs = nodesets.node_bounds(M, codes[:2])
    model = abi.Model(lib, M)
    out = g.solve_nodes(model, lbs, ubs, lean=True, gaptol=1e-5, feastol=1e-5, verbose=2)
    print(name, "iterations", out["results"]["iterations"], "device ms", out["results"]["device_ms"][0], flush=True)
P
done

