set -x
cat > /tmp/fr.py <<'P'
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in sys.argv[1:]:
    r = bench.gpu_node_workload(g, lib, name, 0, t, 1)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"], 2), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3), 2), flush=True)
P
SDPCUDA_UPLOAD_PROFILE=1 SDPCUDA_BATCH_PROFILE=1 timeout 300 python /tmp/fr.py CLS-syn 2>&1 | tail -40
