set -x
cat > /tmp/fr.py <<'P'
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in sys.argv[1:]:
    r = bench.gpu_node_workload(g, lib, name, 0, t, 3)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"], 1), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3), 1), "span ms", round(r["device_ms_first_launch_to_last_result"], 2), "counted", r["counted"], r["max_rel_diff_to_oracle"], flush=True)
P
SDPCUDA_BATCH_PROFILE=1 timeout 300 python /tmp/fr.py example_TT 2>&1 | grep "\[batch\]\|\[nodes\]\|e2e" | tail -22
SDPCUDA_BATCH_CHUNKS=2 SDPCUDA_BATCH_PROFILE=1 timeout 300 python /tmp/fr.py example_CLS 2>&1 | grep "\[batch\]\|\[nodes\]\|e2e" | tail -8
