set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zfrontier.py tests/test_gpu_boundary.py -m gpu -q -x 2>&1 | tail -2
cat > /tmp/fr.py <<'P'
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in ("example_TT", "example_MkP", "example_CLS"):
    r = bench.gpu_node_workload(g, lib, name, 0, t, 3)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"], 1), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3), 1), "span ms", round(r["device_ms_first_launch_to_last_result"], 2), "counted", r["counted"], r["max_rel_diff_to_oracle"], flush=True)
P
for ch in "" 1 2 3; do echo "CHUNKS=$ch"; if [ -z "$ch" ]; then timeout 300 python /tmp/fr.py 2>&1 | tail -3; else SDPCUDA_BATCH_CHUNKS=$ch timeout 300 python /tmp/fr.py 2>&1 | tail -3; fi; done
SDPCUDA_BATCH_PROFILE=1 timeout 300 python /tmp/fr.py 2>&1 | grep "\[batch\]\|\[nodes\]" | tail -24
