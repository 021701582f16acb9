set -x
mkdir -p gpurun_out
bash tools/r2_phase_probe.sh 2>&1 | grep -E "TINY|iterations|cycles"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest_all.log 2>&1; tail -5 gpurun_out/r2m_pytest_all.log
timeout 300 python - <<'P' 2>&1 | tail -8
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in ("example_TT", "example_MkP", "example_CLS"):
    r = bench.gpu_node_workload(g, lib, name, 0, t, 3)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"], 1), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3), 1), "counted", r["counted"], r["max_rel_diff_to_oracle"], flush=True)
for f, s in bench.TREES.values():
    print(f, bench.gpu_tree(g, lib, f, s), flush=True)
P
SDPCUDA_BATCH_TINY=0 timeout 300 python - <<'P' 2>&1 | tail -4
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
r = bench.gpu_node_workload(g, lib, "example_MkP", 0, t, 3)
print("TINY=0 example_MkP e2e nodes/s", round(r["counted"] / r["wall_s"], 1), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3), 1), flush=True)
print(bench.gpu_tree(g, lib, "example_MkP.dat-s.gz", -95.0), flush=True)
P
timeout 300 python tests/tools/bnb_bench.py cuda 2>&1 | tail -4
SDPCUDA_SMALL_M=128 timeout 300 python tests/tools/bnb_bench.py cuda 2>&1 | tail -1
