import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in sys.argv[1:]:
    r = bench.gpu_node_workload(g, lib, name, 0, t, 3)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"], 2), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3), 2), "counted", r["counted"], r["max_rel_diff_to_oracle"], r["resolved_with_stable_settings"], flush=True)
