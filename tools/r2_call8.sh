# Round 2, GPU call 8: leaf (branch-free rows) re-test, batch kernel switches on the kernel-bound node workloads
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "potrf or cholesky" 2>&1 | tail -3
timeout 300 python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
g.time_kernel(9, 64, 1); g.time_kernel(9, 128, 1)
print(g.time_kernel(10, 2000, 1))
for n in (2000, 7140):
    for kind, name in ((3, "potrf"), (2, "potrf+inverse")):
        ms, fl = g.time_kernel(kind, n, 5)
        print(n, name, round(ms, 3), "ms", round(fl / ms / 1e9, 2), "TF/s")
P
for sw in "" "SDPCUDA_BATCH_TINY=1" "SDPCUDA_BATCH_TINY=1 SDPCUDA_BATCH_SMEM=1" "SDPCUDA_BATCH_SMEM=1"; do
env $sw timeout 300 python - <<'P'
import os, sys, json
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
print({k: os.environ.get(k) for k in ("SDPCUDA_BATCH_TINY", "SDPCUDA_BATCH_SMEM")})
for name in ("example_TT", "example_MkP", "example_CLS"):
    r = bench.gpu_node_workload(g, lib, name, 0, t, 3)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"]), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3)), "counted", r["counted"], "maxrel", r["max_rel_diff_to_oracle"])
P
done
