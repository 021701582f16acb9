# Round 2, GPU call 9: GPU suite after the packing/TINY changes, node workloads, headline (grid of the DAG kernel)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_all.log 2>&1; tail -6 gpurun_out/r2e_pytest_all.log
timeout 300 python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in ("example_TT", "example_MkP", "example_CLS"):
    r = bench.gpu_node_workload(g, lib, name, 0, t, 3)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"]), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3)), "counted", r["counted"], "maxrel", r["max_rel_diff_to_oracle"])
for n in (2000,):
    for kind, name in ((3, "potrf"), (2, "potrf+inverse")):
        ms, fl = g.time_kernel(kind, n, 5)
        print(n, name, round(ms, 3), "ms", round(fl / ms / 1e9, 2), "TF/s")
P
timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2e_bench_nonodes.json 2> gpurun_out/r2e_bench.err; cut -c1-250 gpurun_out/r2e_bench_nonodes.json; tail -3 gpurun_out/r2e_bench.err
