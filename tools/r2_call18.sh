set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest_all.log 2>&1; tail -12 gpurun_out/r2l_pytest_all.log
timeout 300 python - <<'P' 2>&1 | tail -8
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in ("example_TT", "example_MkP", "example_CLS", "TT-500"):
    r = bench.gpu_node_workload(g, lib, name, 0, t, 3)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"], 1), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3), 1), "counted", r["counted"], flush=True)
P
SDPCUDA_DAG_WATCHDOG_S=0 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_chol_gemm.log python -m pytest tests/test_gpu_kernels.py -q -k "(tile_dag and 1000 and dag) or (large_dgemm and 1100 and tma)" > gpurun_out/r2_sanitizer_racecheck_chol_gemm.out 2>&1
tail -3 gpurun_out/r2_sanitizer_racecheck_chol_gemm.log; tail -2 gpurun_out/r2_sanitizer_racecheck_chol_gemm.out
timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2l_bench.json 2>> gpurun_out/r2l_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
