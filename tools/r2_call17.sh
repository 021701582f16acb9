set -x
mkdir -p gpurun_out
SDPCUDA_BATCH_PROFILE=1 timeout 300 python - <<'P' 2>&1 | tail -40
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in ("example_TT", "example_MkP", "example_CLS"):
    r = bench.gpu_node_workload(g, lib, name, 0, t, 2)
    print(name, "e2e nodes/s", round(r["counted"] / r["wall_s"]), "device nodes/s", round(r["counted"] / (r["device_ms"] / 1e3)), flush=True)
P
for sw in "" "SDPCUDA_DEVICE_CHECK=1"; do
env $sw timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2k_bench.json 2>> gpurun_out/r2k_bench.err; echo "$sw"; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'])"
done
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_${tool}_chol_gemm.log python -m pytest tests/test_gpu_kernels.py -q -k "(tile_dag and 1000 and dag) or (large_dgemm and 1100 and tma)" > gpurun_out/r2_sanitizer_${tool}_chol_gemm.out 2>&1
  tail -3 gpurun_out/r2_sanitizer_${tool}_chol_gemm.log; tail -2 gpurun_out/r2_sanitizer_${tool}_chol_gemm.out
done
