# Round 2: ncu evidence. (1) launch list of ONE max-cut 2000 solve, (2) full captures of the TMA GEMM (2000^3, 4096^3), the tile-DAG
# Cholesky (2000, 7140) and the 256-thread frontier-batch kernel (example_TT, 592 nodes).  Numbers under ncu are never bench values.
set -x
mkdir -p gpurun_out
cat > /tmp/one_solve.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi, generators
fp, _ = generators.maxcut(2000, 0.01, seed=4004).flatten()
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
r = g.solve(fp, fetch=False, gaptol=1e-5, feastol=1e-5, absgaptol=5e-6)
print(r["phase_name"], r["iterations"], r["launches"], r["device_ms"])
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_maxcut2000.csv python /tmp/one_solve.py > gpurun_out/r2_ncu_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_maxcut2000.csv > gpurun_out/r2_launches_maxcut2000.txt 2>/dev/null; cat gpurun_out/r2_launches_maxcut2000.txt
cat > /tmp/kern.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
kind, n = int(sys.argv[1]), int(sys.argv[2])
print(g.time_kernel(kind, n, 1))
P
for n in 2000 4096; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_tma_kernel --launch-skip 3 -c 1 -o gpurun_out/r2_gemm_tma_$n -f python /tmp/kern.py 0 $n > gpurun_out/r2_ncu_gemm_$n.log 2>&1
done
for n in 2000 7140; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_dag_kernel --launch-skip 3 -c 1 -o gpurun_out/r2_potrf_dag_$n -f python /tmp/kern.py 3 $n > gpurun_out/r2_ncu_dag_$n.log 2>&1
done
cat > /tmp/front.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0)
print(bench.gpu_node_workload(g, lib, "example_TT", 0, nodesets.golden(), 1)["counted"])
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ipm_tiny_batch_kernel --launch-skip 1 -c 1 -o gpurun_out/r2_ipm_tiny_batch_tt -f python /tmp/front.py > gpurun_out/r2_ncu_tiny.log 2>&1
ls -la gpurun_out/*.ncu-rep
