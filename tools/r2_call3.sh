# Round 2, GPU call 3: chain profile of the tile-DAG Cholesky, whole GPU suite
set -x
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2c_dag_chain.log 2>&1 <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
for n in (2000, 4096):
    print(g.time_kernel(10, n, 1))
g.time_kernel(9, 64, 1)
P
cat gpurun_out/r2c_dag_chain.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_all.log 2>&1; tail -25 gpurun_out/r2c_pytest_all.log
