# Round 2, session 2: ncu evidence after the chain-CTA Cholesky. (1) launch list of ONE max-cut 2000 solve (same command as tools/r2_ncu.sh),
# (2) full capture of potrf_dag_kernel<true> at n = 2000 with the inverse.  Numbers under ncu are never bench values.
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
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2b_launches_maxcut2000.csv python /tmp/one_solve.py > gpurun_out/r2b_ncu_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_launches_maxcut2000.csv > gpurun_out/r2b_launches_maxcut2000.txt 2>/dev/null; head -14 gpurun_out/r2b_launches_maxcut2000.txt
cat > /tmp/kern.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
kind, n = int(sys.argv[1]), int(sys.argv[2])
print(g.time_kernel(kind, n, 1))
P
SDPCUDA_DAG_WATCHDOG_S=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_dag_kernel --launch-skip 3 -c 1 -o gpurun_out/r2b_potrf_dag_chain_2000 -f python /tmp/kern.py 2 2000 > gpurun_out/r2b_ncu_dag_2000.log 2>&1
tail -3 gpurun_out/r2b_ncu_dag_2000.log
