set -x
mkdir -p gpurun_out
cat > /tmp/potrf_t.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
for n in (1000, 2000, 3000):
    for kind, name in ((3, "potrf"), (2, "potrf+inverse")):
        ms, fl = g.time_kernel(kind, n, 10)
        print(f"  {name:14s} n={n:5d} {ms:8.3f} ms {fl / ms / 1e9:6.2f} TF/s", flush=True)
g.time_kernel(10, 2000, 1)
P
for sw in "SDPCUDA_DAG_SMEM_KB=116" "SDPCUDA_DAG_SMEM_KB=116 SDPCUDA_DAG_GRID=74" "SDPCUDA_DAG_GRID=74" "SDPCUDA_DAG_SMEM_KB=116 SDPCUDA_DAG_GRID=48"; do
echo "== $sw"
env $sw timeout 300 python /tmp/potrf_t.py 2>&1 | grep -v "dag chain\] j"
done
for sw in "SDPCUDA_DAG_SMEM_KB=116 SDPCUDA_DAG_GRID=74" "SDPCUDA_DAG_GRID=74"; do
env $sw timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2w_bench.json 2>> gpurun_out/r2w_bench.err; echo "$sw"; python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['profiled_solve_ms'], r['share_of_step'])"
done
