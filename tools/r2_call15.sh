set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest_all.log 2>&1; tail -5 gpurun_out/r2i_pytest_all.log
for sw in "" "SDPCUDA_CHOL_PAIR=0" "SDPCUDA_CHOL_PAIR=0 SDPCUDA_CHOL_INV=levels"; do
env $sw timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2i_bench.json 2>> gpurun_out/r2i_bench.err; echo "$sw"; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['profiled_solve_ms'], r['share_of_step'])"
done
tail -3 gpurun_out/r2i_bench.err
timeout 300 python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
for n in (1000, 2000, 4096, 7140):
    for kind, name in ((3, "potrf"), (2, "potrf+inverse")):
        ms, fl = g.time_kernel(kind, n, 5)
        print(f"  {name:14s} n={n:5d} {ms:8.3f} ms {fl / ms / 1e9:6.2f} TF/s")
P
