set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
cat > /tmp/potrf_t.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
for n in (1000, 1501, 2000, 3000):
    for kind, name in ((3, "potrf"), (2, "potrf+inverse")):
        ms, fl = g.time_kernel(kind, n, 10)
        print(f"  {name:14s} n={n:5d} {ms:8.3f} ms {fl / ms / 1e9:6.2f} TF/s", flush=True)
P
for sw in "" "SDPCUDA_DAG_WHELP=0"; do echo "$sw"; env $sw timeout 120 python /tmp/potrf_t.py 2>&1 | grep inverse; done
for sw in "" "SDPCUDA_DAG_WHELP=0"; do
env $sw timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2ah_bench.json 2>> gpurun_out/r2ah_bench.err; echo "$sw"; python -c "
import json; d=json.load(open('gpurun_out/r2ah_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['iterations_per_step'], d['objective'], r['frac'])"
done
