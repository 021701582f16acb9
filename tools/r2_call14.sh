set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "potrf or cholesky" 2>&1 | tail -4
for v in levels kernel; do
SDPCUDA_CHOL_INV=$v timeout 300 python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
print("inverse:", os.environ.get("SDPCUDA_CHOL_INV"))
for n in (512, 1000, 1501, 2000, 4096, 7140):
    for kind, name in ((3, "potrf"), (2, "potrf+inverse")):
        ms, fl = g.time_kernel(kind, n, 5)
        print(f"  {name:14s} n={n:5d} {ms:8.3f} ms {fl / ms / 1e9:6.2f} TF/s")
P
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest_all.log 2>&1; tail -5 gpurun_out/r2h_pytest_all.log
timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2h_bench.json 2>> gpurun_out/r2h_bench.err; cut -c1-200 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
