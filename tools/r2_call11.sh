set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "dgemm or potrf or cholesky" 2>&1 | tail -4
SDPCUDA_GEMM=tma timeout 300 python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
for kind, name in ((0, "dgemm NN"), (1, "dgemm NT"), (5, "syrk lower NT")):
    for n in (1000, 2000, 4096):
        ms, fl = g.time_kernel(kind, n, 5)
        print(f"  {name:14s} n={n:5d} {ms:8.3f} ms {fl / ms / 1e9:6.2f} TF/s")
for n in (2000, 7140):
    ms, fl = g.time_kernel(2, n, 5); print(f"  potrf+inverse n={n} {ms:.3f} ms")
P
timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2g_bench_tma.json 2>> gpurun_out/r2g_bench.err; cut -c1-200 gpurun_out/r2g_bench_tma.json
