# Round 2, GPU call 2: tile-DAG Cholesky (parity vs LAPACK, timing vs the recursive chain), GPU suite, headline bench without node workloads
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf or cholesky" > gpurun_out/r2b_pytest_chol.log 2>&1; tail -15 gpurun_out/r2b_pytest_chol.log
timeout 300 python - > gpurun_out/r2b_potrf_timing.log 2>&1 <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
pk_ms, pk_fl = g.time_kernel(4, 0, 3)
print("dmma peak TF", pk_fl / pk_ms / 1e9)
for var in ("rec", "dag"):
    os.environ["SDPCUDA_CHOL"] = var
    for n in (512, 1000, 1501, 2000, 4096, 7140):
        for kind, name in ((3, "potrf"), (2, "potrf+inverse")):
            ms, fl = g.time_kernel(kind, n, 5)
            print(f"{var} {name:14s} n={n:5d}  {ms:8.3f} ms  {fl / ms / 1e9:6.2f} TF/s")
P
cat gpurun_out/r2b_potrf_timing.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_all.log 2>&1; tail -15 gpurun_out/r2b_pytest_all.log
timeout 300 python bench.py --no-nodes > gpurun_out/r2b_bench_n1_nonodes.json 2> gpurun_out/r2b_bench_n1.err; cut -c1-300 gpurun_out/r2b_bench_n1_nonodes.json; tail -3 gpurun_out/r2b_bench_n1.err
SDPCUDA_CHOL=rec timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2b_bench_n1_nonodes_rec.json 2>> gpurun_out/r2b_bench_n1.err; cut -c1-300 gpurun_out/r2b_bench_n1_nonodes_rec.json
