set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
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
timeout 300 python /tmp/potrf_t.py 2>&1 | grep -v "dag chain\] j [1-9]"
