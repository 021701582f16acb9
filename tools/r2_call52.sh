set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
python -c "
import os,sys
sys.path.insert(0,os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
for n in (1000, 2000, 3000, 4096, 7140):
    print(n, 'potrf', round(g.time_kernel(3, n, 5)[0],3), 'with inverse', round(g.time_kernel(2, n, 5)[0],3), 'panels', round(g.time_kernel(13, n, 5)[0],3), flush=True)
" 2>&1 | tail -5
timeout 200 python tools/solve_once.py mkp120 2>&1 | tail -1
