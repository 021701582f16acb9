set -x
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2am_pytest_all.log 2>&1; tail -2 gpurun_out/r2am_pytest_all.log
timeout 300 python tools/frontier_rates.py example_CLS example_TT example_MkP 2>&1 | tail -3
bash tools/r2_phase_probe.sh 2>&1 | grep "directions\|example\|cycles\]" | head -9
timeout 300 python -c "
import os,sys
sys.path.insert(0,os.getcwd())
import bench
from scip_sdp_b200 import abi
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0)
for f, s in bench.TREES.values():
    r = bench.gpu_tree(g, lib, f, s); print(f, r['nodes'], round(r['nodes']/r['wall_s'],1), r['rounds'], flush=True)
" 2>&1 | tail -2
