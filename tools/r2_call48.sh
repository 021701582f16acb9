set -x
cat > /tmp/tree.py <<'P'
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0)
for f, s in bench.TREES.values():
    r = bench.gpu_tree(g, lib, f, s)
    print(f, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, "nodes/s", round(r["nodes"] / r["wall_s"], 1), flush=True)
P
for sw in "" "SDPCUDA_BATCH_TINY=0"; do echo "$sw"; env $sw SDPCUDA_BATCH_PROFILE=1 timeout 300 python /tmp/tree.py 2>/tmp/err.log | tail -2; grep "\[batch\]" /tmp/err.log | tail -16 | cut -c1-120; done
