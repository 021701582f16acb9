set -x
cat > /tmp/ch.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
g.time_kernel(10, 2000, 1)
g.time_kernel(11, 2000, 1)
print(g.time_kernel(2, 2000, 10), g.time_kernel(3, 2000, 10))
P
for sw in "SDPCUDA_DAG_WHELP=0" ""; do echo "$sw"; env $sw timeout 120 python /tmp/ch.py 2>&1 | grep "chain\] n\|^(" ; done
