# ncu --set full of the frontier kernels with the final code: ipm_tiny_batch_kernel on the 592-node example_MkP frontier (4 nodes per SM)
# and ipm_small_batch_kernel on the 148-node example_CLS frontier (dense Schur path).  Numbers under ncu are never bench values.
mkdir -p gpurun_out
cat > /tmp/front.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0)
print(bench.gpu_node_workload(g, lib, sys.argv[1], 0, nodesets.golden(), 1)["counted"])
P
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ipm_tiny_batch_kernel --launch-skip 1 -c 1 -o gpurun_out/r2c_ipm_tiny_batch_mkp -f python /tmp/front.py example_MkP > gpurun_out/r2c_ncu_tiny.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ipm_small_batch_kernel --launch-skip 1 -c 1 -o gpurun_out/r2c_ipm_small_batch_cls -f python /tmp/front.py example_CLS > gpurun_out/r2c_ncu_small.log 2>&1
ls -la gpurun_out/r2c_*.ncu-rep; tail -2 gpurun_out/r2c_ncu_tiny.log gpurun_out/r2c_ncu_small.log
