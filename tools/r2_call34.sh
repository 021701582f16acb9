set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do timeout 120 python tools/solve_once.py tt500 2>&1 | tail -1; done
timeout 60 python tools/leaf_probe.py 66 100 128 2>&1 | tail -3
