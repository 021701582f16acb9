set -x
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
for sw in "" "SDPCUDA_MPANEL=lookahead"; do echo "$sw"; env $sw timeout 200 python tools/solve_once.py mkp120 2>&1 | tail -1; env $sw timeout 200 python tools/phase_probe.py mkp120 2>&1 | grep "phases" | tail -1; done
SDPCUDA_MINV_MAX=0 timeout 200 python tools/solve_once.py tt500 2>&1 | tail -1
SDPCUDA_MINV_MAX=0 SDPCUDA_MPANEL=lookahead timeout 200 python tools/solve_once.py tt500 2>&1 | tail -1
