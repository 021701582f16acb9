set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solver.py tests/test_gpu_sdpi.py -m gpu -q -x > gpurun_out/r2aa_pytest.log 2>&1; tail -3 gpurun_out/r2aa_pytest.log
for sw in "" "SDPCUDA_RANK1=0" "" "SDPCUDA_RANK1=0"; do echo "$sw"; env $sw timeout 120 python tools/solve_once.py tt500 2>&1 | tail -1; done
timeout 120 python tools/phase_probe.py tt500 maxcut2000 2>&1 | grep "phases\|OPT" | tail -4
