set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_pytest_all.log 2>&1; tail -3 gpurun_out/r2z_pytest_all.log
for w in cls tt500; do timeout 120 python tools/solve_once.py $w 2>&1 | tail -1; done
timeout 120 python tools/phase_probe.py cls 2>&1 | grep phases | tail -2
