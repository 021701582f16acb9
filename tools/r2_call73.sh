mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2bb_pytest_all.log 2>&1; tail -4 gpurun_out/r2bb_pytest_all.log
timeout 600 python tools/frontier_rates.py example_CLS example_TT example_MkP 2>&1 | tail -4
timeout 300 python tools/small_phases.py 2>&1 | grep "cycles"
timeout 300 python tools/concurrent_sdpi_probe.py example_CLS 128 1,8
