mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2bj_pytest_all.log 2>&1; tail -3 gpurun_out/r2bg_pytest_all.log
timeout 600 python tools/frontier_rates.py example_MkP example_TT example_CLS 2>&1 | tail -4
