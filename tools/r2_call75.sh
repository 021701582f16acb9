SDPCUDA_BATCH_PROFILE=1 timeout 300 python tools/frontier_rates.py example_CLS 2>&1 | tail -12
