set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2an_pytest_all.log 2>&1; tail -5 gpurun_out/r2an_pytest_all.log
( echo "== default (packed + blocking wait with >= 3 objects alive)"; timeout 600 python tools/concurrent_sdpi_probe.py example_TT,example_MkP,example_CLS 128 1,2,4,8,16,32
echo "== SDPCUDA_BLOCKING_WAIT=0"; SDPCUDA_BLOCKING_WAIT=0 timeout 600 python tools/concurrent_sdpi_probe.py example_TT,example_MkP,example_CLS 128 8,16,32 ) > gpurun_out/r2an_concurrent_sdpi.log 2>&1
cat gpurun_out/r2an_concurrent_sdpi.log
