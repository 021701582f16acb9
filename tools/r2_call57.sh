set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2am_pytest_all.log 2>&1; tail -3 gpurun_out/r2am_pytest_all.log
( echo "== default"; timeout 600 python tools/concurrent_sdpi_probe.py example_TT,example_MkP,example_CLS 128 1,4,8,16,32
echo "== SDPCUDA_PACKED_SOLVE=1"; SDPCUDA_PACKED_SOLVE=1 timeout 600 python tools/concurrent_sdpi_probe.py example_TT,example_MkP,example_CLS 128 1,8,32
echo "== SDPCUDA_SMALL_M=128"; SDPCUDA_SMALL_M=128 timeout 600 python tools/concurrent_sdpi_probe.py example_MkP 128 1,8,32 ) > gpurun_out/r2am_concurrent_sdpi.log 2>&1
cat gpurun_out/r2am_concurrent_sdpi.log
