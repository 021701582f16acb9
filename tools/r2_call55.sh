set -x
mkdir -p gpurun_out
timeout 900 python tools/concurrent_sdpi_probe.py example_TT,example_MkP,example_CLS 128 1,2,4,8,16,32 > gpurun_out/r2ak_concurrent_sdpi.log 2>&1
cat gpurun_out/r2ak_concurrent_sdpi.log | tail -30
