mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"
timeout 300 python tools/concurrent_sdpi_probe.py example_TT,example_MkP,example_CLS 128 1,8,16,32 > gpurun_out/r2bl_concurrent_sdpi_final.log 2>&1
cat gpurun_out/r2bl_concurrent_sdpi_final.log
