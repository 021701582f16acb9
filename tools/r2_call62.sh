mkdir -p gpurun_out
timeout 600 python tools/lanes_slices.py TT-500 CLS-syn > gpurun_out/r2ar_lanes_slices.log 2>&1
cat gpurun_out/r2ar_lanes_slices.log
