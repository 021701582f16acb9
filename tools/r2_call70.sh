mkdir -p gpurun_out
timeout 300 python tools/small_phases.py > gpurun_out/r2ay_small_phases.log 2>&1
grep -v "^ *[0-9]* *[0-9.e+-]* *[0-9.e+-]*  *[0-9.e+-]* " gpurun_out/r2ay_small_phases.log | tail -12
