mkdir -p gpurun_out
( echo "== default"; timeout 300 python tools/small_overhead_probe.py
echo "== SDPCUDA_DEVICE_CHECK=0"; SDPCUDA_DEVICE_CHECK=0 timeout 300 python tools/small_overhead_probe.py
echo "== SDPCUDA_PACKED_SOLVE=1"; SDPCUDA_PACKED_SOLVE=1 timeout 300 python tools/small_overhead_probe.py ) > gpurun_out/r2ax_small_overhead.log 2>&1
cat gpurun_out/r2ax_small_overhead.log
