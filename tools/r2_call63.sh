mkdir -p gpurun_out
for k in 1 2 3; do SDPCUDA_LANE_TRACE=1 timeout 600 python tools/lanes_slices.py CLS-syn TT-500 ; done > gpurun_out/r2as_lanes_trace.log 2>&1
grep -v "^\[lanes\]" gpurun_out/r2as_lanes_trace.log | awk '{ if ($11+0 > 150) print }' | head -20
grep "^\[lanes\]" gpurun_out/r2as_lanes_trace.log | awk '{ for(i=1;i<=NF;i++) if($i=="solve") { if ($(i+1)+0 > 100) print } }' | head -40
echo total lines; wc -l gpurun_out/r2as_lanes_trace.log
