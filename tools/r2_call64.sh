mkdir -p gpurun_out
for k in 1 2; do SDPCUDA_UPLOAD_PROFILE=1 SDPCUDA_LANE_TRACE=1 timeout 600 python tools/lanes_slices.py CLS-syn ; done > gpurun_out/r2at_upload_trace.log 2>&1
grep "^\[upload\]" gpurun_out/r2at_upload_trace.log | awk '{ if ($(NF-1)+0 > 30) print }' | sort | uniq -c | sort -rn | head -30
echo; grep "^\[upload\]" gpurun_out/r2at_upload_trace.log | awk '{ v=$(NF-1)+0; $NF=""; $(NF-1)=""; s[$0]+=v; n[$0]++; if (v>m[$0]) m[$0]=v } END { for (k in s) printf "%-40s n=%d mean %.2f max %.2f\n", k, n[k], s[k]/n[k], m[k] }'
