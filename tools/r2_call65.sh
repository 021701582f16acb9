mkdir -p gpurun_out
for k in 1 2 3; do timeout 600 python tools/lanes_slices.py CLS-syn TT-500; done > gpurun_out/r2au_lanes_slices.log 2>&1
awk '{ print $1, $5, $7, $11 }' gpurun_out/r2au_lanes_slices.log | sort -k1,1 -k4,4n | awk '{ a[$1]=a[$1] " " $4 } END { for (k in a) print k, a[k] }'
