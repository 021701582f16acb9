mkdir -p gpurun_out
( echo "== default"; timeout 900 python tests/tools/bnb_bench.py cuda oracle; echo "== SDPCUDA_PACKED_SOLVE=1"; SDPCUDA_PACKED_SOLVE=1 timeout 600 python tests/tools/bnb_bench.py cuda ) > gpurun_out/r2aw_bnb_through_sdpi.log 2>&1
cat gpurun_out/r2aw_bnb_through_sdpi.log
