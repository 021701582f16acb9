set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
python bench.py > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1b_bench_ref.json 2> gpurun_out/r1b_bench_ref.err
ORACLE=1 python tests/tools/run_configs.py tt500 cls mkp120 mkp60 > gpurun_out/r1b_configs.log 2>&1
for w in frontier-tt500 frontier-cls frontier-mkp60 frontier-mkp120; do python bench.py --workload $w > gpurun_out/r1b_$w.json 2>> gpurun_out/r1b_frontier.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r1b_launches.csv python tools/ncu_target.py 2000 100 > gpurun_out/ncu_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r1b_launches.csv > gpurun_out/r1b_launches_maxcut2000.txt 2>/dev/null
if [ -z "$SKIP_FULL" ]; then
ncu --set full --clock-control none --import-source on -k regex:leaf_kernel -s 40 -c 2 -o gpurun_out/r1b_leaf -f python tools/ncu_target.py 2000 3 > gpurun_out/ncu_leaf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lzb_step_kernel|lz_coldot_kernel|lz_smem_kernel" -s 10 -c 3 -o gpurun_out/r1b_lanczos -f python tools/ncu_target.py 2000 3 > gpurun_out/ncu_lz.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_kernel -s 300 -c 3 -o gpurun_out/r1b_gemm -f python tools/ncu_target.py 2000 3 > gpurun_out/ncu_gemm.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep | tail -5
