# Round 2, GPU call 1: first device run of the frontier batch kernels + every switch, sanitizer, ncu
set -x
mkdir -p gpurun_out
nproc > gpurun_out/r2_nproc.txt; nvidia-smi -L >> gpurun_out/r2_nproc.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -25 gpurun_out/r2_pytest_all.log
timeout 300 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-400 gpurun_out/r2_bench_n1.json
for inst in tt mkp small cls; do
  for sw in "" "SDPCUDA_BATCH_TINY=1" "SDPCUDA_BATCH_TINY=1 SDPCUDA_BATCH_SMEM=1" "SDPCUDA_BATCH_SMEM=1"; do
    tag=$(echo "$sw" | tr -d ' =1' | tr 'A-Z' 'a-z'); 
    env $sw timeout 200 python bench.py --workload frontier-example-$inst --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline \
      > gpurun_out/r2_fr_${inst}_batch_$tag.json 2>> gpurun_out/r2_fr.err
    cut -c1-200 gpurun_out/r2_fr_${inst}_batch_$tag.json
  done
  timeout 300 python bench.py --workload frontier-example-$inst --frontier-mode batch --nodes-per-gpu 592 > gpurun_out/r2_fr_${inst}_batch_cpu.json 2>> gpurun_out/r2_fr.err
  timeout 600 python bench.py --workload bnb-example-$inst --frontier-mode batch --native-nodes --objlimit --steps 2 --warmup 1 --no-cpu-baseline \
      > gpurun_out/r2_bnb_${inst}_batch_native.json 2>> gpurun_out/r2_bnb.err
  cut -c1-300 gpurun_out/r2_bnb_${inst}_batch_native.json
done
timeout 200 python bench.py --workload frontier-example-tt --frontier-mode serial --nodes-per-gpu 592 --no-cpu-baseline > gpurun_out/r2_fr_tt_serial.json 2>> gpurun_out/r2_fr.err
timeout 200 python bench.py --workload frontier-example-tt --frontier-mode threads --handles-per-gpu 16 --nodes-per-gpu 592 --no-cpu-baseline > gpurun_out/r2_fr_tt_threads.json 2>> gpurun_out/r2_fr.err
SDPCUDA_PACKED_SOLVE=1 timeout 300 python tests/tools/bnb_bench.py cuda > gpurun_out/r2_bnb_sdpi_packed.log 2>&1; tail -6 gpurun_out/r2_bnb_sdpi_packed.log
timeout 300 python tests/tools/bnb_bench.py cuda > gpurun_out/r2_bnb_sdpi_classic.log 2>&1; tail -6 gpurun_out/r2_bnb_sdpi_classic.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_batch.csv \
  python bench.py --workload frontier-example-tt --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline > gpurun_out/r2_ncu_batch_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_batch.csv > gpurun_out/r2_launches_batch.txt 2>/dev/null; cat gpurun_out/r2_launches_batch.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ipm_small_batch_kernel -c 1 -o gpurun_out/r2_ipm_small_batch -f \
  python bench.py --workload frontier-example-tt --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline > gpurun_out/r2_ncu_batch_full.log 2>&1
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_$tool.log python -m pytest tests/test_gpu_zfrontier.py -q -k "test_batched_nodes_match_oracle_and_single_solves and small" > gpurun_out/r2_sanitizer_$tool.out 2>&1
  tail -5 gpurun_out/r2_sanitizer_$tool.log
done
ls -la gpurun_out/
