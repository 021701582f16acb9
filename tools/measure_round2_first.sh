# First GPU call of round 2 (one B200, ~20 minutes; every python start pays ~10 s of torch import): bash tools/measure_round2_first.sh
# 1. the whole GPU suite (the last file, tests/test_gpu_zfrontier.py, holds everything written after the GPU budget of round 1 ran out)
# 2. the headline bench (the upload sort changed: e2e should lose ~10 ms per solve)
# 3. the frontier of small nodes: node-by-node vs threads vs ONE launch (sdpcuda_solve_batch), and complete B&B runs
# 4. mid-size frontier nodes with several handles per GPU
# 5. ncu: launch list of a batched frontier and one full capture of ipm_small_batch_kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_all.log 2>&1; tail -15 gpurun_out/r2_pytest_all.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-400 gpurun_out/r2_bench_n1.json
for inst in tt cls mkp small; do
  for mode in serial threads batch; do
    timeout 300 python bench.py --workload frontier-example-$inst --frontier-mode $mode --nodes-per-gpu 592 --handles-per-gpu 16 \
      > gpurun_out/r2_frontier_example_${inst}_$mode.json 2>> gpurun_out/r2_frontier_example.err
    cut -c1-160 gpurun_out/r2_frontier_example_${inst}_$mode.json
  done
  SDPCUDA_BATCH_TINY=1 timeout 300 python bench.py --workload frontier-example-$inst --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline \
      > gpurun_out/r2_frontier_example_${inst}_batch_tiny.json 2>> gpurun_out/r2_frontier_example.err
  cut -c1-160 gpurun_out/r2_frontier_example_${inst}_batch_tiny.json
  SDPCUDA_BATCH_TINY=1 SDPCUDA_BATCH_SMEM=1 timeout 300 python bench.py --workload frontier-example-$inst --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline \
      > gpurun_out/r2_frontier_example_${inst}_batch_tiny_smem.json 2>> gpurun_out/r2_frontier_example.err
  cut -c1-160 gpurun_out/r2_frontier_example_${inst}_batch_tiny_smem.json
  SDPCUDA_BATCH_SMEM=1 timeout 300 python bench.py --workload frontier-example-$inst --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline \
      > gpurun_out/r2_frontier_example_${inst}_batch_smem.json 2>> gpurun_out/r2_frontier_example.err
  cut -c1-160 gpurun_out/r2_frontier_example_${inst}_batch_smem.json
  timeout 600 python bench.py --workload bnb-example-$inst --frontier-mode batch --steps 2 --warmup 1 > gpurun_out/r2_bnb_example_${inst}_batch.json 2>> gpurun_out/r2_bnb.err
  cut -c1-300 gpurun_out/r2_bnb_example_${inst}_batch.json
  timeout 600 python bench.py --workload bnb-example-$inst --frontier-mode batch --native-nodes --objlimit --steps 2 --warmup 1 --no-cpu-baseline \
      > gpurun_out/r2_bnb_example_${inst}_batch_native.json 2>> gpurun_out/r2_bnb.err
  cut -c1-300 gpurun_out/r2_bnb_example_${inst}_batch_native.json
done
for w in frontier-tt500 frontier-cls frontier-mkp60; do
  for k in 1 4; do
    timeout 600 python bench.py --workload $w --frontier-mode threads --handles-per-gpu $k --nodes-per-gpu 32 --no-cpu-baseline \
      > gpurun_out/r2_${w}_threads$k.json 2>> gpurun_out/r2_frontier_threads.err
    cut -c1-160 gpurun_out/r2_${w}_threads$k.json
  done
done
# single small relaxations through the binding: classic upload vs packed path (+ staged work space)
python tests/tools/bnb_bench.py > gpurun_out/r2_bnb_sdpi_classic.log 2>&1; tail -6 gpurun_out/r2_bnb_sdpi_classic.log
SDPCUDA_PACKED_SOLVE=1 python tests/tools/bnb_bench.py cuda > gpurun_out/r2_bnb_sdpi_packed.log 2>&1; tail -6 gpurun_out/r2_bnb_sdpi_packed.log
SDPCUDA_PACKED_SOLVE=1 SDPCUDA_BATCH_SMEM=1 SDPCUDA_BATCH_TINY=1 python tests/tools/bnb_bench.py cuda > gpurun_out/r2_bnb_sdpi_packed_smem.log 2>&1; tail -6 gpurun_out/r2_bnb_sdpi_packed_smem.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_batch.csv \
  python bench.py --workload frontier-example-tt --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline > gpurun_out/r2_ncu_batch_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_batch.csv > gpurun_out/r2_launches_batch.txt 2>/dev/null; cat gpurun_out/r2_launches_batch.txt
ncu --set full --clock-control none --import-source on -k regex:ipm_small_batch_kernel -c 1 -o gpurun_out/r2_ipm_small_batch -f \
  python bench.py --workload frontier-example-tt --frontier-mode batch --nodes-per-gpu 592 --no-cpu-baseline > gpurun_out/r2_ncu_batch_full.log 2>&1
ls -la gpurun_out/r2_*
