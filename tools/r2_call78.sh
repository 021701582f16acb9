# compute-sanitizer memcheck + racecheck of the one-launch kernels after the rework of the Schur phase (dense products, pair dots, ballot
# scan, LP rows per warp): batched nodes of example_CLS (dense path) and example_MkP (LP rows, packed factor)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --log-file gpurun_out/r2c_sanitizer_${tool}_batch_cls_mkp.log python -m pytest tests/test_gpu_zfrontier.py -q -k "test_batched_nodes_match_oracle_and_single_solves and (CLS or MkP)" > gpurun_out/r2c_sanitizer_${tool}.out 2>&1
  tail -2 gpurun_out/r2c_sanitizer_${tool}.out; tail -3 gpurun_out/r2c_sanitizer_${tool}_batch_cls_mkp.log
done
