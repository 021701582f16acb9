set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest_all.log 2>&1; tail -3 gpurun_out/r2y_pytest_all.log
for sw in "" "SDPCUDA_CHOL_CHAIN=0"; do
env $sw timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2y_bench.json 2>> gpurun_out/r2y_bench.err; echo "$sw"; python -c "
import json; d=json.load(open('gpurun_out/r2y_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['profiled_solve_ms'], r['share_of_step'], d['kernels']['potrf_2000']['ms'], d['kernels']['potrf_with_inverse_2000']['ms'])"
done
tail -3 gpurun_out/r2y_bench.err
timeout 120 python tools/leaf_probe.py 64 128 256 1501 2>&1 | tail -6
