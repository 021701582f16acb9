set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2ab_pytest_all.log 2>&1; tail -3 gpurun_out/r2ab_pytest_all.log
timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2ab_bench.json 2>> gpurun_out/r2ab_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2ab_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['device_ms_per_solve'], r['profiled_solve_ms'], r['share_of_step'])"
timeout 120 python tools/phase_probe.py maxcut2000 2>&1 | grep "phases" | tail -3
