set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2ae_pytest_all.log 2>&1; tail -3 gpurun_out/r2ae_pytest_all.log
for sw in "" "SDPCUDA_APAT=0"; do
env $sw timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2ae_bench.json 2>> gpurun_out/r2ae_bench.err; echo "$sw"; python -c "
import json; d=json.load(open('gpurun_out/r2ae_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['iterations_per_step'], d['objective'], r['frac'], r['device_ms_per_solve'], r['share_of_step'])"
done
timeout 120 python tools/phase_probe.py maxcut2000 2>&1 | grep "phases" | tail -3
