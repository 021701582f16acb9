set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solver.py -m gpu -q -x 2>&1 | tail -2
for sw in "" "SDPCUDA_GEMM_BALANCE=0"; do
env $sw timeout 300 python bench.py --no-nodes --no-cpu-baseline > gpurun_out/r2ai_bench.json 2>> gpurun_out/r2ai_bench.err; echo "$sw"; python -c "
import json; d=json.load(open('gpurun_out/r2ai_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['iterations_per_step'], d['objective'], r['frac'], r['device_ms_per_solve'])"
done
for sw in "" "SDPCUDA_GEMM_BALANCE=0"; do env $sw timeout 120 python tools/phase_probe.py maxcut2000 2>&1 | grep "phases" | tail -2; done
