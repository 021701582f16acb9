# Round 2, GPU call 16 (2 GPUs): the driver's multi-GPU commands — default bench under torchrun, reference arm under torchrun, 2-GPU tests
set -x
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2j_bench_n2.err; python - <<'P'
import json
d = json.load(open("gpurun_out/r2j_bench_n2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
for k, v in d.get("bnb", {}).items():
    print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "nodes", "counted", "ms_per_frontier", "not_converged", "rounds")}, v.get("max_rel_diff_to_oracle"))
print(json.dumps(d.get("sharded"), indent=0)[:1500])
P
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2j_bench_ref_n2.json 2> gpurun_out/r2j_bench_ref_n2.err ) 2>&1 | tail -3
cut -c1-300 gpurun_out/r2j_bench_ref_n2.json; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_ref_n2.json')); print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:120])"
timeout 600 python -m pytest tests/test_gpu_solver.py -q -k "shard or nccl or two" 2>&1 | tail -4
