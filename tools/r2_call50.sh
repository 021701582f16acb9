set -x
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2ak_bench_n2.json 2> gpurun_out/r2ak_bench_n2.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2ak_bench_n2.err; python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2ak_bench_n2.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["roofline"]["frac"])
for k, v in d.get("bnb", {}).items():
    print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "nodes", "counted", "ms_per_frontier", "device_nodes_per_s", "not_converged", "rounds")}, v.get("max_rel_diff_to_oracle"))
print(json.dumps(d.get("sharded"))[:1200])
P
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "two_gpus" 2>&1 | tail -2
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2ak_bench_ref_n2.json 2> gpurun_out/r2ak_bench_ref_n2.err ) 2>&1 | tail -3
cut -c1-300 gpurun_out/r2ak_bench_ref_n2.json
