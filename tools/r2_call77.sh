set -x
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2bf_bench_n8.json 2> gpurun_out/r2bf_bench_n8.err ) 2>&1 | tail -3
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2bf_bench_n8.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
for k, v in d.get("bnb", {}).items():
    print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "nodes", "counted", "not_converged", "ms_per_frontier")})
for k, v in d.get("sharded", {}).items():
    print(k, v["value"], v["ms_per_relaxation"], v["rel_diff_to_oracle"])
P
