mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2bm_bench_n2.json 2> gpurun_out/r2bm_bench_n2.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2bm_bench_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"])
for k, v in d.get("bnb", {}).items():
    print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "nodes", "counted", "not_converged", "ms_per_frontier")})
for k, v in d.get("sharded", {}).items():
    print(k, v["value"], v["ms_per_relaxation"])
P
