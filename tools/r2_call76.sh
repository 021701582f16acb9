mkdir -p gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2bk_bench_n1.json 2> gpurun_out/r2bk_bench_n1.err ) 2>&1 | tail -3
python - <<'P'
import json
d = json.load(open("gpurun_out/r2bk_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value"), d["clocks"])
for k, v in d.get("bnb", {}).items():
    print(k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "nodes", "counted", "ms_per_frontier", "device_nodes_per_s", "not_converged", "rounds")}, v.get("max_rel_diff_to_oracle"), "cpu", round(v.get("cpu_baseline", {}).get("value", 0), 2))
P
