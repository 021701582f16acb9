set -x
mkdir -p gpurun_out
for T in 1 8; do
for W in example_TT example_CLS; do
SDPCUDA_UPLOAD_PROFILE=1 timeout 300 python tools/concurrent_sdpi_probe.py $W 32 $T > gpurun_out/r2al_up_${W}_$T.out 2> gpurun_out/r2al_up_${W}_$T.err
cat gpurun_out/r2al_up_${W}_$T.out
python - <<P
import collections
d = collections.defaultdict(list)
for l in open("gpurun_out/r2al_up_${W}_$T.err"):
    if l.startswith("[upload]"):
        p = l[8:].rsplit(None, 2)
        d[p[0].strip()].append(float(p[1]))
tot = 0
for k, v in d.items():
    v = v[len(v)//3:]
    print(f"  {k:30s} n={len(v):4d} mean {sum(v)/len(v):8.3f} ms max {max(v):8.3f}")
    tot += sum(v)/len(v)
print("  total mean", tot)
P
done; done
