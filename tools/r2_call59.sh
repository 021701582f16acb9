set -x
mkdir -p gpurun_out
for L in 1 2 3 4; do
echo "== SDPCUDA_LONER_LANES=$L"; SDPCUDA_LONER_LANES=$L timeout 600 python tools/frontier_rates.py TT-500 CLS-syn
done > gpurun_out/r2ao_loner_lanes.log 2>&1
cat gpurun_out/r2ao_loner_lanes.log
