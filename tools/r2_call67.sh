# Round 2, end of session 2: ncu launch list of ONE max-cut 2000 solve with the final code (same command as tools/r2_call39.sh).
# Numbers under ncu are never bench values: per-launch times are cold-cache and serialised, the SHARES are what is compared.
set -x
mkdir -p gpurun_out
cat > /tmp/one_solve.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
from scip_sdp_b200 import abi, generators
fp, _ = generators.maxcut(2000, 0.01, seed=4004).flatten()
g = abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)
r = g.solve(fp, fetch=False, gaptol=1e-5, feastol=1e-5, absgaptol=5e-6)
print(r["phase_name"], r["iterations"], r["launches"], r["device_ms"])
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c_launches_maxcut2000.csv python /tmp/one_solve.py > gpurun_out/r2c_ncu_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c_launches_maxcut2000.csv > gpurun_out/r2c_launches_maxcut2000.txt 2>/dev/null; head -24 gpurun_out/r2c_launches_maxcut2000.txt
tail -2 gpurun_out/r2c_ncu_list.log
