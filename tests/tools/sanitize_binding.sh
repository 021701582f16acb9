#!/bin/bash
# Runs the binding sdpisolver_cuda.c (+ the reference's sdpi.c layer compiled in place + the CPU oracle back end) under
# AddressSanitizer or UBSan:  bash tests/tools/sanitize_binding.sh [address|undefined]
# Build products go to /tmp; needs /root/reference (build container only).  Covers the ported checksdpi cases, the boundary
# cases (penalty patterns, primal getters, warm start / preoptimal), the sdpi-level warm start at a node with removed block rows
# and four B&B runs.
set -e
KIND=${1:-address}
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
REF=${REF:-/root/reference}
OUT=/tmp/sanitize_$KIND
mkdir -p $OUT && cd $OUT
INC="-I$ROOT/shim -I$REF/src -I$REF/src/scipsdp -I$ROOT/include"
FL="-O1 -g -fPIC -fsanitize=$KIND -fno-omit-frame-pointer"
for f in sdpi/sdpi.c sdpi/sdpsolchecker.c sdpi/solveonevarsdp.c sdpi/sdpiclock.c scipsdp/SdpVarfixer.c; do gcc $FL $INC -c $REF/src/$f -o $(basename $f .c).o; done
gcc $FL $INC -c $ROOT/scip-sdp_b200/sdpi/sdpisolver_cuda.c -o sdpisolver_cuda.o
gcc $FL $INC -c $ROOT/shim/bms_shim.c -o bms_shim.o
cp $ROOT/oracle/_ref/obj/lapack_interface.o .
SCIPYLIB=$(python -c "import scipy, os; print(os.path.join(os.path.dirname(scipy.__file__), '..', 'scipy.libs'))")
OB=$(ls $SCIPYLIB/libscipy_openblas*.so | head -1)
gcc -shared -fsanitize=$KIND -o libsdpi_oracle_san.so *.o -L$ROOT/oracle -loracle_sdp -Wl,-rpath,$ROOT/oracle $OB -Wl,-rpath,$SCIPYLIB -lm
cat > run.py <<PY
import os, sys
sys.path.insert(0, '$ROOT'); sys.path.insert(0, '$ROOT/tests')
os.environ["SHIM_QUIET"] = "1"
from harness import sdpi_ref, boundary_cases, checksdpi_port, bnb
from golden.checksdpi_cases import CASES
from scip_sdp_b200 import misdp
SAN = "$OUT/libsdpi_oracle_san.so"
sdpi_ref.LIB_ORACLE = SAN
lib = sdpi_ref.SdpiLib(SAN)
for name in sorted(CASES):
    checksdpi_port.run_case(lib, CASES[name], name)
boundary_cases.run_penalty_patterns(SAN); boundary_cases.run_primal_getters(SAN); boundary_cases.run_warmstart_and_preoptimal(SAN)
import test_oracle_golden as t
for nf in (0, 20):
    t.test_warmstart_and_preoptimal_through_the_reference_sdpi_layer(lib, nf)
for name in ["example_small.dat-s", "example_MkP.dat-s.gz", "example_cbf_mix.cbf", "example_tightenmatrices.dat-s"]:
    r = bnb.solve_misdp(lib, misdp.read_instance(os.path.join('$ROOT/tests/golden', name)), timelimit=300)
    print(name, r["status"], r["objval"], r["nodes"])
print("SANITIZER RUN CLEAN ($KIND)")
PY
LIBSAN=$(gcc -print-file-name=$([ "$KIND" = address ] && echo libasan.so || echo libubsan.so))
LD_PRELOAD=$LIBSAN ASAN_OPTIONS=detect_leaks=0 python run.py
