#!/bin/bash
# The native node marshalling (csrc/node_marshal.hpp) and the batch stand-in of the checker library under AddressSanitizer + UBSan:
#   bash tests/tools/sanitize_node_marshal.sh
# Builds a sanitized copy of oracle/ipm_oracle.cpp (which includes the shared header) in /tmp and runs the array-equality cases,
# sdpcuda_solve_nodes and native B&B trees on it.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=/tmp/sanitize_node_marshal
mkdir -p $OUT && cd $OUT
SCIPYLIB=$(python -c "import scipy, os; print(os.path.join(os.path.dirname(scipy.__file__), '..', 'scipy.libs'))")
OB=$(ls $SCIPYLIB/libscipy_openblas*.so | head -1)
g++ -O1 -g -fPIC -shared -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -I$ROOT/oracle -o liboracle_san.so $ROOT/oracle/ipm_oracle.cpp $OB -Wl,-rpath,$SCIPYLIB
cat > run.py <<PY
import os, sys
sys.path.insert(0, '$ROOT'); sys.path.insert(0, '$ROOT/tests')
import numpy as np
from scip_sdp_b200 import abi, misdp, frontier
abi.ORACLE_LIB = abi.PRODUCT_LIB = "$OUT/liboracle_san.so"
import test_node_marshal as t
for name, make in t._models():
    t.test_native_node_problem_equals_the_python_restatement(name, make)
t.test_solve_nodes_equals_the_python_path()
for name in ["example_small.dat-s", "example_TT.dat-s.gz", "example_MkP.dat-s.gz", "example_small_ind.dat-s"]:
    M = misdp.read_instance(os.path.join('$ROOT/tests/golden', name))
    r = frontier.branch_and_bound(abi.Solver(abi.Lib(abi.ORACLE_LIB)), M, mode="batch", width=64, native=True, use_objlimit=True)
    print(name, r["status"], r["objval"], r["nodes"])
print("SANITIZER RUN CLEAN (address,undefined)")
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 python run.py
