#!/bin/bash
# Host half of sdpcuda_solve_batch / sdpcuda_solve_nodes in the PRODUCT library under AddressSanitizer + UBSan (no GPU needed: only the
# packing hooks and the node marshalling are called):  bash tests/tools/sanitize_batch_host.sh
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=/tmp/sanitize_batch_host
mkdir -p $OUT && cd $OUT
FL="-gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fno-omit-frame-pointer"
for f in gemm chol eig ops ipm_small ipm_tiny ipm; do nvcc $FL -c $ROOT/scip-sdp_b200/csrc/$f.cu -o $f.o & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libsdpcuda_san.so *.o -cudart static -ldl -Xlinker -lasan -Xlinker -lubsan
cat > run.py <<PY
import os, sys
sys.path.insert(0, '$ROOT'); sys.path.insert(0, '$ROOT/tests')
from scip_sdp_b200 import abi
abi.PRODUCT_LIB = "$OUT/libsdpcuda_san.so"
import ctypes as C
import test_batch_pack as bp, test_cuemu_batch as ce, test_node_marshal as nm
L = bp.lib.__wrapped__() if hasattr(bp.lib, "__wrapped__") else None
lib = abi.Lib(abi.PRODUCT_LIB)
lib.lib.sdpcuda_debug_pack_node.argtypes = [C.POINTER(abi.Problem), C.POINTER(abi.Params), C.c_ulonglong, C.c_ulonglong, C.c_ulonglong,
                                            C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                            C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
for name, make in bp._cases():
    bp.test_packed_node_reproduces_the_operators(lib, name, make)
bp.test_relaxations_outside_the_single_cta_limits_do_not_fit(lib)
emu = C.CDLL(os.path.join(ce.EMUDIR, "_build", "libcuemu_ipm.so"))
for f in (emu.cuemu_run_small_batch, emu.cuemu_run_tiny_batch):
    f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
for t in (False, True):
    ce.test_planned_batch_with_mixed_sizes(lib, emu, t)
    ce.test_work_space_staged_in_shared_memory(lib, emu, t)
    ce.test_staged_multipliers_are_copied_back_for_a_packed_single_solve(lib, emu, t)
for name, make in nm._models():
    nm.test_native_node_problem_equals_the_python_restatement(name, make)
print("SANITIZER RUN CLEAN (product library host code: address,undefined)")
PY
make -C $ROOT/tests/harness/cuemu >/dev/null
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 python run.py
