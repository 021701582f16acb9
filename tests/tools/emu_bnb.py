"""Complete frontier-synchronous B&B trees of example_TT and example_MkP with every round executed by the CPU emulation of the batch
kernel (256-thread instantiation, work space staged in shared memory, the library's own batch plan).  Too slow for the test suite
(2 and 14 minutes); result of 2026-10-17: TT optimal 2.118035 (645 nodes, 19 launches), MkP optimal -94.99996 (553 nodes, 14 launches),
no unsolved node."""
import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import ctypes as C, numpy as np
import test_cuemu_batch as T
from scip_sdp_b200 import abi, misdp, frontier
L = abi.Lib(abi.PRODUCT_LIB)
emu = C.CDLL(os.path.join(T.EMUDIR, "_build", "libcuemu_ipm.so"))
for f in (emu.cuemu_run_small_batch, emu.cuemu_run_tiny_batch): f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
class S:
    def __init__(s): s.cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB)); s.nb = 0; s.nn = 0
    def solve_batch(s, probs, fetch=True, **kw):
        got, nb, nt = T.run_planned_batch(L, emu, probs, True, stage=True, **kw)
        s.nb += 1; s.nn += len(probs)
        assert nb == len(probs) == nt
        return [got[i] for i in range(len(probs))]
    def solve(s, fp, **kw): return s.cpu.solve(fp, **kw)
    def get_y(s): return s.cpu.get_y()
for name, want in (("example_TT.dat-s.gz", 2.11803), ("example_MkP.dat-s.gz", -95.0)):
    M = misdp.read_instance(os.path.join(T.GOLDEN, name))
    s = S(); t = time.time()
    r = frontier.branch_and_bound(s, M, mode="batch", width=592, gaptol=1e-5, feastol=1e-5, timelimit=3000)
    print(name, r["status"], M.file_objective(r["objval"]), want, "nodes", r["nodes"], "rounds", r["rounds"], "unsolved", r["unsolved"], f"{time.time()-t:.0f}s", flush=True)
