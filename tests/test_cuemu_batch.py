"""CPU suite, part 6: the frontier-batch kernel executed on the CPU.  tests/harness/cuemu compiles the SOURCE of csrc/ipm_small.cu
(both instantiations) against a fiber-based emulation of a CUDA thread block (every thread a fiber, __syncthreads / warp shuffles as
scheduler barriers); the descriptors and the packed image come from the product library's own packing code
(sdpcuda_debug_pack_node, the host half of sdpcuda_solve_batch) bound to host buffers.  Together this runs the whole batch path —
packing, descriptor staging in shared memory, in-kernel cold start and dense-matrix expansion, the interior-point iteration, the
result and y write-back — without a GPU, and compares with the CPU oracle at the north-star tolerance (1e-5 relative).
It checks logic, not timing or races; the GPU suite (tests/test_gpu_zfrontier.py) remains the parity test proper."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from scip_sdp_b200 import abi, generators, misdp
from test_batch_pack import SmallArgs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
EMUDIR = os.path.join(os.path.dirname(__file__), "harness", "cuemu")
KW = dict(gaptol=1e-6, feastol=1e-6)


class SmallResult(C.Structure):      # sdpk::SmallResult (csrc/ipm_small.cuh)
    _fields_ = [("phase", C.c_int), ("stop", C.c_int), ("iterations", C.c_int), ("backtracks", C.c_int),
                ("pobj", C.c_double), ("dobj", C.c_double), ("relgap", C.c_double), ("pinf", C.c_double), ("dinf", C.c_double), ("mu", C.c_double)]


@pytest.fixture(scope="module")
def emu():
    r = subprocess.run(["make", "-C", EMUDIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    L = C.CDLL(os.path.join(EMUDIR, "_build", "libcuemu_ipm.so"))
    for f in (L.cuemu_run_small_batch, L.cuemu_run_tiny_batch):
        f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
    return L


@pytest.fixture(scope="module")
def lib():
    L = abi.Lib(abi.PRODUCT_LIB)
    L.lib.sdpcuda_debug_pack_node.argtypes = [C.POINTER(abi.Problem), C.POINTER(abi.Params), C.c_ulonglong, C.c_ulonglong, C.c_ulonglong,
                                              C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                              C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    return L


CANARY = 1.2345e300


def run_batch(lib, emu, probs, tiny=False, **kw):
    """the device half of sdpcuda_solve_batch on the emulator: -> list of dict(phase_name, dobj, ..., y)"""
    par = lib.default_params(**kw)
    n = len(probs)
    descs = (SmallArgs * n)()
    res = (SmallResult * n)()
    keep = []
    for i, fp in enumerate(probs):
        st = fp.struct()
        nimg, nwork, ndesc, fits = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_int(-1)
        assert lib.lib.sdpcuda_debug_pack_node(C.byref(st), C.byref(par), 0, 0, 0, None, 0, C.byref(nimg), C.byref(nwork), None, 0,
                                               C.byref(ndesc), C.byref(fits)) == 0
        assert fits.value == 1 and ndesc.value == C.sizeof(SmallArgs)
        img = np.zeros(nimg.value + 64, dtype=np.uint8)
        work = np.zeros(nwork.value + 64)
        work[nwork.value:] = CANARY
        y = np.full(fp.m + 17, CANARY)
        assert lib.lib.sdpcuda_debug_pack_node(C.byref(st), C.byref(par), img.ctypes.data, work.ctypes.data, y.ctypes.data, img.ctypes.data,
                                               img.size, C.byref(nimg), C.byref(nwork), C.byref(descs[i]), C.sizeof(SmallArgs),
                                               C.byref(ndesc), C.byref(fits)) == 0
        descs[i].out = C.addressof(res[i])
        keep.append((img, work, y, nwork.value))
    run = emu.cuemu_run_tiny_batch if tiny else emu.cuemu_run_small_batch
    assert run(n, C.addressof(descs), C.sizeof(SmallArgs), 0) == 0, "emulator reported a deadlock or a shared-memory overrun"
    out = []
    for i, fp in enumerate(probs):
        img, work, y, nw = keep[i]
        assert np.all(work[nw:] == CANARY), "the kernel wrote behind the node's work space"
        assert np.all(y[fp.m + 1:] == CANARY), "the kernel wrote behind the node's y"
        r = res[i]
        out.append(dict(phase_name=abi.PHASES[r.phase], stop_name=abi.STOPS[r.stop], iterations=r.iterations, pobj=r.pobj, dobj=r.dobj,
                        relgap=r.relgap, pinf=r.pinf, dinf=r.dinf, y=y[:fp.m].copy()))
    return out


def _nodes(M, q):
    ints = np.flatnonzero(M.integer)[:q]
    for code in range(1 << q):
        lb, ub = M.lb.copy(), M.ub.copy()
        for b, j in enumerate(ints):
            lb[j] = ub[j] = float((code >> b) & 1)
        fp, info = M.flatten_fast(lb, ub)
        if fp.m > 0:
            yield fp


def _compare(fp, r, ref):
    assert (r["phase_name"] == "pdOPT") == (ref["phase_name"] == "pdOPT"), (r["phase_name"], r["stop_name"], ref["phase_name"])
    if ref["phase_name"] in ("pFEAS_dINF", "dINF"):
        assert r["phase_name"] in ("pFEAS_dINF", "dINF")
    if ref["phase_name"] == "pdOPT":
        assert abs(r["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
        assert np.allclose(r["y"], ref["y"], atol=1e-3 * max(1.0, np.abs(ref["y"]).max()))
        assert r["relgap"] <= 1e-5 and r["pinf"] <= 1e-5 and r["dinf"] <= 1e-5


@pytest.mark.parametrize("tiny", [False, True], ids=["1024-threads", "256-threads"])
@pytest.mark.parametrize("name,q", [("example_small.dat-s", 2), ("example_TT.dat-s.gz", 2), ("example_MkP.dat-s.gz", 1)])
def test_batch_kernel_on_the_emulator(lib, emu, name, q, tiny):
    """nodes of the shipped instances with blocks of order <= 16 through both instantiations of the batch kernel"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, name)).rows_to_bounds()
    probs = list(_nodes(M, q))
    got = run_batch(lib, emu, probs, tiny=tiny, **KW)
    cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    for fp, r in zip(probs, got):
        _compare(fp, r, cpu.solve(fp, **KW))


def test_batch_kernel_with_dense_constraint_matrices_on_the_emulator(lib, emu):
    """a small cardinality-constrained least-squares relaxation: dense constraint matrices expanded by the CTA itself, block of
    order 13, and a max-cut relaxation without LP block"""
    cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    probs = [generators.cls(12, 9, 3, seed=5).flatten()[0], generators.maxcut(24, 0.3, seed=3).flatten()[0]]
    for fp, r in zip(probs, run_batch(lib, emu, probs, **KW)):
        _compare(fp, r, cpu.solve(fp, **KW))


def test_root_relaxations_of_the_larger_shipped_instances_on_the_emulator(lib, emu):
    """example_CLS (block of order 43, dense constraint matrices) and example_MkP (m = 105 > 64: Schur factor in global memory) in
    ONE emulated launch of two CTAs: same iteration counts as the oracle, objectives within 1e-7 relative"""
    cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    probs = [misdp.read_sdpa(os.path.join(GOLDEN, f)).rows_to_bounds().flatten()[0] for f in ("example_CLS.dat-s.gz", "example_MkP.dat-s.gz")]
    for fp, r in zip(probs, run_batch(lib, emu, probs, **KW)):
        ref = cpu.solve(fp, **KW)
        _compare(fp, r, ref)
        assert abs(r["iterations"] - ref["iterations"]) <= 1 and abs(r["dobj"] - ref["dobj"]) <= 1e-7 * max(1.0, abs(ref["dobj"]))


class EmuSolver:
    """stands in for abi.Solver in frontier.branch_and_bound: batches run on the emulated kernel, single solves (stable-settings
    retries, penalty ladder) on the oracle"""

    def __init__(self, lib, emu, tiny):
        self.lib, self.emu, self.tiny = lib, emu, tiny
        self.cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
        self.batches = 0

    def solve_batch(self, probs, fetch=True, objlimits=None, **kw):
        assert objlimits is None
        self.batches += 1
        return run_batch(self.lib, self.emu, probs, tiny=self.tiny, **kw)

    def solve(self, fp, **kw):
        return self.cpu.solve(fp, **kw)

    def get_y(self):
        return self.cpu.get_y()


@pytest.mark.parametrize("name,want", [("example_small.dat-s", -8.0), ("example_inf.dat-s", None), ("example_small_ind.dat-s", -18.0)])
def test_branch_and_bound_on_the_emulated_batch_kernel(lib, emu, name, want):
    """complete frontier-synchronous trees with every round solved by one emulated launch of the 256-thread instantiation"""
    from scip_sdp_b200 import frontier
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    s = EmuSolver(lib, emu, tiny=True)
    r = frontier.branch_and_bound(s, M, mode="batch", width=64, gaptol=1e-5, feastol=1e-5)
    assert s.batches == r["rounds"] or s.batches <= r["rounds"]
    if want is None:
        assert r["status"] == "infeasible"
    else:
        assert r["status"] == "optimal" and abs(M.file_objective(r["objval"]) - want) <= 1e-4 * max(1.0, abs(want))


@pytest.mark.parametrize("tiny", [False, True], ids=["1024-threads", "256-threads"])
def test_result_does_not_depend_on_the_thread_visiting_order(lib, emu, tiny, monkeypatch):
    """the emulator visits warps and lanes in ascending or (CUEMU_REVERSE=1) descending order between barriers: bit-identical results
    mean that no two threads exchange data through memory without a barrier in between on these inputs"""
    probs = [misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds().flatten()[0], generators.cls(12, 9, 3, seed=5).flatten()[0]]
    fwd = run_batch(lib, emu, probs, tiny=tiny, **KW)
    monkeypatch.setenv("CUEMU_REVERSE", "1")
    rev = run_batch(lib, emu, probs, tiny=tiny, **KW)
    for a, b in zip(fwd, rev):
        assert a["phase_name"] == b["phase_name"] and a["iterations"] == b["iterations"]
        assert a["dobj"] == b["dobj"] and a["pobj"] == b["pobj"] and np.array_equal(a["y"], b["y"])


def run_planned_batch(lib, emu, probs, usetiny, stage=False, copyback=False, **kw):
    """exactly what sdpcuda_solve_batch does, with the CUDA calls replaced: ONE image, ONE zeroed work buffer, ONE y buffer and the
    descriptor order come from the library's own plan (sdpcuda_debug_pack_batch); the two launches run on the emulator"""
    par = lib.default_params(**kw)
    n = len(probs)
    F = lib.lib.sdpcuda_debug_pack_batch
    F.argtypes = [C.c_int, C.POINTER(C.POINTER(abi.Problem)), C.POINTER(abi.Params), C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong,
                  C.c_ulonglong, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p,
                  C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    flags = int(usetiny) + 2 * int(stage) + 4 * int(copyback)
    stagebytes = (C.c_size_t * 2)()
    structs = [p.struct() for p in probs]
    ps = (C.POINTER(abi.Problem) * n)(*[C.pointer(s) for s in structs])
    nimg, nwork, ny, nb, nt = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_int(0), C.c_int(0)
    assert F(n, ps, C.byref(par), flags, 0, 0, 0, 0, None, 0, C.byref(nimg), C.byref(nwork), C.byref(ny), None, 0, C.byref(nb), C.byref(nt), None, None, None) == 0
    img = np.zeros(nimg.value + 64, dtype=np.uint8)
    work = np.zeros(nwork.value + 64); work[nwork.value:] = CANARY
    ybuf = np.full(ny.value + 64, CANARY)
    res = (SmallResult * max(nb.value, 1))()
    descs = (SmallArgs * max(nb.value, 1))()
    owner = (C.c_int * max(nb.value, 1))()
    yoff = (C.c_size_t * max(nb.value, 1))()
    assert F(n, ps, C.byref(par), flags, img.ctypes.data, work.ctypes.data, ybuf.ctypes.data, C.addressof(res), img.ctypes.data, img.size,
             C.byref(nimg), C.byref(nwork), C.byref(ny), C.byref(descs), C.sizeof(descs), C.byref(nb), C.byref(nt), owner, yoff, stagebytes) == 0
    if not any(p.m > 64 for p in probs):           # (Schur complements above 64 reserve room for their packed factor as well)
        assert (stagebytes[0] > 0 or stagebytes[1] > 0) == (stage and nb.value > 0)
    base = C.addressof(descs)
    if nt.value:
        assert emu.cuemu_run_tiny_batch(nt.value, base, C.sizeof(SmallArgs), stagebytes[0]) == 0
    if nb.value - nt.value:
        assert emu.cuemu_run_small_batch(nb.value - nt.value, base + nt.value * C.sizeof(SmallArgs), C.sizeof(SmallArgs), stagebytes[1]) == 0
    assert np.all(work[nwork.value:] == CANARY) and np.all(ybuf[ny.value:] == CANARY)
    out = {}
    for k in range(nb.value):
        fp, r = probs[owner[k]], res[k]
        out[owner[k]] = dict(phase_name=abi.PHASES[r.phase], stop_name=abi.STOPS[r.stop], iterations=r.iterations, pobj=r.pobj, dobj=r.dobj,
                             relgap=r.relgap, pinf=r.pinf, dinf=r.dinf, y=ybuf[yoff[k]:yoff[k] + fp.m].copy())
    def arr(addr, n):
        lo = (addr - work.ctypes.data) // 8
        return work[lo:lo + n].copy()
    for k in range(nb.value):
        # descriptor k is in launch order; its result slot tells which node it describes
        d = descs[k]
        node = (d.out - C.addressof(res)) // C.sizeof(SmallResult)
        out[owner[node]].update(X=arr(d.X, d.arena), S=arr(d.S, d.arena), x=arr(d.x, d.nlp), s=arr(d.s, d.nlp))
        if stage and not copyback:
            # staged arrays never reach the global work space: the staged head of the node's slice is still all zeros
            lo = (d.workbase - work.ctypes.data) // 8
            assert d.stage_doubles > 0 and not work[lo:lo + d.stage_doubles].any()
    return out, nb.value, nt.value


@pytest.mark.parametrize("usetiny", [False, True])
def test_planned_batch_with_mixed_sizes(lib, emu, usetiny):
    """one plan for nodes of different shapes — blocks of order 2, 10, 13 (dense) and 24, one relaxation outside the limits —
    executed as the two launches of sdpcuda_solve_batch: every batched node matches the oracle, the large one is left to the
    ordinary solve, and with SDPCUDA_BATCH_TINY the three relaxations with blocks <= 16 go first"""
    S = misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s")).rows_to_bounds().flatten()[0]
    T = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds().flatten()[0]
    probs = [generators.maxcut(24, 0.3, seed=3).flatten()[0], S, generators.maxcut(96, 0.1, seed=7).flatten()[0], T,
             generators.cls(12, 9, 3, seed=5).flatten()[0], S]
    got, nbatched, ntiny = run_planned_batch(lib, emu, probs, usetiny, **KW)
    assert nbatched == 5 and sorted(got) == [0, 1, 3, 4, 5] and ntiny == (4 if usetiny else 0)
    cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    for i, r in got.items():
        _compare(probs[i], r, cpu.solve(probs[i], **KW))
    assert got[1]["dobj"] == got[5]["dobj"] and np.array_equal(got[1]["y"], got[5]["y"])


@pytest.mark.parametrize("usetiny", [False, True])
def test_work_space_staged_in_shared_memory(lib, emu, usetiny):
    """SDPCUDA_BATCH_SMEM: the entry kernel moves the head of every node's work space (vectors, block matrices, M and its factor as far
    as the budget reaches) into shared memory and redirects the descriptor's pointers; same results as from global memory bit for
    bit, nothing of the staged head is ever written to the global work space, no overrun of the enlarged shared memory"""
    T = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds().flatten()[0]       # everything fits
    K = misdp.read_sdpa(os.path.join(GOLDEN, "example_MkP.dat-s.gz")).rows_to_bounds().flatten()[0]      # M (2 x 88 KB) stays global
    Cl = generators.cls(12, 9, 3, seed=5).flatten()[0]                                                   # dense-path buffers
    W = generators.maxcut(40, 0.2, seed=3).flatten()[0]                                                  # 1024-thread kernel, partly staged
    probs = [T, K, Cl, W]
    plain, nb, nt = run_planned_batch(lib, emu, probs, usetiny, stage=False, **KW)
    staged, nb2, nt2 = run_planned_batch(lib, emu, probs, usetiny, stage=True, **KW)
    assert (nb, nt) == (nb2, nt2) == (4, 3 if usetiny else 0)
    for i in range(4):
        a, b = plain[i], staged[i]
        assert a["phase_name"] == b["phase_name"] == "pdOPT" and a["iterations"] == b["iterations"]
        assert a["dobj"] == b["dobj"] and np.array_equal(a["y"], b["y"])


@pytest.mark.parametrize("usetiny", [False, True])
def test_staged_multipliers_are_copied_back_for_a_packed_single_solve(lib, emu, usetiny):
    """descriptor flag copyback (SDPCUDA_PACKED_SOLVE: one relaxation through the packed path, its X, S, x, s served by the getters
    afterwards): the staged copies reach their global addresses, bit-identical to the run without staging"""
    probs = [misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds().flatten()[0], generators.maxcut(40, 0.2, seed=3).flatten()[0]]
    plain, _, _ = run_planned_batch(lib, emu, probs, usetiny, stage=False, **KW)
    staged, _, _ = run_planned_batch(lib, emu, probs, usetiny, stage=True, copyback=True, **KW)
    for i in range(len(probs)):
        for key in ("X", "S", "x", "s", "y"):
            assert np.array_equal(plain[i][key], staged[i][key]), key
        assert np.abs(plain[i]["X"]).max() > 0 and plain[i]["dobj"] == staged[i]["dobj"]


@pytest.mark.parametrize("extra", [dict(setting=2), dict(setting=3), dict(lambdastar=50.0), dict(absgaptol=1e-7), dict(objlimit=0.1), dict(maxiter=5)],
                         ids=lambda d: "-".join(f"{k}={v}" for k, v in d.items()))
def test_parameters_reach_the_packed_kernel(lib, emu, extra):
    """the solver parameters as the batch descriptor carries them (settings of SCIP_SDPSOLVERSETTING, prescribed lambdastar, absolute
    gap, objective limit, iteration limit): same status, stop reason, iteration count and objective as the oracle"""
    fp, _ = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds().flatten()
    kw = dict(KW, **extra)
    r = run_batch(lib, emu, [fp], tiny=True, **kw)[0]
    ref = abi.Solver(abi.Lib(abi.ORACLE_LIB)).solve(fp, **kw)
    assert (r["phase_name"], r["stop_name"], r["iterations"]) == (ref["phase_name"], ref["stop_name"], ref["iterations"])
    assert abs(r["dobj"] - ref["dobj"]) <= 1e-7 * max(1.0, abs(ref["dobj"]))


def test_packed_factor_shares_the_tile_of_the_small_schur_complements(lib, emu):
    """frontier launch of the 256-thread kernel with Schur complements above 64 (example_MkP, m = 105) next to small ones (example_TT):
    the plan puts the packed factor where the 64 x 65 tile of the m <= 64 variant lives and adds only the excess to the launch
    (three CTAs per SM instead of two on a B200: 64 KB instead of 103 KB per CTA); the emulator guards the shared memory, the bounds
    match the oracle"""
    K = misdp.read_sdpa(os.path.join(GOLDEN, "example_MkP.dat-s.gz")).rows_to_bounds().flatten()[0]
    T = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds().flatten()[0]
    assert K.m == 105
    probs = [T, K, T]
    par = lib.default_params(**KW)
    F = lib.lib.sdpcuda_debug_pack_batch
    structs = [p.struct() for p in probs]
    ps = (C.POINTER(abi.Problem) * 3)(*[C.pointer(s) for s in structs])
    nimg, nwork, ny, nb, nt = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_int(0), C.c_int(0)
    stagebytes = (C.c_size_t * 2)()
    got, nbatched, ntiny = run_planned_batch(lib, emu, probs, True, **KW)        # sets the argument types of F
    assert nbatched == 3 and ntiny == 3
    assert F(3, ps, C.byref(par), 1, 0, 0, 0, 0, None, 0, C.byref(nimg), C.byref(nwork), C.byref(ny), None, 0, C.byref(nb), C.byref(nt), None, None,
             stagebytes) == 0
    packed = 8 * (105 * 106 // 2 + 2)
    tile = 8 * 64 * 65
    assert stagebytes[1] == 0 and packed - tile - 64 <= stagebytes[0] <= packed - tile + 256, (stagebytes[0], packed - tile)      # (64 bytes of slack end the kernel's own buffers)
    cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    for i, r in got.items():
        _compare(probs[i], r, cpu.solve(probs[i], **KW))
    assert got[0]["dobj"] == got[2]["dobj"] and np.array_equal(got[0]["y"], got[2]["y"])
