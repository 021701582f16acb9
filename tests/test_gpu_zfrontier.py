"""GPU suite, last part: several frontier nodes on ONE device at a time — sdpcuda_solve_batch (one kernel launch, one CTA per node)
and one host thread + stream per handle.  NOTE: written after the GPU budget of round 1 was spent; the host plumbing is covered on
the CPU oracle (tests/test_frontier_gloo.py), the device side of these tests has not run on a B200 yet (the file sorts last so
that it cannot mask the verified suites).  Order inside the file: first the cases that only use kernels which have already run on a
B200 (classic path, device-resident check, several handles on host threads), then the batch kernel, its 256-thread instantiation,
the staged work space and the packed single solve — a device fault in a later group cannot hide the earlier ones."""
import os

import numpy as np
import pytest

from scip_sdp_b200 import abi, frontier, generators, misdp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KW = dict(gaptol=1e-6, feastol=1e-6)


def _frontier(M, q):
    """all 0/1 fixings of the first q integer variables (2^q nodes)"""
    ints = np.flatnonzero(M.integer)[:q]
    out = []
    for code in range(1 << q):
        lb, ub = M.lb.copy(), M.ub.copy()
        for b, j in enumerate(ints):
            lb[j] = ub[j] = float((code >> b) & 1)
        out.append((lb, ub))
    return out


@pytest.fixture(scope="module")
def lib():
    return abi.Lib(abi.PRODUCT_LIB)


@pytest.fixture(scope="module")
def cpu():
    return abi.Solver(abi.Lib(abi.ORACLE_LIB))


@pytest.mark.parametrize("name,want", [("example_inf.dat-s", None), ("example_small_ind.dat-s", -18.0)])
def test_bnb_through_the_reference_sdpi_layer_later_cases(name, want):
    """the two short.test instances whose harness-level reading changed after the last GPU run of round 1 (example_inf: block-size
    line with a glued comment, now read like reader_sdpa.c does; example_small_ind: indicator entries), through the reference's
    sdpi.c + sdpisolver_cuda.c + libsdpcuda exactly like tests/test_gpu_sdpi.py"""
    from harness import bnb, sdpi_ref
    os.environ.setdefault("SHIM_QUIET", "1")
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    r = bnb.solve_misdp(sdpi_ref.SdpiLib(sdpi_ref.LIB_CUDA), M, timelimit=900)
    if want is None:
        assert r["status"] == "infeasible"
    else:
        assert r["status"] == "optimal" and r["unsolved"] == 0
        assert abs(M.file_objective(r["objval"]) - want) <= 1e-4 * max(1.0, abs(want))


def test_resident_psd_check_on_gpu(lib):
    """sdpcuda_check_psd_resident (device assemble + Cholesky of the resident problem) against numpy's eigenvalues"""
    from test_boundary_cpu import _resident_check_cases
    gpu = abi.Solver(lib, device=0)
    _resident_check_cases(gpu)
    gpu.close()


def test_checksdpi_known_answers_with_the_device_resident_post_check(monkeypatch):
    """the ported unittests/src/checksdpi.c cases with SDPCUDA_DEVICE_CHECK=1: the binding's post-check of every converged solve
    then runs on the device-resident problem instead of shipping a dense Z(y)"""
    from golden.checksdpi_cases import CASES
    from harness import checksdpi_port, sdpi_ref
    os.environ.setdefault("SHIM_QUIET", "1")
    monkeypatch.setenv("SDPCUDA_DEVICE_CHECK", "1")
    L = sdpi_ref.SdpiLib(sdpi_ref.LIB_CUDA)
    for name in sorted(CASES):
        if L.solver_name() in CASES[name].get("skip_for", []):
            continue
        checksdpi_port.run_case(L, CASES[name], name)


@pytest.mark.parametrize("mode", ["threads", "batch"])
def test_frontier_modes_on_one_gpu(lib, mode, monkeypatch):
    """frontier.solve_frontier on one device: same statuses and bounds (1e-7 relative) as the serial loop; "threads" runs a
    mid-size truss relaxation (multi-kernel path, CUDA graphs captured per thread) on 4 host threads"""
    if mode == "batch":
        M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    else:
        M = generators.truss(4, 4, 60, seed=13)
        monkeypatch.setenv("SDPCUDA_PATH", "m")
    nodes = _frontier(M, 3)
    pool = [abi.Solver(lib, device=0) for _ in range(4)]
    serial = frontier.solve_frontier(pool[0], M, nodes, **KW)
    got = frontier.solve_frontier(pool[0], M, nodes, pool=pool[1:], mode=mode, **KW)
    for a, b in zip(serial, got):
        assert a["status"] == b["status"]
        assert abs(a["bound"] - b["bound"]) <= 1e-7 * max(1.0, abs(a["bound"]))
    for s in pool:
        s.close()


@pytest.mark.parametrize("name,q", [("example_small.dat-s", 2), ("example_TT.dat-s.gz", 4), ("example_CLS.dat-s.gz", 3), ("example_MkP.dat-s.gz", 3)])
def test_batched_nodes_match_oracle_and_single_solves(lib, cpu, name, q, monkeypatch):
    """2^q nodes of a shipped instance in ONE launch: statuses and bounds as the oracle's (1e-5 relative, north_star tolerance),
    and the same numbers as the one-relaxation launch of the same kernel (the batch only differs in who writes the cold start)"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, name)).rows_to_bounds()
    flat = [M.flatten(lb, ub) for lb, ub in _frontier(M, q)]
    keep = [fp for fp, _ in flat if fp.m > 0]
    gpu = abi.Solver(lib, device=0)
    res = gpu.solve_batch(keep, **KW)
    assert sum(r["launches"] for r in res) == 1            # ONE kernel launch for the whole frontier
    monkeypatch.setenv("SDPCUDA_PATH", "s")
    one = abi.Solver(lib, device=0)
    for fp, r in zip(keep, res):
        ref = cpu.solve(fp, **KW)
        # optimal nodes must be optimal; for infeasible nodes any certificate phase counts (the two back ends may stop one iteration apart)
        assert (r["phase_name"] == "pdOPT") == (ref["phase_name"] == "pdOPT"), (r["phase_name"], r["stop_name"], ref["phase_name"])
        if ref["phase_name"] in ("pFEAS_dINF", "dINF"):
            assert r["phase_name"] in ("pFEAS_dINF", "dINF")
        if ref["phase_name"] == "pdOPT":
            assert abs(r["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
            assert np.allclose(r["y"], ref["y"], atol=1e-3 * max(1.0, np.abs(ref["y"]).max()))
        single = one.solve(fp, **KW)
        assert single["phase_name"] == r["phase_name"] and single["iterations"] == r["iterations"]
        assert abs(single["dobj"] - r["dobj"]) <= 1e-9 * max(1.0, abs(r["dobj"]))
        assert np.allclose(single["y"], r["y"], rtol=0, atol=1e-9 * max(1.0, np.abs(r["y"]).max()))
    gpu.close(); one.close()


def test_batch_with_a_node_outside_the_single_cta_limits(lib, cpu):
    """a block of order 96 does not fit the one-CTA kernel: that node is solved by the multi-kernel path inside the same call"""
    big, _ = generators.maxcut(96, 0.1, seed=7).flatten()
    small, _ = misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s")).rows_to_bounds().flatten()
    gpu = abi.Solver(lib, device=0)
    res = gpu.solve_batch([small, big, small], **KW)
    for fp, r in zip([small, big, small], res):
        ref = cpu.solve(fp, **KW)
        assert r["phase_name"] == ref["phase_name"] == "pdOPT"
        assert abs(r["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
        assert np.allclose(r["y"], ref["y"], atol=1e-3 * max(1.0, np.abs(ref["y"]).max()))
    assert res[1]["launches"] > 100 and res[0]["dobj"] == res[2]["dobj"]
    gpu.close()


def test_batch_larger_than_the_sm_count(lib, cpu):
    """300 nodes (two waves of CTAs on 148 SMs) of example_TT: every node's bound equals the one of its duplicate"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    flat = [M.flatten(lb, ub)[0] for lb, ub in _frontier(M, 3)]
    probs = [flat[i % 8] for i in range(300)]
    gpu = abi.Solver(lib, device=0)
    res = gpu.solve_batch(probs, **KW)
    for i, r in enumerate(res):
        assert r["phase_name"] == res[i % 8]["phase_name"] and r["dobj"] == res[i % 8]["dobj"] and np.array_equal(r["y"], res[i % 8]["y"])
    for i in range(8):
        ref = cpu.solve(flat[i], **KW)
        assert abs(res[i]["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
    gpu.close()


@pytest.mark.parametrize("chunks", ["2", "3", "5"])
def test_batch_in_chunks_equals_the_single_launch(lib, chunks, monkeypatch):
    """a frontier sent through in chunks (two buffer sets on two streams, chunk c + 1 packed while chunk c runs; the default for
    frontiers of several waves) returns bit for bit what the single launch returns, also with a node outside the single-CTA limits"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    flat = [M.flatten(lb, ub)[0] for lb, ub in _frontier(M, 3)]
    big, _ = generators.maxcut(96, 0.1, seed=7).flatten()
    probs = [flat[i % 8] for i in range(90)]
    probs[41] = big
    gpu = abi.Solver(lib, device=0)
    monkeypatch.setenv("SDPCUDA_BATCH_CHUNKS", "1")
    one = gpu.solve_batch(probs, **KW)
    monkeypatch.setenv("SDPCUDA_BATCH_CHUNKS", chunks)
    many = gpu.solve_batch(probs, **KW)
    again = gpu.solve_batch(probs, **KW)                  # both buffer sets reused
    for a, b, c in zip(one, many, again):
        assert a["phase_name"] == b["phase_name"] == c["phase_name"] == "pdOPT"
        assert a["dobj"] == b["dobj"] == c["dobj"] and a["iterations"] == b["iterations"] == c["iterations"]
        assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["y"], c["y"])
    gpu.close()


@pytest.mark.parametrize("lanes", ["2", "4"])
def test_midsize_nodes_on_several_lanes_equal_one_after_the_other(lib, lanes, monkeypatch):
    """nodes outside the single-CTA limits run side by side on helper handles (SDPCUDA_LONER_LANES, default 4): bit for bit the
    results of one node after the other on the caller's handle, small nodes of the same call untouched, handle reusable afterwards"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    flat = [M.flatten(lb, ub)[0] for lb, ub in _frontier(M, 3)]
    mids = [generators.maxcut(80 + 8 * k, 0.1, seed=11 + k).flatten()[0] for k in range(5)] + [generators.truss(4, 4, 150, seed=5).flatten()[0]]
    probs = [flat[i % 8] for i in range(12)]
    for k, q in enumerate(mids):
        probs.insert(2 * k + 1, q)
    gpu = abi.Solver(lib, device=0)
    monkeypatch.setenv("SDPCUDA_LONER_LANES", "1")
    one = gpu.solve_batch(probs, **KW)
    monkeypatch.setenv("SDPCUDA_LONER_LANES", lanes)
    many = gpu.solve_batch(probs, **KW)
    again = gpu.solve_batch(probs, **KW)
    assert sum(1 for r in one if r["launches"] > 100) >= len(mids) - 1      # the mid-size nodes took the multi-kernel path
    for a, b, c in zip(one, many, again):
        assert a["phase_name"] == b["phase_name"] == c["phase_name"] == "pdOPT"
        assert a["dobj"] == b["dobj"] == c["dobj"] and a["iterations"] == b["iterations"] == c["iterations"]
        assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["y"], c["y"])
    single = gpu.solve(mids[0], **KW)                                       # the caller's handle after the lanes
    assert single["dobj"] == one[1]["dobj"]
    gpu.close()


def test_objective_limits_per_node_on_gpu(lib):
    """per-node objective limits in the batch call: nodes with a limit below their value stop early with phase pUNBD"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    probs = [fp for fp, _ in (M.flatten(lb, ub) for lb, ub in _frontier(M, 3)) if fp.m > 0]
    gpu = abi.Solver(lib, device=0)
    free = gpu.solve_batch(probs, **KW)
    limits = [r["dobj"] - 0.05 if i % 2 == 0 else 1e20 for i, r in enumerate(free)]
    cut = gpu.solve_batch(probs, objlimits=limits, **KW)
    for i, (a, b) in enumerate(zip(free, cut)):
        if i % 2 == 0:
            assert b["phase_name"] == "pUNBD" and b["iterations"] < a["iterations"] and b["pobj"] > limits[i]
        else:
            assert b["phase_name"] == a["phase_name"] and b["dobj"] == a["dobj"]
    gpu.close()



@pytest.mark.parametrize("name,want", [("example_small.dat-s", -8.0), ("example_inf.dat-s", None), ("example_TT.dat-s.gz", 2.11803),
                                       ("example_CLS.dat-s.gz", 7.1485), ("example_MkP.dat-s.gz", -95.0), ("example_small_ind.dat-s", -18.0)])
def test_frontier_branch_and_bound_on_gpu(lib, name, want):
    """frontier-synchronous B&B with all open nodes of a round in one launch: the optimal values of check/testset/short.solu"""
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    gpu = abi.Solver(lib, device=0)
    r = frontier.branch_and_bound(gpu, M, mode="batch", width=592, timelimit=300)
    if want is None:
        assert r["status"] == "infeasible"
    else:
        assert r["status"] == "optimal", r
        assert abs(M.file_objective(r["objval"]) - want) <= 1e-4 * max(1.0, abs(want))
    gpu.close()


@pytest.mark.parametrize("name,want", [("example_TT.dat-s.gz", 2.11803), ("example_MkP.dat-s.gz", -95.0)])
def test_branch_and_bound_with_native_nodes_on_gpu(lib, name, want):
    """the rounds handed to the library as bound vectors (sdpcuda_solve_nodes: C++ presolve + marshalling + one launch per round)"""
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    gpu = abi.Solver(lib, device=0)
    r = frontier.branch_and_bound(gpu, M, mode="batch", width=592, native=True, use_objlimit=True, timelimit=300)
    assert r["status"] == "optimal" and abs(M.file_objective(r["objval"]) - want) <= 1e-4 * max(1.0, abs(want))
    gpu.close()


@pytest.mark.parametrize("name,q", [("example_small.dat-s", 2), ("example_TT.dat-s.gz", 4), ("example_MkP.dat-s.gz", 3)])
def test_tiny_instantiation_of_the_batch_kernel(lib, cpu, name, q, monkeypatch):
    """SDPCUDA_BATCH_TINY=1: relaxations with blocks of order <= 16 run in the 256-thread instantiation (four nodes per SM,
    csrc/ipm_tiny.cu): same statuses, bounds within 1e-7 relative of the 1024-thread kernel (reduction orders differ) and within
    the north-star tolerance of the oracle"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, name)).rows_to_bounds()
    keep = [fp for fp, _ in (M.flatten(lb, ub) for lb, ub in _frontier(M, q)) if fp.m > 0]
    gpu = abi.Solver(lib, device=0)
    regular = gpu.solve_batch(keep, **KW)
    monkeypatch.setenv("SDPCUDA_BATCH_TINY", "1")
    tiny = gpu.solve_batch(keep, **KW)
    for fp, a, b in zip(keep, regular, tiny):
        assert a["phase_name"] == b["phase_name"], (a["phase_name"], b["phase_name"], b["stop_name"])
        if a["phase_name"] == "pdOPT":
            assert abs(a["dobj"] - b["dobj"]) <= 1e-7 * max(1.0, abs(a["dobj"]))
            ref = cpu.solve(fp, **KW)
            assert abs(b["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
    gpu.close()


@pytest.mark.parametrize("tiny", ["0", "1"])
def test_work_space_staged_in_shared_memory_on_gpu(lib, tiny, monkeypatch):
    """SDPCUDA_BATCH_SMEM=1: the head of every node's work space lives in shared memory; the arithmetic is the same, so statuses,
    iteration counts, objectives and y are bit-identical to the run from global memory (the CPU emulation shows the same)"""
    T = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    probs = [fp for fp, _ in (T.flatten(lb, ub) for lb, ub in _frontier(T, 3)) if fp.m > 0]
    probs += [misdp.read_sdpa(os.path.join(GOLDEN, "example_MkP.dat-s.gz")).rows_to_bounds().flatten()[0],
              generators.cls(12, 9, 3, seed=5).flatten()[0], generators.maxcut(40, 0.2, seed=3).flatten()[0]]
    monkeypatch.setenv("SDPCUDA_BATCH_TINY", tiny)
    gpu = abi.Solver(lib, device=0)
    plain = gpu.solve_batch(probs, **KW)
    monkeypatch.setenv("SDPCUDA_BATCH_SMEM", "1")
    staged = gpu.solve_batch(probs, **KW)
    for a, b in zip(plain, staged):
        assert a["phase_name"] == b["phase_name"] and a["iterations"] == b["iterations"]
        assert a["dobj"] == b["dobj"] and np.array_equal(a["y"], b["y"])
    gpu.close()


@pytest.mark.parametrize("smem", ["0", "1"])
@pytest.mark.parametrize("name", ["example_small.dat-s", "example_TT.dat-s.gz", "example_CLS.dat-s.gz", "example_MkP.dat-s.gz"])
def test_packed_single_solve(lib, cpu, name, smem, monkeypatch):
    """SDPCUDA_PACKED_SOLVE=1: one relaxation through the packed path (one copy, one launch); objective, y and the multipliers the
    getters return afterwards (X, S, x, s) against the oracle and the complementarity / feasibility relations"""
    fp, _ = misdp.read_sdpa(os.path.join(GOLDEN, name)).rows_to_bounds().flatten()
    monkeypatch.setenv("SDPCUDA_PACKED_SOLVE", "1")
    monkeypatch.setenv("SDPCUDA_BATCH_SMEM", smem)
    monkeypatch.setenv("SDPCUDA_BATCH_TINY", smem)
    gpu = abi.Solver(lib, device=0)
    r = gpu.solve(fp, **KW)
    ref = cpu.solve(fp, **KW)
    assert r["phase_name"] == ref["phase_name"] == "pdOPT" and r["launches"] == 1
    assert abs(r["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
    Cd = fp.dense_C()
    for k in range(fp.nblocks):
        Z = sum(r["y"][j] * fp.dense_A(j)[k] for j in range(fp.m)) - Cd[k]
        assert np.allclose(r["S"][k], Z, atol=1e-5 * max(1.0, np.abs(Z).max()))                     # S = A'y - C
        assert np.linalg.eigvalsh(r["X"][k])[0] >= -1e-8 and abs(np.sum(r["X"][k] * r["S"][k])) <= 1e-4 * max(1.0, abs(ref["dobj"]))
    if fp.nlp:
        assert np.allclose(r["slp"], fp.dense_D() @ r["y"] - fp.lprhs, atol=1e-5 * max(1.0, np.abs(fp.lprhs).max()))
        assert r["xlp"].min() >= -1e-9
    again = gpu.solve_resident(**KW)
    assert again["dobj"] == r["dobj"] and again["iterations"] == r["iterations"]
    gpu.close()


def test_sdpi_layer_on_the_packed_single_solve(monkeypatch):
    """the reference's sdpi.c over the binding with packed single solves, staged work space and the device-resident post-check
    (which falls back to the shipped Z(y) for packed solves): ported checksdpi.c answers and two B&B optima"""
    from golden.checksdpi_cases import CASES
    from harness import bnb, checksdpi_port, sdpi_ref
    os.environ.setdefault("SHIM_QUIET", "1")
    for k in ("SDPCUDA_PACKED_SOLVE", "SDPCUDA_BATCH_SMEM", "SDPCUDA_BATCH_TINY", "SDPCUDA_DEVICE_CHECK"):
        monkeypatch.setenv(k, "1")
    L = sdpi_ref.SdpiLib(sdpi_ref.LIB_CUDA)
    for name in sorted(CASES):
        if L.solver_name() in CASES[name].get("skip_for", []):
            continue
        checksdpi_port.run_case(L, CASES[name], name)
    for name, want in (("example_small.dat-s", -8.0), ("example_TT.dat-s.gz", 2.11803)):
        M = misdp.read_instance(os.path.join(GOLDEN, name))
        r = bnb.solve_misdp(L, M, timelimit=600)
        assert r["status"] == "optimal" and abs(M.file_objective(r["objval"]) - want) <= 1e-4 * max(1.0, abs(want))
