"""CPU suite, part 3: the N > 1 path (independent node relaxations partitioned over ranks) with world_size 2 on gloo.
The relaxations themselves run on the CPU oracle here; on the GPU box the same code runs with one device handle per rank."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from scip_sdp_b200 import abi, frontier, misdp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _nodes(M):
    """root + the four grandchildren obtained by branching on the first two integer variables"""
    ints = np.flatnonzero(M.integer)[:2]
    out = [(M.lb.copy(), M.ub.copy())]
    for a in (0, 1):
        for b in (0, 1):
            lb, ub = M.lb.copy(), M.ub.copy()
            lb[ints[0]] = ub[ints[0]] = a
            lb[ints[1]] = ub[ints[1]] = b
            out.append((lb, ub))
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
        solver = abi.Solver(abi.Lib(abi.ORACLE_LIB))
        res = frontier.solve_frontier(solver, M, _nodes(M), dist=dist, gaptol=1e-6, feastol=1e-6)
        tmax = frontier.max_over_ranks(1.0 + rank, dist=dist)
        # the channel that carries the NCCL id of the sharded-Schur path (rank 0 -> all)
        ident = frontier.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, dist)
        assert ident == bytes(range(128))

        class FakeHandle:          # stands in for abi.Solver: records what shard_one_sdp hands to sdpcuda_dist_init
            def dist_unique_id(self):
                return bytes([7 + rank]) * 128          # only rank 0's id may be used

            def dist_init(self, nranks, r, idbytes):
                self.args = (nranks, r, idbytes)

        fake = FakeHandle()
        frontier.shard_one_sdp(fake, dist)
        assert fake.args == (world, rank, bytes([7]) * 128)
        q.put((rank, [(r["status"], round(r["bound"], 6)) for r in res], tmax, frontier.partition(5, world, rank)))
    finally:
        dist.destroy_process_group()


def test_partition_is_a_disjoint_cover():
    for world in (1, 2, 4, 8):
        allidx = sorted(i for r in range(world) for i in frontier.partition(13, world, r))
        assert allidx == list(range(13))


def test_schur_shares_are_a_disjoint_cover():
    for world in (1, 2, 3, 8):
        shares = frontier.schur_shares(251, world)
        assert sorted(t for sh in shares for t in sh) == list(range(251))
        assert max(len(sh) for sh in shares) - min(len(sh) for sh in shares) <= 1


def test_oracle_has_no_sharded_path():
    s = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    s.dist_init(1, 0, None)                      # a clique of one is fine
    with pytest.raises(RuntimeError):
        s.dist_init(2, 0, bytes(128))


def test_frontier_world_size_2_matches_single_process():
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    single = frontier.solve_frontier(abi.Solver(abi.Lib(abi.ORACLE_LIB)), M, _nodes(M), gaptol=1e-6, feastol=1e-6)
    expect = [(r["status"], round(r["bound"], 6)) for r in single]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, tmax, part in got:
        assert res == expect                   # every rank ends up with the complete, identical result list
        assert tmax == 2.0                     # max over ranks of (1 + rank)
        assert part == list(range(rank, 5, 2))
    # child bounds can only be worse (larger) than the root relaxation bound
    assert all(b >= expect[0][1] - 1e-6 for st, b in expect[1:] if st == "pdOPT")


@pytest.mark.parametrize("mode", ["batch", "threads"])
def test_frontier_modes_agree_with_serial(mode):
    """the two ways of running several frontier nodes on ONE device at a time — sdpcuda_solve_batch (one launch, one CTA per node
    on the GPU; the oracle's stand-in loops) and one host thread per handle — return exactly what the serial loop returns"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    lib = abi.Lib(abi.ORACLE_LIB)
    kw = dict(gaptol=1e-6, feastol=1e-6)
    serial = frontier.solve_frontier(abi.Solver(lib), M, _nodes(M), **kw)
    pool = [abi.Solver(lib) for _ in range(3)]             # fewer handles than nodes: strided work lists
    got = frontier.solve_frontier(pool[0], M, _nodes(M), pool=pool[1:], mode=mode, chunk=2, **kw)      # chunks of 2: 2 + 2 + 1 nodes
    assert [(r["status"], r["bound"]) for r in got] == [(r["status"], r["bound"]) for r in serial]
    with pytest.raises(ValueError):
        frontier.solve_frontier(pool[0], M, _nodes(M), mode="nonsense", **kw)


def test_solve_batch_boundary():
    """argument checks of sdpcuda_solve_batch (include/sdpcuda.h) and its outputs"""
    import ctypes as C
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s"))
    lib = abi.Lib(abi.ORACLE_LIB)
    nodes = [M.flatten(lb, ub)[0] for lb, ub in _nodes(M)]
    s, ref = abi.Solver(lib), abi.Solver(lib)
    res = s.solve_batch(nodes, gaptol=1e-6, feastol=1e-6)
    for fp, r in zip(nodes, res):
        q = ref.solve(fp, gaptol=1e-6, feastol=1e-6)
        assert r["phase_name"] == q["phase_name"] and r["dobj"] == q["dobj"] and np.array_equal(r["y"], q["y"])
    assert s.solve_batch([]) == []
    assert "y" not in s.solve_batch(nodes[:2], fetch=False)[0]
    par = lib.default_params()
    batch = lib.lib.sdpcuda_solve_batch
    assert batch(s.h, -1, None, C.byref(par), None, None, None) == 1          # SDPCUDA_ERR_ARG
    assert batch(s.h, 0, None, C.byref(par), None, None, None) == 0
    assert batch(s.h, 2, None, C.byref(par), None, None, None) == 1
    assert batch(None, 0, None, C.byref(par), None, None, None) == 1
    assert batch(s.h, 0, None, None, None, None, None) == 1


# check/testset/short.solu (reference): B&B optima of the short.test instances the harness can read (the same 13 that
# tests/test_oracle_golden.py checks through the reference's sdpi.c; None = infeasible)
SHORT_SOLU = {"example_small.dat-s": -8.0, "example_inf.dat-s": None, "example_TT.dat-s.gz": 2.11803,
              "example_CLS.dat-s.gz": 7.1485, "example_MkP.dat-s.gz": -95.0,
              "example_small_cbf.cbf": -8.0, "example_cbf_primal.cbf": 0.75, "example_cbf_mix.cbf": 4.0, "example_cbf_dual.cbf": 4.0,
              "example_multaggr.cbf": -1.0, "example_diagzeroimpl.cbf": -1.0, "example_tightenmatrices.dat-s": -9.0,
              "example_small_ind.dat-s": -18.0}


@pytest.mark.parametrize("name", sorted(SHORT_SOLU))
def test_frontier_branch_and_bound_reaches_short_solu(name):
    """frontier-synchronous B&B straight on the C ABI (node presolve of sdpi.c restated in Misdp.node_problem, all open nodes of a
    round solved by one sdpcuda_solve_batch call, penalty ladder for unacceptable solves): the reference's optimal values"""
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    r = frontier.branch_and_bound(abi.Solver(abi.Lib(abi.ORACLE_LIB)), M, mode="batch", width=64, timelimit=300)
    if SHORT_SOLU[name] is None:
        assert r["status"] == "infeasible"
    else:
        assert r["status"] == "optimal" and r["unsolved"] == 0
        assert abs(M.file_objective(r["objval"]) - SHORT_SOLU[name]) <= 1e-4 * max(1.0, abs(SHORT_SOLU[name]))
        # the incumbent is integral and feasible for the original model
        y = r["sol"]
        assert np.abs(y[M.integer] - np.round(y[M.integer])).max(initial=0.0) <= 1e-5
        assert all(np.linalg.eigvalsh(Z)[0] >= -1e-5 for Z in M.dense_Z(y))


@pytest.mark.parametrize("mode", ["serial", "threads"])
def test_frontier_branch_and_bound_other_modes(mode):
    lib = abi.Lib(abi.ORACLE_LIB)
    M = misdp.read_instance(os.path.join(GOLDEN, "example_MkP.dat-s.gz"))
    r = frontier.branch_and_bound(abi.Solver(lib), M, mode=mode, width=1 if mode == "serial" else 8, pool=[abi.Solver(lib) for _ in range(3)])
    assert r["status"] == "optimal" and abs(r["objval"] + 95.0) <= 1e-2


def test_node_problem_follows_sdpi_presolve():
    """Misdp.node_problem: rows without active variables decide or vanish, one-variable rows become bounds, empty block rows/columns
    are removed (sdpi.c:691-810,1131-1290,3219-3290)"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz"))
    ints = np.flatnonzero(M.integer)
    lb, ub = M.lb.copy(), M.ub.copy()
    ub[ints[:30]] = 0.0
    st, fp, info = M.node_problem(lb, ub)
    full, _ = M.flatten(lb, ub)
    assert st == "solve" and fp.m <= full.m and fp.blocksizes[0] < full.blocksizes[0] and fp.nlp < full.nlp
    lb2 = M.lb.copy(); lb2[ints] = 1.0                       # every bar at every size: the one-size-per-bar rows are violated
    assert M.node_problem(lb2, M.ub.copy())[0] == "infeasible"
    lb3, ub3 = M.lb.copy(), M.ub.copy()
    lb3[0] = 1.0; ub3[0] = 0.0
    assert M.node_problem(lb3, ub3)[0] == "infeasible"       # crossed bounds
    S = misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s"))
    y = np.array([1.0, 1.0, 1.0])
    st, fp, info = S.node_problem(y, y)
    assert st in ("allfixed", "infeasible") and fp is None


def test_penalty_form_of_flatten():
    """the penalty formulation (sdpisolver.h:258-322): r with objective Gamma on every block diagonal and LP row, not on bounds"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s"))
    fp0, _ = M.flatten()
    fp, info = M.flatten(penalty=(7.0, True, True))
    assert fp.m == fp0.m + 1 and fp.obj[-1] == 7.0 and np.array_equal(fp.obj[:-1], fp0.obj)
    r_ent = range(fp.varbeg[fp.m - 1], fp.varbeg[fp.m])
    assert [(fp.entblk[e], fp.entrow[e], fp.entcol[e], fp.entval[e]) for e in r_ent] == [(b, i, i, 1.0) for b, n in enumerate(fp.blocksizes) for i in range(n)]
    nrows = len(info["rowmap"])
    D = fp.dense_D()
    assert np.all(D[:nrows, -1] == 1.0) and np.all(D[nrows:-1, -1] == 0.0) and D[-1, -1] == 1.0 and fp.lprhs[-1] == 0.0
    assert fp.nlp == fp0.nlp + 1
    fpf, _ = M.flatten(penalty=(1.0, False, False))
    assert fpf.nlp == fp0.nlp and np.all(fpf.obj[:-1] == 0.0) and fpf.obj[-1] == 1.0
    # feasible problem: min r is (clearly) negative, i.e. a strictly feasible point exists
    q = abi.Solver(abi.Lib(abi.ORACLE_LIB)).solve(fpf, gaptol=1e-6, feastol=1e-6)
    assert q["phase_name"] == "pdOPT" and q["dobj"] < -1e-3


def _bnb_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M = misdp.read_instance(os.path.join(GOLDEN, "example_MkP.dat-s.gz"))
        r = frontier.branch_and_bound(abi.Solver(abi.Lib(abi.ORACLE_LIB)), M, mode="batch", width=64, dist=dist)
        q.put((rank, r["status"], r["objval"], r["nodes"], r["rounds"]))
    finally:
        dist.destroy_process_group()


def test_branch_and_bound_world_size_2_runs_the_same_tree():
    """the rounds of the B&B driver partitioned over two ranks (gloo): both ranks end with the tree of the single-process run"""
    M = misdp.read_instance(os.path.join(GOLDEN, "example_MkP.dat-s.gz"))
    single = frontier.branch_and_bound(abi.Solver(abi.Lib(abi.ORACLE_LIB)), M, mode="batch", width=64)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bnb_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, status, objval, nodes, rounds in got:
        assert (status, nodes, rounds) == (single["status"], single["nodes"], single["rounds"])
        assert objval == single["objval"]


def test_objective_limits_per_node():
    """sdpcuda_solve_batch with per-node objective limits (relaxing/SDP/objlimit): a relaxation whose lower bound passes its limit
    stops with phase pUNBD; the B&B driver with use_objlimit reaches the same optimum over the same tree with fewer iterations"""
    lib = abi.Lib(abi.ORACLE_LIB)
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    nodes = [M.flatten(lb, ub)[0] for lb, ub in _nodes(M)]
    s = abi.Solver(lib)
    free = s.solve_batch(nodes, gaptol=1e-6, feastol=1e-6)
    limits = [r["dobj"] - 0.05 if i % 2 == 0 else 1e20 for i, r in enumerate(free)]
    cut = s.solve_batch(nodes, objlimits=limits, gaptol=1e-6, feastol=1e-6)
    for i, (a, b) in enumerate(zip(free, cut)):
        if i % 2 == 0:
            assert b["phase_name"] == "pUNBD" and b["stop_name"] == "objlimit" and b["iterations"] < a["iterations"] and b["pobj"] > limits[i]
        else:
            assert b["phase_name"] == a["phase_name"] and b["dobj"] == a["dobj"]
    R = misdp.read_instance(os.path.join(GOLDEN, "example_TT.dat-s.gz"))
    plain = frontier.branch_and_bound(abi.Solver(lib), R, mode="batch", width=64)
    fast = frontier.branch_and_bound(abi.Solver(lib), R, mode="batch", width=64, use_objlimit=True)
    assert (plain["status"], plain["nodes"]) == (fast["status"], fast["nodes"]) and abs(plain["objval"] - fast["objval"]) <= 1e-6
    assert fast["iterations"] < plain["iterations"]
