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
    assert batch(s.h, -1, None, C.byref(par), None, None) == 1          # SDPCUDA_ERR_ARG
    assert batch(s.h, 0, None, C.byref(par), None, None) == 0
    assert batch(s.h, 2, None, C.byref(par), None, None) == 1
    assert batch(None, 0, None, C.byref(par), None, None) == 1
    assert batch(s.h, 0, None, None, None, None) == 1
