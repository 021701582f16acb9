"""Independent node relaxations across GPUs: the open branch-and-bound frontier is partitioned
round-robin over the ranks of a torch.distributed job (one process per GPU, NCCL on the GPU box, gloo in CPU tests); every rank
solves its nodes on its own device handle and only the scalar results (status, bound) are exchanged — there is no data-path
collective.  `bench.py --gpus N` uses the same partition for its weak-scaling measurement."""
import numpy as np


def partition(nnodes, world, rank):
    """indices of the frontier nodes rank `rank` solves"""
    return list(range(rank, nnodes, world))


def flatten_nodes(model, node_bounds, world=1, rank=0):
    """solver-form problems of this rank's nodes ({node index: (FlatProblem, info)}): the host-side marshalling that sdpi.c does
    in C for SCIP-SDP; bench.py does it before the timed region"""
    return {i: model.flatten(*node_bounds[i]) for i in partition(len(node_bounds), world, rank)}


def _solve_nodes_batched(solver, todo, chunk, solve_kw):
    """chunks of `chunk` nodes through sdpcuda_solve_batch: one host->device copy, one kernel launch (one CTA per node) and one
    device->host copy per chunk for the relaxations that fit the single-CTA kernel; others are solved one by one inside the call"""
    out = {}
    for c in range(0, len(todo), chunk):
        part = todo[c:c + chunk]
        res = solver.solve_batch([fp for _, fp, _ in part], fetch=False, **solve_kw)
        for (i, _, info), r in zip(part, res):
            out[i] = dict(status=r["phase_name"], bound=float(r["dobj"] + info["fixedobj"]), iterations=int(r["iterations"]))
    return out


def _solve_nodes_threaded(pool, todo, solve_kw):
    """one host thread per handle (what SCIP's concurrent solver threads do through SCIPsdpiSolverCreate): the handles own
    their streams, so the latency-bound kernel chains of different nodes overlap on the device; ctypes releases the GIL"""
    import threading
    out, errors = {}, []

    def work(k):
        try:
            for i, fp, info in todo[k::len(pool)]:
                r = pool[k].solve(fp, fetch=False, **solve_kw)
                out[i] = dict(status=r["phase_name"], bound=float(r["dobj"] + info["fixedobj"]), iterations=int(r["iterations"]))
        except Exception as e:                       # noqa: BLE001 - re-raised on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(min(len(pool), len(todo)))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out


def solve_frontier(solver, model, node_bounds, dist=None, flat=None, pool=None, mode="serial", chunk=1184, **solve_kw):
    """solver: scip_sdp_b200.abi.Solver bound to this rank's device; node_bounds: list of (lb, ub) arrays, identical on all ranks;
    flat: optional result of flatten_nodes for this rank.  mode "serial": the nodes one after the other on `solver`; "batch": chunks
    of `chunk` nodes (default 8 per SM) in one kernel launch each (sdpcuda_solve_batch); "threads": one host thread and stream per
    handle of [solver] + pool.  Returns a list with one dict(status, bound) per node (complete on every rank)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine, todo = {}, []
    for i in partition(len(node_bounds), world, rank):
        lb, ub = node_bounds[i]
        fp, info = flat[i] if flat is not None else model.flatten(lb, ub)
        if fp.m == 0:
            mine[i] = dict(status="allfixed", bound=float(info["fixedobj"]))
            continue
        todo.append((i, fp, info))
    if mode not in ("serial", "batch", "threads"):
        raise ValueError(f"unknown frontier mode {mode!r}")
    handles = [solver] + list(pool or [])
    if mode == "batch":
        mine.update(_solve_nodes_batched(solver, todo, max(1, int(chunk)), solve_kw))
    elif mode == "threads" and len(handles) > 1:
        mine.update(_solve_nodes_threaded(handles, todo, solve_kw))
    else:
        for i, fp, info in todo:
            r = solver.solve(fp, fetch=False, **solve_kw)
            mine[i] = dict(status=r["phase_name"], bound=float(r["dobj"] + info["fixedobj"]), iterations=int(r["iterations"]))
    if world == 1:
        return [mine[i] for i in range(len(node_bounds))]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(node_bounds))]


def max_over_ranks(value, dist=None, device="cpu"):
    """the timing rule of bench.py: a multi-GPU time is the maximum over ranks"""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------- one large SDP over several GPUs
def schur_shares(nstrips, world):
    """host-side mirror of the device partition of the Schur complement: strip t (8 columns of the entry path, or one chunk
    of dense variables) belongs to rank t % world (ops.cu: schur_light_kernel / ipm.cu: dense chunks)"""
    return [list(range(r, nstrips, world)) for r in range(world)]


def broadcast_bytes(payload, nbytes, dist, src=0, device="cpu"):
    """rank `src` passes `payload` (bytes of length nbytes) to every rank over the job's process group: used for the
    128-byte NCCL unique id that sdpcuda_dist_init needs"""
    import torch
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def shard_one_sdp(solver, dist, device="cpu"):
    """joins the rank's device handle to the NCCL clique of the job: afterwards all ranks call solver.solve with the SAME problem
    and every rank forms only its share of the Schur complement"""
    world, rank = dist.get_world_size(), dist.get_rank()
    ident = solver.dist_unique_id() if rank == 0 else None
    ident = broadcast_bytes(ident, 128, dist, src=0, device=device)
    solver.dist_init(world, rank, ident)
