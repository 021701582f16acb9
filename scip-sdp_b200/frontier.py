"""Independent node relaxations across GPUs: the open branch-and-bound frontier is partitioned
round-robin over the ranks of a torch.distributed job (one process per GPU, NCCL on the GPU box, gloo in CPU tests); every rank
solves its nodes on its own device handle and only the scalar results (status, bound) are exchanged — there is no data-path
collective.  `bench.py --gpus N` uses the same partition for its weak-scaling measurement."""
import numpy as np


def partition(nnodes, world, rank):
    """indices of the frontier nodes rank `rank` solves"""
    return list(range(rank, nnodes, world))


def solve_frontier(solver, model, node_bounds, dist=None, **solve_kw):
    """solver: scip_sdp_b200.abi.Solver bound to this rank's device; node_bounds: list of (lb, ub) arrays, identical on all ranks.
    Returns a list with one dict(status, bound) per node (complete on every rank)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = {}
    for i in partition(len(node_bounds), world, rank):
        lb, ub = node_bounds[i]
        fp, info = model.flatten(lb, ub)
        if fp.m == 0:
            mine[i] = dict(status="allfixed", bound=float(info["fixedobj"]))
            continue
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
