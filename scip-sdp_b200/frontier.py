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
    return {i: model.flatten_fast(*node_bounds[i]) for i in partition(len(node_bounds), world, rank)}


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


def _penalty_ladder(solver, model, lb, ub, r, kw, penaltyparam=1e5, maxpenaltyparam=1e10, npenaltyincr=8, peninfeasadjust=10.0):
    """what SCIPsdpiSolve does after an unacceptable solve (sdpi.c:3437-3619): (ii) min r over the constraints + r I, r free, no
    objective: optimal with r > peninfeasadjust * max(feastol, gaptol), or infeasible => the node is infeasible; otherwise (iii) the
    penalty formulation with objective, r >= 0 and growing Gamma until r < feastol, whose y and bound are then those of the node.
    Returns a result dict like a solve ("pFEAS_dINF" for a node found infeasible, the original r if nothing worked)."""
    fp, info = model.flatten(lb, ub, compress=True, skip_single_rows=True, penalty=(1.0, False, False))
    q = solver.solve(fp, fetch=False, **kw)
    tol = max(kw.get("feastol", 1e-6), kw.get("gaptol", 1e-6))
    if (q["phase_name"] == "pdOPT" and q["dobj"] > peninfeasadjust * tol) or q["phase_name"] in ("pFEAS_dINF", "dINF"):
        return dict(r, phase_name="pFEAS_dINF")
    gamma, fact = penaltyparam, (maxpenaltyparam / penaltyparam) ** (1.0 / npenaltyincr)
    for _ in range(npenaltyincr + 1):
        fp, info = model.flatten(lb, ub, compress=True, skip_single_rows=True, penalty=(gamma, True, True))
        q = solver.solve(fp, fetch=False, **kw)
        if q["phase_name"] == "pdOPT":
            y = solver.get_y()
            if y[-1] < kw.get("feastol", 1e-6):          # feasorig: the solution is feasible for the node itself
                return dict(q, phase_name="pdOPT", dobj=float(q["dobj"] - gamma * y[-1]), y=y[:-1])
        elif q["phase_name"] in ("pFEAS_dINF", "dINF"):
            return dict(r, phase_name="pFEAS_dINF")
        gamma *= fact
    return r


def branch_and_bound(solver, model, mode="batch", width=1184, pool=None, gaptol=1e-5, feastol=1e-5, inttol=1e-5, maxnodes=1000000,
                     timelimit=600.0, verbose=False, dist=None, use_objlimit=False, native=False):
    """Frontier-synchronous branch-and-bound on the C ABI: in every round the (at most `width`) best open nodes are prepared like
    sdpi.c prepares a node (Misdp.node_problem), their relaxations are solved TOGETHER — mode "batch": one kernel launch, one CTA
    per node (sdpcuda_solve_batch); "threads": one host thread + stream per handle; "serial" — and the results are dispatched like
    relax_sdp.c does (relax_sdp.c:4180-4346): infeasibility certificate => cutoff, bound >= incumbent => cutoff, integral => new
    incumbent, otherwise most-infeasible branching (branch_sdpmostinf.c).  A relaxation that does not end optimal or with a
    certificate is solved again alone with the stable settings (the first rung of sdpi.c's ladder); if that fails too the node is
    branched on its first free integer variable with its parent's bound (nothing is lost, the count is reported as `unsolved`).
    use_objlimit (relaxing/SDP/objlimit of the reference, off by default there too): once an incumbent exists every relaxation is
    stopped as soon as its lower bound (the X-side objective of a feasible X-iterate) exceeds the cutoff — phase pUNBD => cutoff.
    native (mode "batch", one rank): the nodes of a round go to the library as bound vectors only (sdpcuda_solve_nodes); node
    presolve and marshalling run in C++ (csrc/node_marshal.hpp, the same arrays as Misdp.node_problem), Python keeps the tree.
    dist: torch.distributed with world size N > 1 (one process per GPU): every rank runs the same deterministic tree, in each round
    rank r solves the nodes r, r + N, ... of the round on its own device and the results (status, bound, y) are all-gathered — the
    partition of SURVEY.md 8e.1, no collective on the data path of a relaxation.
    -> dict(status, objval, sol, nodes, rounds, unsolved, seconds)"""
    import heapq
    import itertools
    import math
    import time
    t0 = time.time()
    ints = np.flatnonzero(model.integer)
    indicators = list(getattr(model, "indicators", []))
    lb0, ub0 = model.lb.copy(), model.ub.copy()
    lb0[ints] = np.ceil(lb0[ints] - inttol)
    ub0[ints] = np.floor(ub0[ints] + inttol)
    best, bestsol = math.inf, None
    tick = itertools.count()
    heap = [(-math.inf, next(tick), lb0, ub0)]
    nodes = rounds = unsolved = iterations = 0
    kw = dict(gaptol=gaptol, feastol=feastol)
    handles = [solver] + list(pool or [])
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0

    def cutoff(bound):
        return bound >= best - 1e-6 * max(1.0, abs(best))

    def push_children(bound, lb, ub, j, value):
        dn_ub = ub.copy(); dn_ub[j] = math.floor(value)
        up_lb = lb.copy(); up_lb[j] = math.ceil(value)
        if math.floor(value) == math.ceil(value):            # integral value (unsolved node / indicator): split below and above it
            up_lb[j] = value + 1.0
        if dn_ub[j] >= lb[j] - inttol:
            heapq.heappush(heap, (bound, next(tick), lb, dn_ub))
        if up_lb[j] <= ub[j] + inttol:
            heapq.heappush(heap, (bound, next(tick), up_lb, ub))

    expired = False
    use_native = native and mode == "batch" and world == 1
    nmodel = None
    if use_native:
        from . import abi
        nmodel = abi.Model(solver.L, model)
    while use_native and heap and nodes < maxnodes and not expired:
        rounds += 1
        expired = time.time() - t0 >= timelimit
        lbs, ubs, bounds = [], [], []
        while heap and len(lbs) < width:
            bound, _, lb, ub = heapq.heappop(heap)
            if cutoff(bound):
                continue
            nodes += 1
            for sl, z in indicators:
                if lb[z] > 0.5 and ub[sl] > 0.0:
                    ub = ub.copy(); ub[sl] = 0.0
            lbs.append(lb); ubs.append(ub); bounds.append(bound)
        if not lbs:
            continue
        cut = None
        if use_objlimit and best < math.inf:
            cut = np.full(len(lbs), best - 1e-6 * max(1.0, abs(best)))
        out = solver.solve_nodes(nmodel, np.array(lbs), np.array(ubs), cutoff=cut, **kw)
        for k in range(len(lbs)):
            st, r = int(out["status"][k]), out["results"][k]
            lb, ub, y = out["lb"][k], out["ub"][k], out["y"][k]
            if st == 1:
                continue
            if st == 2:
                if out["bound"][k] < best and not cutoff(out["bound"][k]):
                    best, bestsol = float(out["bound"][k]), y.copy()
                continue
            iterations += int(r["iterations"])
            obj = float(out["bound"][k])
            if r["phase_name"] not in ("pdOPT", "pFEAS_dINF", "dINF", "pINF_dFEAS", "pUNBD"):
                # unacceptable solve: the few such nodes go through the Python path (stable settings, then the penalty ladder)
                status, fp, info = model.node_problem_fast(lbs[k], ubs[k], feastol=feastol)
                r = solver.solve(fp, fetch=False, setting=3, **kw)
                r["y"] = solver.get_y()
                if r["phase_name"] not in ("pdOPT", "pFEAS_dINF", "dINF", "pINF_dFEAS"):
                    r = _penalty_ladder(solver, model, info["lb"], info["ub"], r, kw)
                if r["phase_name"] == "pdOPT":
                    y = info["lb"].copy(); y[info["active"]] = r["y"]
                    obj = r["dobj"] + info["fixedobj"]
            if r["phase_name"] in ("pFEAS_dINF", "dINF", "pUNBD"):
                continue
            if r["phase_name"] == "pINF_dFEAS":
                return dict(status="unbounded", objval=-math.inf, sol=None, nodes=nodes, rounds=rounds, unsolved=unsolved, iterations=iterations,
                            seconds=time.time() - t0)
            free = [j for j in ints if ub[j] - lb[j] > 0.5]
            if r["phase_name"] != "pdOPT":
                unsolved += 1
                if free:
                    push_children(bounds[k], lb, ub, free[0], math.floor(0.5 * (lb[free[0]] + ub[free[0]])) + 0.5)
                continue
            if cutoff(obj):
                continue
            frac = np.abs(y[ints] - np.round(y[ints])) if len(ints) else np.zeros(0)
            if len(ints) and frac.max() > inttol:
                j = ints[int(np.argmax(frac))]
                push_children(obj, lb, ub, j, y[j])
                continue
            viol = [z for sl, z in indicators if y[z] > 0.5 and ub[z] - lb[z] > 0.5 and y[sl] > 1e-6]
            if viol:
                push_children(obj, lb, ub, viol[0], 0.5)
                continue
            best, bestsol = obj, y.copy()
    while not use_native and heap and nodes < maxnodes and not expired:
        rounds += 1
        expired = time.time() - t0 >= timelimit            # acted upon after this round (and, with several ranks, agreed upon)
        todo = []
        while heap and len(todo) < width:
            bound, _, lb, ub = heapq.heappop(heap)
            if cutoff(bound):
                continue
            nodes += 1
            for sl, z in indicators:                         # binary = 1 => slack = 0 (cons_indicator), through the slack's bound
                if lb[z] > 0.5 and ub[sl] > 0.0:
                    ub = ub.copy(); ub[sl] = 0.0
            status, fp, info = model.node_problem_fast(lb, ub, feastol=feastol)
            if status == "infeasible":
                continue
            if status == "allfixed":
                if info["fixedobj"] < best and not cutoff(info["fixedobj"]):
                    best, bestsol = info["fixedobj"], info["y"]
                continue
            todo.append((bound, fp, info))
        if not todo:
            continue
        alltodo = todo
        if world > 1:
            todo = alltodo[rank::world]
        if not todo:
            results = []
        elif mode == "batch":
            results = []
            for c in range(0, len(todo), width):
                part = todo[c:c + width]
                limits = None
                if use_objlimit and best < math.inf:
                    limits = [best - 1e-6 * max(1.0, abs(best)) - info["fixedobj"] for _, _, info in part]
                results += solver.solve_batch([fp for _, fp, _ in part], objlimits=limits, **kw)
        elif mode == "threads" and len(handles) > 1:
            import threading
            results = [None] * len(todo)

            def work(k):
                for i in range(k, len(todo), len(handles)):
                    results[i] = handles[k].solve(todo[i][1], fetch=False, **kw)
                    results[i]["y"] = handles[k].get_y()

            th = [threading.Thread(target=work, args=(k,)) for k in range(min(len(handles), len(todo)))]
            for t in th:
                t.start()
            for t in th:
                t.join()
        else:
            results = []
            for _, fp, _ in todo:
                r = solver.solve(fp, fetch=False, **kw)
                r["y"] = solver.get_y()
                results.append(r)
        for k, ((bound, fp, info), r) in enumerate(zip(todo, results)):       # repair unacceptable solves on the rank that owns the node
            if r["phase_name"] not in ("pdOPT", "pFEAS_dINF", "dINF", "pINF_dFEAS", "pUNBD"):
                r = solver.solve(fp, fetch=False, setting=3, **kw)
                r["y"] = solver.get_y()
            if r["phase_name"] not in ("pdOPT", "pFEAS_dINF", "dINF", "pINF_dFEAS", "pUNBD"):
                r = _penalty_ladder(solver, model, info["lb"], info["ub"], r, kw)
            results[k] = r
        if world > 1:
            slim = [dict(phase_name=r["phase_name"], dobj=float(r["dobj"]), iterations=int(r.get("iterations", 0)),
                         y=np.asarray(r["y"], dtype=float)) for r in results]
            gathered = [None] * world
            dist.all_gather_object(gathered, (expired, slim))
            results = [None] * len(alltodo)
            for q, (flag, part) in enumerate(gathered):
                results[q::world] = part
                expired = expired or flag
            todo = alltodo
        for (bound, fp, info), r in zip(todo, results):
            lb, ub = info["lb"], info["ub"]
            iterations += int(r.get("iterations", 0))
            if r["phase_name"] in ("pFEAS_dINF", "dINF", "pUNBD"):      # infeasible, or bound above the cutoff (objective limit)
                continue
            if r["phase_name"] == "pINF_dFEAS":
                return dict(status="unbounded", objval=-math.inf, sol=None, nodes=nodes, rounds=rounds, unsolved=unsolved, seconds=time.time() - t0)
            y = lb.copy()
            y[info["active"]] = r["y"]
            free = [j for j in ints if ub[j] - lb[j] > 0.5]
            if r["phase_name"] != "pdOPT":
                unsolved += 1
                if free:
                    push_children(bound, lb, ub, free[0], math.floor(0.5 * (lb[free[0]] + ub[free[0]])) + 0.5)
                continue
            obj = r["dobj"] + info["fixedobj"]
            if cutoff(obj):
                continue
            frac = np.abs(y[ints] - np.round(y[ints])) if len(ints) else np.zeros(0)
            if len(ints) and frac.max() > inttol:
                j = ints[int(np.argmax(frac))]
                push_children(obj, lb, ub, j, y[j])
                continue
            viol = [z for sl, z in indicators if y[z] > 0.5 and ub[z] - lb[z] > 0.5 and y[sl] > 1e-6]
            if viol:
                push_children(obj, lb, ub, viol[0], 0.5)
                continue
            best, bestsol = obj, y
            if verbose:
                print(f"round {rounds}: incumbent {best:.8g} ({nodes} nodes)")
    done = not heap or all(cutoff(h[0]) for h in heap)
    status = ("optimal" if bestsol is not None else "infeasible") if done else "limit"
    return dict(status=status, objval=best, sol=bestsol, nodes=nodes, rounds=rounds, unsolved=unsolved, iterations=iterations,
                seconds=time.time() - t0)


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
