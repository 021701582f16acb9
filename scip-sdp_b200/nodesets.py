"""The B&B node workloads of bench.py's default line (BASELINE.json configs 2-4): a deterministic frontier of open nodes per
instance — all 0/1 fixings of the first q integer variables, enumerated by a code; rank r of a multi-GPU run takes the codes
[r * per_gpu, (r + 1) * per_gpu) (weak scaling, distinct nodes on every GPU, no data-path collective).  The oracle bounds of
these nodes are committed in tests/golden/frontier_bounds.npz (make_frontier_bounds.py)."""
import os

import numpy as np

from . import generators, misdp

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
MAX_RANKS = 8

# name -> (model constructor, nodes per GPU, how the nodes are solved)
#   "nodes":  sdpcuda_solve_nodes — bound vectors in, node presolve + marshalling + ONE launch (one CTA per node) in the library
#   "serial": one relaxation after the other through sdpcuda_solve (multi-kernel path), marshalled by Misdp.node_problem_fast
WORKLOADS = {
    "example_TT": (lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds(), 592, "nodes"),
    "example_MkP": (lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_MkP.dat-s.gz")).rows_to_bounds(), 592, "nodes"),
    "example_CLS": (lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_CLS.dat-s.gz")).rows_to_bounds(), 148, "nodes"),
    "TT-500": (lambda: generators.truss(6, 6, 500, seed=1001), 4, "serial"),
    "CLS-syn": (lambda: generators.cls(199, 99, 10, seed=2002), 4, "serial"),
    # MkP-120: the root relaxation only (code -1 = no fixing), one per GPU: deeper nodes of this instance lose strict feasibility, end
    # without convergence in the oracle as well (dFEAS after ~30 iterations) and would go through sdpi.c's penalty ladder
    "MkP-120": (lambda: generators.mkp(120, seed=3003), 1, "serial"),
}


def node_bounds(M, codes, q=16):
    """(lb, ub) arrays [len(codes) x nvars]: the code seeds a random partial fixing of the first q integer variables — free with
    probability 1/2, fixed to 0 with 3/8, to 1 with 1/8 (a frontier of mixed depth; 1s are rarer because the instances' rows bound
    how many variables may be 1); a value outside the variable's bounds leaves it alone; one integer variable always stays free"""
    ints = np.flatnonzero(M.integer)
    ints = ints[:min(q, max(1, len(ints) - 1))]
    lbs = np.tile(M.lb, (len(codes), 1))
    ubs = np.tile(M.ub, (len(codes), 1))
    for k, code in enumerate(codes):
        if code < 0:                       # the root node
            continue
        u = np.random.default_rng(7919 * int(code) + 13).random(len(ints))
        for t, j in enumerate(ints):
            v = 0.0 if u[t] < 0.375 else (1.0 if u[t] >= 0.875 else None)
            if v is not None and M.lb[j] <= v <= M.ub[j]:
                lbs[k, j] = ubs[k, j] = v
    return lbs, ubs


def golden():
    """the committed oracle results: per workload `<name>_codes` (the codes whose nodes reach the SDP solver and are solved to
    optimality by the oracle, in increasing order: the frontier), `<name>_bound` (their lower bounds)"""
    return np.load(os.path.join(GOLDEN, "frontier_bounds.npz"))


def frontier_of_rank(name, rank, per_gpu=None, table=None):
    """-> (codes, oracle bounds) of the nodes rank `rank` solves: slice [rank * per_gpu, (rank + 1) * per_gpu) of the frontier"""
    per = per_gpu or WORKLOADS[name][1]
    table = table if table is not None else golden()
    codes, bound = table[name + "_codes"], table[name + "_bound"]
    lo = (rank * per) % max(1, len(codes) - per + 1)
    return codes[lo:lo + per], bound[lo:lo + per]
