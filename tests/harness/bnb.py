"""Minimal best-first branch-and-bound over the reference's SDPI (stand-in for the SCIP tree that is not available
here).  Per node it does what relax_sdp.c does around calcRelax (relax_sdp.c:4010-4346):
push the local bounds with SCIPsdpiChgBounds, call SCIPsdpiSolve, then dispatch on the status predicates
(IsAcceptable -> IsDualInfeasible => cutoff | feasible => lower bound from SCIPsdpiGetDualSol)."""
import heapq
import itertools
import math
import time

import numpy as np

from .sdpi_ref import Sdpi


def solve_misdp(lib, M, gaptol=1e-5, feastol=1e-5, inttol=1e-5, maxnodes=100000, verbose=False, timelimit=600.0):
    """-> dict(status, objval, sol, nodes, sdpcalls, iterations, seconds)"""
    s = Sdpi(lib, gaptol=gaptol, sdpsolverfeastol=feastol, feastol=feastol)
    t0 = time.time()
    try:
        s.load_model(M)
        ints = np.flatnonzero(M.integer)
        idx = np.arange(M.nvars, dtype=np.int32)
        best, bestsol = math.inf, None
        counter = itertools.count()
        lb0, ub0 = M.lb.copy(), M.ub.copy()
        lb0[ints] = np.ceil(lb0[ints] - inttol)
        ub0[ints] = np.floor(ub0[ints] + inttol)
        heap = [(-math.inf, next(counter), lb0, ub0)]
        nodes = calls = iters = unsolved = 0
        while heap and nodes < maxnodes and time.time() - t0 < timelimit:
            bound, _, lb, ub = heapq.heappop(heap)
            if bound >= best - 1e-6 * max(1.0, abs(best)):
                continue
            nodes += 1
            # indicator constraints (binary = 1 => slack = 0, SCIP's cons_indicator): enforced through the slack's upper bound
            for sl, z in getattr(M, "indicators", []):
                if lb[z] > 0.5:
                    ub = ub.copy(); ub[sl] = 0.0
            s.chg_bounds(idx, lb, ub)
            s.solve()
            st = s.stats()
            calls += st["sdpcalls"]; iters += st["iterations"]
            if not s.flag("WasSolved") or not s.flag("IsAcceptable"):
                unsolved += 1
                if verbose:
                    print(f"node {nodes}: relaxation not solved (status unknown) - node dropped")
                continue
            if s.flag("IsDualInfeasible"):
                continue
            if s.flag("IsDualUnbounded"):
                return dict(status="unbounded", objval=-math.inf, nodes=nodes, sdpcalls=calls, iterations=iters,
                            seconds=time.time() - t0, sol=None, unsolved=unsolved)
            if not (s.flag("IsPrimalFeasible") and s.flag("IsDualFeasible")):
                unsolved += 1
                continue
            obj, y = s.dual_sol()
            if obj >= best - 1e-6 * max(1.0, abs(best)):
                continue
            frac = np.abs(y[ints] - np.round(y[ints]))
            violated = [z for sl, z in getattr(M, "indicators", []) if frac.max(initial=0.0) <= inttol and y[z] > 0.5 and ub[z] - lb[z] > 0.5
                        and y[sl] > 1e-6]
            if violated:
                # integral but an indicator with value 1 has a positive slack: branch on that binary variable
                j = violated[0]
                dn_ub = ub.copy(); dn_ub[j] = 0.0
                up_lb = lb.copy(); up_lb[j] = 1.0
                heapq.heappush(heap, (obj, next(counter), lb, dn_ub))
                heapq.heappush(heap, (obj, next(counter), up_lb, ub))
                continue
            if len(ints) == 0 or frac.max() <= inttol:
                best, bestsol = obj, y.copy()
                if verbose:
                    print(f"node {nodes}: new incumbent {best:.8g}")
                continue
            j = ints[int(np.argmax(frac))]          # most infeasible branching (branch_sdpmostinf.c)
            dn_ub = ub.copy(); dn_ub[j] = math.floor(y[j])
            up_lb = lb.copy(); up_lb[j] = math.ceil(y[j])
            heapq.heappush(heap, (obj, next(counter), lb, dn_ub))
            heapq.heappush(heap, (obj, next(counter), up_lb, ub))
        status = "optimal" if bestsol is not None and not heap else ("infeasible" if bestsol is None and not heap else "limit")
        if heap and all(h[0] >= best - 1e-6 * max(1.0, abs(best)) for h in heap):
            status = "optimal"
        return dict(status=status, objval=best, sol=bestsol, nodes=nodes, sdpcalls=calls, iterations=iters,
                    seconds=time.time() - t0, unsolved=unsolved)
    finally:
        s.close()
