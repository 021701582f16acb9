"""Checks of the solver boundary itself (SCIPsdpiSolver*, sdpisolver.h), shared by the CPU (oracle back end) and GPU test files:
the penalty formulation call patterns of sdpi.c (patterns ii-iv) and the primal-matrix getters."""
import os

import numpy as np

from scip_sdp_b200 import misdp, sdpisolver_host

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")


def run_penalty_patterns(libpath):
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    bp = sdpisolver_host.BoundaryProblem(M)
    s = sdpisolver_host.SdpiSolver(libpath, gaptol=1e-6, feastol=1e-6)
    try:
        s.load_and_solve(bp)
        assert s.flag("IsAcceptable") and s.flag("IsOptimal")
        obj0, y0 = s.dual_sol()
        # (iii) penalty formulation with objective and r >= 0: exact for a large Gamma -> r = 0, same optimum, feasorig
        feasorig, penbound = s.load_and_solve_with_penalty(bp, 1e5, True, True)
        assert s.flag("IsAcceptable") and feasorig and not penbound
        obj1, y1 = s.dual_sol()
        assert abs(obj1 - obj0) <= 1e-5 * max(1.0, abs(obj0))
        assert np.all(np.isfinite(y1)) and len(y1) == len(y0)       # the optimal face of this relaxation is not a single point
        # (ii) feasibility phase: min r, r free, no objective: strictly feasible problem -> optimal r < 0 (Slater holds)
        feasorig, _ = s.load_and_solve_with_penalty(bp, 1.0, False, False)
        assert s.flag("WasSolved")
        if s.flag("IsOptimal"):
            assert s.objval() < 1e-6 and feasorig
        else:
            assert s.flag("IsDualUnbounded") or not s.flag("IsAcceptable")   # r can go to -infinity when y is unbounded below
        # (ii) on an infeasible problem: [[y1, 1], [1, 0.75 y2]] psd, |y| <= 1  ->  r* > 0
        I = misdp.Misdp(2, [-1.0, 0.0], [2])
        I.A[0][0] = [(0, 0, 1.0)]; I.A[0][1] = [(1, 1, 0.75)]; I.C[0] = [(1, 0, -1.0)]
        I.lb[:] = -1.0; I.ub[:] = 1.0
        bpi = sdpisolver_host.BoundaryProblem(I)
        s.load_and_solve(bpi)
        assert s.flag("IsDualInfeasible") or not s.flag("IsAcceptable")
        feasorig, _ = s.load_and_solve_with_penalty(bpi, 1.0, False, False)
        assert s.flag("IsOptimal") and not feasorig
        assert s.objval() > 1e-3            # decision rule of sdpi.c:3484: objective > tolerance => node infeasible
    finally:
        s.close()


def run_primal_getters(libpath):
    """GetPrimalMatrix (sparse, original indices, LP block with the 2i/2i+1 convention) vs GetPrimalSolutionMatrix (dense) vs
    GetPrimalBoundVars, and dual feasibility of the multipliers: sum_k A_j.X + D'x + w - v = obj_j"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s")).rows_to_bounds()
    bp = sdpisolver_host.BoundaryProblem(M)
    s = sdpisolver_host.SdpiSolver(libpath, gaptol=1e-7, feastol=1e-7)
    try:
        s.load_and_solve(bp)
        assert s.flag("IsOptimal")
        dense = s.primal_matrix_dense(bp)
        sparse = s.primal_matrix_sparse(bp)
        for b, X in enumerate(dense):
            R = np.zeros_like(X)
            r, c, v = sparse[b]
            assert np.all(r >= c)
            R[r, c] = v; R[c, r] = v
            assert np.abs(R - X).max() <= 1e-8
            assert np.linalg.eigvalsh(X).min() >= -1e-8
        lbv, ubv = s.bound_multipliers()
        r, c, v = sparse[-1]
        lp = dict(zip(r.tolist(), v.tolist()))
        nrows = len(M.rows)
        for j in range(M.nvars):
            assert abs(lp.get(2 * nrows + 2 * j, 0.0) - lbv[j]) <= 1e-8 and abs(lp.get(2 * nrows + 2 * j + 1, 0.0) - ubv[j]) <= 1e-8
        # stationarity in the variables: A_j . X + sum_rows (+-d_ij) x + w_j - v_j = obj_j
        for j in range(M.nvars):
            t = lbv[j] - ubv[j]
            for b in range(len(M.blocksizes)):
                for (rr, cc, vv) in M.A[b].get(j, []):
                    t += vv * dense[b][rr, cc] * (1.0 if rr == cc else 2.0)
            for i, (coefs, lhs, rhs) in enumerate(M.rows):
                t += coefs.get(j, 0.0) * (lp.get(2 * i, 0.0) - lp.get(2 * i + 1, 0.0))
            assert abs(t - M.obj[j]) <= 1e-6 * max(1.0, abs(M.obj[j]))
    finally:
        s.close()
