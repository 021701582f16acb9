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


def run_resident_resolves(libpath, device=False):
    """the re-solves of one node (sdpi.c:3437-3619) reuse the problem that is resident on the device.  Pattern: node
    solve, the same node again with a tighter gap tolerance, then the penalty formulation (iii) with three growing penalty
    parameters — uploads happen only when the structure changes (node -> penalty formulation), every other solve ships obj and
    lprhs only; the results equal those of a fresh solver object that uploads every time."""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    bp = sdpisolver_host.BoundaryProblem(M)
    s = sdpisolver_host.SdpiSolver(libpath, gaptol=1e-5, feastol=1e-6)
    try:
        s.load_and_solve(bp)
        obj0, y0 = s.dual_sol()
        assert s.transfer_stats()[:2] == (1, 0)
        bytes_first = s.transfer_stats()[2]
        s.lib.SCIPsdpiSolverSetRealpar(s.s, 1, 1e-7)                   # GAPTOL tightened between two solves of the node (sdpi.c:3553)
        s.load_and_solve(bp)
        obj1, _ = s.dual_sol()
        up, pa, by = s.transfer_stats()
        assert (up, pa) == (1, 1) and abs(obj1 - obj0) <= 1e-5 * max(1.0, abs(obj0))
        if device:
            assert by - bytes_first <= 0.3 * bytes_first, (by, bytes_first)              # obj + lprhs only (the first solve shipped every array)
        objs = []
        for gamma in (1e4, 1e5, 1e6):
            feasorig, _ = s.load_and_solve_with_penalty(bp, gamma, True, True)
            assert s.flag("IsAcceptable") and feasorig
            objs.append(s.dual_sol()[0])
        up, pa, by2 = s.transfer_stats()
        assert up == 2 and pa >= 3, (up, pa)                          # one upload for the penalty structure, Gamma only changes obj
        fresh = []
        for gamma in (1e4, 1e5, 1e6):
            t = sdpisolver_host.SdpiSolver(libpath, gaptol=1e-7, feastol=1e-6)
            t.load_and_solve_with_penalty(bp, gamma, True, True)
            fresh.append(t.dual_sol()[0])
            assert t.transfer_stats()[0] == 1
            t.close()
        assert np.allclose(objs, fresh, rtol=1e-9, atol=1e-9), (objs, fresh)
        # a different node (one more fixed variable): the structure changes, so it is uploaded
        lb, ub = M.lb.copy(), M.ub.copy()
        j = int(np.flatnonzero(M.integer)[0]); ub[j] = lb[j]
        s.load_and_solve(sdpisolver_host.BoundaryProblem(M, lb, ub))
        assert s.transfer_stats()[0] == 3
    finally:
        s.close()


def run_primal_inner_products(libpath):
    """the quantities computeConflictCut (relax_sdp.c:1030-1099) forms from the dense primal matrices — <A_v, X> per block
    variable, <A_0, X>, min(lambda_min(X), 0) — from the device-resident X, against the same sums formed on the host from
    SCIPsdpiSolverGetPrimalSolutionMatrix (the reference's route)"""
    for name, fix in (("example_small.dat-s", 0), ("example_TT.dat-s.gz", 3), ("example_MkP.dat-s.gz", 5)):
        M = misdp.read_sdpa(os.path.join(GOLDEN, name)).rows_to_bounds()
        lb, ub = M.lb.copy(), M.ub.copy()
        for j in np.flatnonzero(M.integer)[:fix]:
            ub[j] = lb[j]                                   # a node with fixed variables (their matrices leave the device problem)
        bp = sdpisolver_host.BoundaryProblem(M, lb, ub)
        s = sdpisolver_host.SdpiSolver(libpath, gaptol=1e-6, feastol=1e-6)
        try:
            s.load_and_solve(bp)
            assert s.flag("WasSolved")
            dense = s.primal_matrix_dense(bp)
            prods, const, mineig = s.primal_inner_products(bp)
            for b, X in enumerate(dense):
                scale = max(1.0, np.abs(X).max())
                vs = sorted(M.A[b])
                want = [sum(v * X[r, c] * (1.0 if r == c else 2.0) for r, c, v in M.A[b][j]) for j in vs]
                assert len(prods[b]) == len(vs)
                assert np.allclose(prods[b], want, rtol=0, atol=1e-10 * scale * max(1.0, max(abs(v) for j in vs for _, _, v in M.A[b][j])))
                wantc = sum(v * X[r, c] * (1.0 if r == c else 2.0) for r, c, v in M.C[b])
                assert abs(const[b] - wantc) <= 1e-10 * scale * max(1.0, max([abs(v) for _, _, v in M.C[b]] + [1.0]))
                lam = np.linalg.eigvalsh(X).min()
                assert mineig[b] <= min(lam, 0.0) + 1e-12 * scale and mineig[b] >= min(lam, 0.0) - 1e-10 * scale - 10 * abs(min(lam, 0.0))
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


def _sparse_lower(A, tol=0.0):
    r, c = np.nonzero(np.tril(np.abs(A) > tol))
    return r.astype(np.int32), c.astype(np.int32), A[r, c]


def run_warmstart_and_preoptimal(libpath):
    """row a10 of the scope table: (1) WARMSTARTPOGAP makes the solver keep the first iterate inside that gap and
    GetPreoptimalSol returns it (y with fixed variables filled, X sparse in original indices, LP block with the 2i/2i+1
    convention); (2) a primal-dual start point (y, Z, X) is used: starting next to the optimum needs fewer iterations than the
    cold start and ends at the same optimum; (3) a start point outside the cone is ignored, not fatal."""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    bp = sdpisolver_host.BoundaryProblem(M)
    s = sdpisolver_host.SdpiSolver(libpath, gaptol=1e-6, feastol=1e-6)
    try:
        assert s.lib.SCIPsdpiSolverDoesWarmstartNeedPrimal()
        s.load_and_solve(bp)
        assert s.flag("IsOptimal")
        assert s.preoptimal_sol(bp) is None                     # not asked for: first entry -1 / success FALSE
        obj0, y0 = s.dual_sol()
        it_cold, _ = s.iterations()

        s.set_warmstart_pogap(1e-2)
        s.load_and_solve(bp)
        assert s.flag("IsOptimal")
        obj1, y1 = s.dual_sol()
        assert abs(obj1 - obj0) <= 1e-6 * max(1.0, abs(obj0))
        pre = s.preoptimal_sol(bp)
        assert pre is not None
        ypre, Xpre = pre
        objpre = float(np.dot(M.obj, ypre))
        # an earlier iterate: close to the optimum within the requested gap, but not the final point
        assert abs(objpre - obj0) <= 5e-2 * max(1.0, abs(obj0)) and abs(objpre - obj0) > 1e-7
        for b, n in enumerate(M.blocksizes):
            r, c, v = Xpre[b]
            assert np.all(r >= c) and np.all(r < n)
            X = np.zeros((n, n)); X[r, c] = v; X[c, r] = v
            assert np.linalg.eigvalsh(X).min() > 0.0             # an interior iterate
        assert np.all(Xpre[-1][2] > 0.0)                         # LP multipliers of an interior iterate
        s.set_warmstart_pogap(-1.0)

        # start point next to the optimum, pushed into the cone like relax_sdp.c does (convex combination with the identity)
        dense = s.primal_matrix_dense(bp)
        lp = s.primal_matrix_sparse(bp)[-1]
        lam = 0.05
        Zs, Xs = [], []
        Zopt = M.dense_Z(y0)
        for b, n in enumerate(M.blocksizes):
            Zs.append(_sparse_lower((1 - lam) * Zopt[b] + lam * np.eye(n)))
            Xs.append(_sparse_lower((1 - lam) * dense[b] + lam * np.eye(n)))
        # LP block: slack of every finite side / bound (Z part) and its multiplier (X part), all made positive
        idx, zval, xval = [], [], []
        xmap = dict(zip(lp[0].tolist(), lp[2].tolist()))
        nrows = len(M.rows)
        for i, (coefs, lhs, rhs) in enumerate(M.rows):
            act = sum(a * y0[j] for j, a in coefs.items())
            if lhs > -1e20:
                idx.append(2 * i); zval.append(act - lhs)
            if rhs < 1e20:
                idx.append(2 * i + 1); zval.append(rhs - act)
        for j in range(M.nvars):
            if M.lb[j] > -1e20:
                idx.append(2 * nrows + 2 * j); zval.append(y0[j] - M.lb[j])
            if M.ub[j] < 1e20:
                idx.append(2 * nrows + 2 * j + 1); zval.append(M.ub[j] - y0[j])
        idx = np.array(idx, dtype=np.int32)
        zval = (1 - lam) * np.maximum(np.array(zval), 0.0) + lam
        xval = (1 - lam) * np.array([max(xmap.get(int(i), 0.0), 0.0) for i in idx]) + lam
        Zs.append((idx, idx, zval)); Xs.append((idx, idx, xval))
        s.load_and_solve(bp, start=dict(y=y0, Z=Zs, X=Xs))
        assert s.flag("IsOptimal")
        obj2, _ = s.dual_sol()
        it_warm, _ = s.iterations()
        assert abs(obj2 - obj0) <= 1e-5 * max(1.0, abs(obj0))
        assert it_warm < it_cold, (it_warm, it_cold)

        # a start point outside the cone: solved from the default point instead
        Xbad = [(r, c, -v) for (r, c, v) in Xs]
        s.load_and_solve(bp, start=dict(y=y0, Z=Zs, X=Xbad))
        assert s.flag("IsOptimal")
        obj3, _ = s.dual_sol()
        assert abs(obj3 - obj0) <= 1e-5 * max(1.0, abs(obj0))
    finally:
        s.close()
