"""CPU suite, part 1: the oracle against the reference's golden vectors.
(a) ported unittests/src/checksdpi.c known answers, driven through the REFERENCE's own sdpi.c + our sdpisolver_cuda.c
    binding + the CPU oracle (oracle/_ref/libsdpi_oracle.so);
(b) B&B optima of check/testset/short.solu for the BASELINE instances, with copies of the four instance files
    regenerated into tests/golden by tests/golden/make_instances.py."""
import os

import numpy as np
import pytest

from golden.checksdpi_cases import CASES
from harness import bnb, checksdpi_port, sdpi_ref
from scip_sdp_b200 import abi, misdp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not os.path.exists(sdpi_ref.LIB_ORACLE), reason="oracle/_ref/libsdpi_oracle.so not built")


@pytest.fixture(scope="module")
def lib():
    os.environ.setdefault("SHIM_QUIET", "1")
    return sdpi_ref.SdpiLib(sdpi_ref.LIB_ORACLE)


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_checksdpi_known_answers(lib, name):
    case = CASES[name]
    if lib.solver_name() in case.get("skip_for", []):
        pytest.skip("skipped upstream for this solver")
    st = checksdpi_port.run_case(lib, case, name)
    if case["reaches_solver"]:
        assert st["sdpcalls"] >= 1


# check/testset/short.solu: every instance of check/testset/short.test that needs neither rank-1 constraints nor indicator
# constraints (those rely on SCIP's own constraint handlers, which the B&B stand-in does not have)
SHORT_SOLU = {"example_small.dat-s": -8.0, "example_inf.dat-s": None, "example_TT.dat-s.gz": 2.11803,
              "example_CLS.dat-s.gz": 7.1485, "example_MkP.dat-s.gz": -95.0,
              "example_small_cbf.cbf": -8.0, "example_cbf_primal.cbf": 0.75, "example_cbf_mix.cbf": 4.0, "example_cbf_dual.cbf": 4.0,
              "example_multaggr.cbf": -1.0, "example_diagzeroimpl.cbf": -1.0, "example_tightenmatrices.dat-s": -9.0,
              "example_small_ind.dat-s": -18.0}      # indicator constraint (binary = 1 => slack = 0), handled by the harness


@needs_ref
@pytest.mark.parametrize("name", sorted(SHORT_SOLU))
def test_bnb_optimum_matches_short_solu(lib, name):
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    r = bnb.solve_misdp(lib, M, timelimit=300)
    if SHORT_SOLU[name] is None:
        assert r["status"] == "infeasible"
    else:
        assert r["status"] == "optimal" and r["unsolved"] == 0
        assert abs(M.file_objective(r["objval"]) - SHORT_SOLU[name]) <= 1e-4 * max(1.0, abs(SHORT_SOLU[name]))


def test_oracle_relaxation_kkt():
    """a-posteriori KKT residuals of the oracle on the root relaxation of example_TT (parity unpinned at relaxation level)"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    fp, _ = M.flatten()
    r = abi.Solver(abi.Lib(abi.ORACLE_LIB)).solve(fp, gaptol=1e-7, feastol=1e-7)
    assert r["phase_name"] == "pdOPT"
    y, X, S = r["y"], r["X"], r["S"]
    C = fp.dense_C()
    Z = [sum(y[j] * fp.dense_A(j)[k] for j in range(fp.m)) - C[k] for k in range(fp.nblocks)]
    for k in range(fp.nblocks):
        assert np.linalg.norm(Z[k] - S[k]) <= 1e-6 * (1 + np.linalg.norm(C[k]))
        assert np.linalg.eigvalsh(X[k]).min() >= -1e-9 and np.linalg.eigvalsh(S[k]).min() >= -1e-9
    D = fp.dense_D()
    AX = np.array([sum(np.vdot(fp.dense_A(j)[k], X[k]) for k in range(fp.nblocks)) for j in range(fp.m)]) + D.T @ r["xlp"]
    assert np.linalg.norm(AX - fp.obj) <= 1e-6 * (1 + np.linalg.norm(fp.obj))
    assert abs(r["pobj"] - r["dobj"]) <= 1e-6 * max(1, abs(r["dobj"]))


def test_oracle_eigen_matches_reference_lapack_interface():
    """the reference's own lapack_interface.c (oracle/_ref/liblapack_ref.so): SURVEY.md section 0 probe values and random matrices"""
    import ctypes as C
    ref_path = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "liblapack_ref.so")
    if not os.path.exists(ref_path):
        pytest.skip("oracle/_ref/liblapack_ref.so not built")
    L = C.CDLL(ref_path, mode=C.RTLD_LOCAL)
    dp = C.POINTER(C.c_double)
    L.BMScreateBufferMemory.restype = C.c_void_p
    L.BMScreateBufferMemory.argtypes = [C.c_double, C.c_int, C.c_uint]
    L.SCIPlapackComputeEigenvectorDecomposition.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
    buf = C.c_void_p(L.BMScreateBufferMemory(1.2, 4, 0))
    cpu = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    rng = np.random.default_rng(0)
    for A in [np.array([[1.0, 2.0], [2.0, 4.0]])] + [(lambda G: G + G.T)(rng.standard_normal((n, n))) for n in (10, 15, 43)]:
        n = A.shape[0]
        Ac, w, V = A.copy(), np.zeros(n), np.zeros(n * n)
        assert L.SCIPlapackComputeEigenvectorDecomposition(buf, n, Ac.ctypes.data_as(dp), w.ctypes.data_as(dp), V.ctypes.data_as(dp)) == 1
        wo, Vo = cpu.syev(A)
        assert np.abs(w - wo).max() <= 1e-10 * max(1.0, np.abs(w).max())
        V = V.reshape(n, n)
        for k in range(n):                      # eigenvectors as rows, equal up to sign where the eigenvalue is simple
            assert abs(abs(V[k] @ Vo[k]) - 1.0) <= 1e-6
    assert np.allclose(w[:0], [])               # (keeps flake8 quiet about w)


@needs_ref
@pytest.mark.parametrize("nfixed", [0, 20])
def test_warmstart_and_preoptimal_through_the_reference_sdpi_layer(lib, nfixed):
    """row a10 driven through the REFERENCE's sdpi.c (SCIPsdpiSolve start point, SCIPsdpiGetPreoptimalSol, sdpi.c:3123-3405,
    4250-4370) on top of our binding: WARMSTARTPOGAP gives an earlier interior iterate; a start point next to the optimum
    (y*, Z*, X* pushed into the cone, as relax_sdp.c does) converges in fewer iterations to the same optimum.
    nfixed = 20: a B&B node with 20 bars fixed to zero, where sdpi.c removes two rows/columns of the block (indchanges), so the
    start matrices in ORIGINAL indices have to be compressed by the binding (and entries of removed rows ignored)"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz"))
    Mb = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    mem_before = lib.memory_used()
    s = sdpi_ref.Sdpi(lib, gaptol=1e-6, sdpsolverfeastol=1e-6, feastol=1e-6)
    try:
        s.load_model(M)
        if nfixed:
            fix = np.flatnonzero(M.integer)[:nfixed]
            Mb.lb[fix] = 0.0; Mb.ub[fix] = 0.0
            s.chg_bounds(np.arange(M.nvars, dtype=np.int32), Mb.lb, Mb.ub)
        s.solve()
        assert s.flag("IsOptimal")
        obj0, y0 = s.dual_sol()
        it_cold = s.stats()["iterations"]
        assert s.preoptimal_sol() is None

        s.set_real("WARMSTARTPOGAP", 1e-2)
        s.solve()
        pre = s.preoptimal_sol()
        assert pre is not None
        ypre, Xpre = pre
        assert np.array_equal(s.preoptimal_y_only(), ypre)
        objpre = float(np.dot(M.obj, ypre))
        assert 1e-7 < abs(objpre - obj0) <= 5e-2 * max(1.0, abs(obj0))
        for b, n in enumerate(M.blocksizes):
            r, c, v = Xpre[b]
            X = np.zeros((n, n)); X[r, c] = v; X[c, r] = v
            kept = np.flatnonzero(np.abs(X).sum(axis=0) > 0)          # removed rows/columns come back as zero rows (original indices)
            assert len(kept) == (n if nfixed == 0 else n - 2)
            assert np.linalg.eigvalsh(X[np.ix_(kept, kept)]).min() > 0.0
        s.set_real("WARMSTARTPOGAP", -1.0)

        # start point: the optimal (y, Z, X) moved 5 % towards the identity
        s.solve()
        Xopt = s.primal_matrix_sparse()
        dense, ok = s.primal_matrices()            # GetPrimalSolutionMatrix: original size, zero rows where sdpi.c removed them
        assert ok
        for b, n in enumerate(M.blocksizes):
            r, c, v = Xopt[b]
            R = np.zeros((n, n)); R[r, c] = v; R[c, r] = v
            assert np.abs(R - dense[b]).max() <= 1e-8
        lam = 0.05
        Zd = M.dense_Z(y0)
        startZ, startX = [], []
        for b, n in enumerate(M.blocksizes):
            r, c, v = Xopt[b]
            X = np.zeros((n, n)); X[r, c] = v; X[c, r] = v
            for A, out in (((1 - lam) * Zd[b] + lam * np.eye(n), startZ), ((1 - lam) * X + lam * np.eye(n), startX)):
                rr, cc = np.nonzero(np.tril(np.ones((n, n))))
                out.append((rr.astype(np.int32), cc.astype(np.int32), A[rr, cc]))
        # LP block: index 2i / 2i+1 for lhs / rhs of the ORIGINAL row i (sdpi.c hands all rows on and marks the removed ones in
        # lpindchanges; rows with one nonzero were turned into bounds, sdpi.c:1131), then 2 nrows + 2j (+1) for lb (ub) of
        # variable j; every finite side and bound gets a positive slack and multiplier
        lp = dict(zip(Xopt[-1][0].tolist(), Xopt[-1][2].tolist()))
        nrows = len(M.rows)
        idx, zval = [], []
        for i, (coefs, lhs, rhs) in enumerate(M.rows):
            if sum(1 for a in coefs.values() if a != 0.0) <= 1:
                continue
            act = sum(a * y0[j] for j, a in coefs.items())
            if lhs > -1e20:
                idx.append(2 * i); zval.append(act - lhs)
            if rhs < 1e20:
                idx.append(2 * i + 1); zval.append(rhs - act)
        for j in range(M.nvars):
            if Mb.ub[j] - Mb.lb[j] <= 1e-9:
                continue                                        # fixed: not a variable of the solver problem
            if Mb.lb[j] > -1e20:
                idx.append(2 * nrows + 2 * j); zval.append(y0[j] - Mb.lb[j])
            if Mb.ub[j] < 1e20:
                idx.append(2 * nrows + 2 * j + 1); zval.append(Mb.ub[j] - y0[j])
        idx = np.array(idx, dtype=np.int32)
        zval = (1 - lam) * np.maximum(np.array(zval), 0.0) + lam
        xval = (1 - lam) * np.array([max(lp.get(int(i), 0.0), 0.0) for i in idx]) + lam
        startZ.append((idx, idx, zval)); startX.append((idx, idx, xval))
        s.solve(starty=y0, startZ=startZ, startX=startX)
        assert s.flag("IsOptimal")
        obj2, _ = s.dual_sol()
        assert abs(obj2 - obj0) <= 1e-5 * max(1.0, abs(obj0))
        assert s.stats()["iterations"] < it_cold, (s.stats()["iterations"], it_cold)
    finally:
        s.close()
    assert lib.memory_used() == mem_before, "BMS memory leak (preoptimal buffers?)"      # like unittests/src/checksdpi.c:117


@needs_ref
def test_exhausted_time_limit_is_not_an_error(lib):
    """time limit already exhausted => SCIP_OKAY, timelimit flag set, nothing acceptable (sdpisolver_dsdp.c:880-890);
    the same object solves normally afterwards"""
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz"))
    s = sdpi_ref.Sdpi(lib, gaptol=1e-6, sdpsolverfeastol=1e-6, feastol=1e-6)
    try:
        s.load_model(M)
        s.solve(timelimit=1e-12)
        assert s.flag("IsTimelimExc") and not s.flag("IsAcceptable") and not s.flag("IsConverged")
        assert s.stats()["sdpcalls"] == 0
        s.solve()
        assert not s.flag("IsTimelimExc") and s.flag("IsOptimal")
        obj, _ = s.dual_sol()
        assert abs(obj - 0.16447) <= 1e-4          # root relaxation of example_TT
    finally:
        s.close()


@needs_ref
@pytest.mark.parametrize("name,dual_expected", [("example_small.dat-s", 1), ("example_TT.dat-s.gz", 1), ("example_MkP.dat-s.gz", 1),
                                                ("example_CLS.dat-s.gz", 1), ("example_inf.dat-s", 1), ("example_inf.dat-s:relaxation-infeasible", -2)])
def test_slater_checks_of_the_reference_run_through_the_binding(lib, name, dual_expected):
    """SURVEY.md 8b invocation patterns (iv) and (v): sdpi.c's dual Slater check (penalty formulation, r free, no objective) and
    primal Slater check (LoadAndSolve without constant matrices, all sides 0, one extra LP row; sdpi.c:1518-1870) with
    relaxing/SDP/slatercheck = 2: SCIP_SDPSLATER_HOLDS (1) for the instances with a strictly feasible relaxation (example_inf is
    infeasible only in the integers: y1 y2 >= 4/3 with |y| <= sqrt 2), SCIP_SDPSLATER_INF (-2) on the dual side of a variant of
    example_inf whose first block loses its constant entry (1,1) (then S_11 = 0 forces y1 = 0 and block 2 cannot be psd);
    the main solve is unaffected"""
    import ctypes as C
    name, _, variant = name.partition(":")
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    if variant:
        assert M.C[0][0] == (0, 0, -1.0)
        del M.C[0][0]
    s = sdpi_ref.Sdpi(lib, gaptol=1e-6, sdpsolverfeastol=1e-6, feastol=1e-6)
    try:
        s.load_model(M)
        assert s.L.lib.SCIPsdpiSetIntpar(s.sdpi, sdpi_ref.PAR["SLATERCHECK"], 2) == sdpi_ref.SCIP_OKAY
        s.solve(enforceslater=True)
        p, d = C.c_int(-9), C.c_int(-9)
        s.L.lib.SCIPsdpiSlater.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        assert s.L.lib.SCIPsdpiSlater(s.sdpi, C.byref(p), C.byref(d)) == sdpi_ref.SCIP_OKAY
        assert p.value == 1 and d.value == dual_expected
        assert s.flag("IsAcceptable")
        assert s.flag("IsDualInfeasible") == (dual_expected == -2)
    finally:
        s.close()


# unittests/src/check1dsdp.c:122-400: one-variable SDPs  min x  s.t.  A x - B psd, lb <= x <= ub  (lower triangles, 0-based)
CHECK1D = {
    "test1": dict(lb=0.0, ub=1.0, n=2, B=[(0, 0, 1.0)], A=[(0, 0, 1.0)], x=1.0),
    "test2": dict(lb=0.0, ub=1.5, n=2, B=[(0, 0, 1.0), (1, 1, -1.0)], A=[(0, 0, 1.0), (1, 1, -1.0)], x=1.0),
    "test3": dict(lb=0.0, ub=2.0, n=2, B=[(0, 0, 1.0), (1, 1, -0.99)], A=[(0, 0, 1.0), (1, 1, -1.0)], x=None),
    "test4": dict(lb=0.0, ub=2.0, n=2, B=[(0, 0, 0.89496), (1, 0, -0.44498), (1, 1, -0.88496)],
                  A=[(0, 0, 0.89443), (1, 0, -0.44721), (1, 1, -0.89443)], x=None),
    "test5": dict(lb=0.0, ub=2.0, n=2, B=[(0, 0, -2.0), (1, 0, 1.0), (1, 1, 3.0)], A=[(0, 0, 1.0), (1, 0, -2.0), (1, 1, 5.0)], x=1.541381),
}


@needs_ref
@pytest.mark.parametrize("name", sorted(CHECK1D))
def test_check1dsdp_known_answers(lib, name):
    """the five known answers of unittests/src/check1dsdp.c: (a) the reference's own SCIPsolveOneVarSDP (solveonevarsdp.c,
    compiled in place; the shortcut sdpi.c takes for one-variable problems), (b) where the problem has an interior, the same
    answer from the interior-point oracle through the C ABI (this is what the GPU path is compared with)"""
    import ctypes as C
    c = CHECK1D[name]
    L = lib.lib
    _dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.SCIPsolveOneVarSDP.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, _ip, _ip, _dp,
                                     C.c_double, C.c_double, _dp, _dp, _dp, _dp]
    buf = C.c_void_p(L.BMScreateBufferMemory(1.2, 4, 0))

    def arrs(ents):
        r = (C.c_int * len(ents))(*[e[0] for e in ents]); cc = (C.c_int * len(ents))(*[e[1] for e in ents])
        v = (C.c_double * len(ents))(*[e[2] for e in ents])
        return len(ents), r, cc, v
    nb, br, bc, bv = arrs(c["B"])
    na, ar, ac, av = arrs(c["A"])
    objval, optval = C.c_double(0), C.c_double(0)
    rc = L.SCIPsolveOneVarSDP(buf, 1.0, c["lb"], c["ub"], c["n"], nb, br, bc, bv, na, ar, ac, av, 1e20, 1e-6, None, None,
                              C.byref(objval), C.byref(optval))
    assert rc == sdpi_ref.SCIP_OKAY
    L.BMSdestroyBufferMemory(C.byref(buf))
    if c["x"] is None:
        assert objval.value >= 1e20                        # infeasible (check1dsdp.c:306, 334)
    else:
        assert abs(optval.value - c["x"]) <= 1e-6
    # (b) the same problem through the interior-point oracle
    M = misdp.Misdp(1, [1.0], [c["n"]])
    M.A[0][0] = list(c["A"]); M.C[0] = list(c["B"])
    M.lb[0], M.ub[0] = c["lb"], c["ub"]
    fp, _ = M.flatten()
    r = abi.Solver(abi.Lib(abi.ORACLE_LIB)).solve(fp, gaptol=1e-8, feastol=1e-8)
    if name == "test5":
        assert r["phase_name"] == "pdOPT" and abs(r["y"][0] - c["x"]) <= 1e-5
    elif c["x"] is None:
        assert r["phase_name"] in ("pFEAS_dINF", "dINF", "noINFO", "pFEAS")        # never reported optimal
