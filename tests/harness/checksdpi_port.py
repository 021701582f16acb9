"""Port of solveTest() from the reference's unittests/src/checksdpi.c:140-356 onto tests/harness/sdpi_ref.py.
Same assertions in the same order; Criterion's cr_assert becomes a Python assert."""
import numpy as np

from .sdpi_ref import Sdpi

EPS = 1e-6


def run_case(lib, case, name=""):
    before = lib.memory_used()
    s = Sdpi(lib, gaptol=EPS, sdpsolverfeastol=EPS)
    try:
        s.load(case["nvars"], case["obj"], case["lb"], case["ub"], case["blocksizes"], case["A"], case["C"], case["rows"])
        assert not s.flag("WasSolved")
        s.solve()
        assert s.flag("WasSolved"), name
        assert not s.flag("IsObjlimExc") and not s.flag("IsIterlimExc") and not s.flag("IsTimelimExc"), name
        pfeas, dfeas = s.sol_feasibility()
        P, D = case["primal"], case["dual"]
        if P == "feas" and D == "feas":
            assert s.flag("IsOptimal"), name
            assert s.flag("IsDualFeasible") and not s.flag("IsDualInfeasible") and not s.flag("IsDualUnbounded"), name
            assert s.flag("IsPrimalFeasible") and not s.flag("IsPrimalInfeasible") and not s.flag("IsPrimalUnbounded"), name
        if P == "feas":
            assert pfeas and s.flag("IsPrimalFeasible") and not s.flag("IsPrimalInfeasible"), name
            assert not s.flag("IsDualUnbounded"), name
        elif P == "unbounded":
            assert not s.flag("IsPrimalInfeasible"), name
            assert not s.flag("IsDualFeasible") and s.flag("IsDualInfeasible"), name
        elif P == "ray":
            assert not s.flag("IsDualFeasible") and s.flag("IsDualInfeasible"), name
        elif P == "infeas":
            assert not pfeas and not s.flag("IsPrimalFeasible") and not s.flag("IsPrimalUnbounded"), name
        if D == "feas":
            assert dfeas and s.flag("IsDualFeasible") and not s.flag("IsDualInfeasible"), name
            assert not s.flag("IsPrimalUnbounded"), name
        elif D in ("unbounded", "ray"):
            assert not s.flag("IsPrimalFeasible") and s.flag("IsPrimalInfeasible"), name
        elif D == "infeas":
            assert not dfeas and not s.flag("IsDualFeasible") and not s.flag("IsDualUnbounded"), name
        if D == "feas":
            _, y = s.dual_sol()
            np.testing.assert_allclose(y, case["dualsol"], atol=EPS, rtol=0, err_msg=f"{name} dual solution")
        if "lbvals" in case or "ubvals" in case:
            lbv, ubv, ok = s.primal_bound_vars()
            assert ok
            if "lbvals" in case:
                np.testing.assert_allclose(lbv, case["lbvals"], atol=EPS, rtol=0, err_msg=f"{name} lb multipliers")
            if "ubvals" in case:
                np.testing.assert_allclose(ubv, case["ubvals"], atol=EPS, rtol=0, err_msg=f"{name} ub multipliers")
        if "lhsvals" in case or "rhsvals" in case:
            l, r, ok = s.primal_lp_sides()
            assert ok
            if "lhsvals" in case:
                np.testing.assert_allclose(l, case["lhsvals"], atol=EPS, rtol=0, err_msg=f"{name} lhs multipliers")
            if "rhsvals" in case:
                np.testing.assert_allclose(r, case["rhsvals"], atol=EPS, rtol=0, err_msg=f"{name} rhs multipliers")
        if "primalmatrix" in case:
            mats, ok = s.primal_matrices()
            assert ok
            np.testing.assert_allclose(mats[0].ravel(), case["primalmatrix"], atol=EPS, rtol=0, err_msg=f"{name} primal matrix")
        st = s.stats()
    finally:
        s.close()
    assert lib.memory_used() == before, f"{name}: BMS memory leak ({lib.memory_used() - before} bytes)"   # checksdpi.c:117
    return st
