"""CPU suite, part 4: SCIPsdpiSolver* boundary checks of our binding linked to the CPU oracle (oracle/_ref/libsdpi_oracle.so)."""
import os

import pytest

from harness import boundary_cases, sdpi_ref

pytestmark = pytest.mark.skipif(not os.path.exists(sdpi_ref.LIB_ORACLE), reason="oracle/_ref/libsdpi_oracle.so not built")


def test_penalty_formulation_call_patterns():
    boundary_cases.run_penalty_patterns(sdpi_ref.LIB_ORACLE)


def test_primal_matrix_getters_are_consistent():
    boundary_cases.run_primal_getters(sdpi_ref.LIB_ORACLE)


def test_resolves_of_a_node_reuse_the_resident_problem():
    boundary_cases.run_resident_resolves(sdpi_ref.LIB_ORACLE)


def test_warmstart_and_preoptimal_solution():
    boundary_cases.run_warmstart_and_preoptimal(sdpi_ref.LIB_ORACLE)


def test_parameters_and_penalty_formulas():
    """row a11: Get/SetRealpar, Get/SetIntpar, Infinity and the penalty-parameter policy (same constants as the SDPA binding,
    sdpisolver_sdpa.cpp:101-105: Gamma = clamp(10 maxcoeff, 1e5, 1e12), max Gamma = min(1e6 Gamma, 1e15))"""
    import ctypes as C
    from scip_sdp_b200 import sdpisolver_host
    s = sdpisolver_host.SdpiSolver(sdpi_ref.LIB_ORACLE)
    L = s.lib
    try:
        L.SCIPsdpiSolverGetRealpar.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.SCIPsdpiSolverGetIntpar.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.SCIPsdpiSolverSetIntpar.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.SCIPsdpiSolverInfinity.restype = C.c_double
        L.SCIPsdpiSolverInfinity.argtypes = [C.c_void_p]
        L.SCIPsdpiSolverIsInfinity.restype = C.c_uint
        L.SCIPsdpiSolverIsInfinity.argtypes = [C.c_void_p, C.c_double]
        L.SCIPsdpiSolverComputePenaltyparam.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        L.SCIPsdpiSolverComputeMaxPenaltyparam.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        L.SCIPsdpiSolverGetDefaultSdpiSolverNpenaltyIncreases.restype = C.c_int
        v = C.c_double(0)
        # EPSILON 0, GAPTOL 1, FEASTOL 2, SDPSOLVERFEASTOL 3, OBJLIMIT 4, LAMBDASTAR 10, WARMSTARTPOGAP 12 (type_sdpi.h:49-66)
        for par, val in ((0, 1e-8), (1, 3e-5), (2, 2e-6), (3, 4e-7), (4, 123.5), (10, 7.0), (12, 0.01)):
            assert L.SCIPsdpiSolverSetRealpar(s.s, par, val) == sdpisolver_host.SCIP_OKAY
            assert L.SCIPsdpiSolverGetRealpar(s.s, par, C.byref(v)) == sdpisolver_host.SCIP_OKAY and v.value == val
        for par in (14, 15, 16):                  # USEPRESOLVING, USESCALING, SCALEOBJ: unknown like in the DSDP/SDPA bindings
            assert L.SCIPsdpiSolverSetRealpar(s.s, par, 1.0) == -12 or L.SCIPsdpiSolverSetIntpar(s.s, par, 1) == -12
        iv = C.c_int(-7)
        assert L.SCIPsdpiSolverSetIntpar(s.s, 5, 1) == sdpisolver_host.SCIP_OKAY            # SDPINFO
        assert L.SCIPsdpiSolverGetIntpar(s.s, 5, C.byref(iv)) == sdpisolver_host.SCIP_OKAY and iv.value == 1
        assert L.SCIPsdpiSolverSetIntpar(s.s, 5, 0) == sdpisolver_host.SCIP_OKAY
        assert L.SCIPsdpiSolverSetIntpar(s.s, 11, 4) == sdpisolver_host.SCIP_OKAY           # NTHREADS: accepted, no effect
        assert L.SCIPsdpiSolverInfinity(s.s) == 1e20
        assert L.SCIPsdpiSolverIsInfinity(s.s, 1e20) and L.SCIPsdpiSolverIsInfinity(s.s, -3e20) and not L.SCIPsdpiSolverIsInfinity(s.s, 9e19)
        for maxcoeff, expect in ((1.0, 1e5), (5e4, 5e5), (1e13, 1e12)):
            assert L.SCIPsdpiSolverComputePenaltyparam(s.s, maxcoeff, C.byref(v)) == sdpisolver_host.SCIP_OKAY and v.value == expect
        for gamma, expect in ((1e5, 1e11), (1e10, 1e15)):
            assert L.SCIPsdpiSolverComputeMaxPenaltyparam(s.s, gamma, C.byref(v)) == sdpisolver_host.SCIP_OKAY and v.value == expect
        assert L.SCIPsdpiSolverGetDefaultSdpiSolverNpenaltyIncreases() == 8
        assert s.name() == "CUDA-IPM"
    finally:
        s.close()


def test_getters_before_a_solve_return_lperror():
    """error behaviour of the boundary: SCIP_LPERROR for "asked for a solution before a solve" (CHECK_IF_SOLVED,
    sdpisolver_dsdp.c:145-164) and for the unimplemented file interface, never a crash"""
    import ctypes as C
    import numpy as np
    from scip_sdp_b200 import sdpisolver_host
    SCIP_LPERROR = -6
    s = sdpisolver_host.SdpiSolver(sdpi_ref.LIB_ORACLE)
    L = s.lib
    try:
        assert not s.flag("WasSolved")
        v = C.c_double(0)
        assert L.SCIPsdpiSolverGetObjval(s.s, C.byref(v)) == SCIP_LPERROR
        y = np.zeros(4)
        assert L.SCIPsdpiSolverGetDualSol(s.s, C.byref(v), y.ctypes.data_as(C.POINTER(C.c_double))) == SCIP_LPERROR
        cnt = (C.c_int * 2)()
        assert L.SCIPsdpiSolverGetPrimalNonzeros(s.s, 2, cnt) == SCIP_LPERROR
        L.SCIPsdpiSolverReadSDP.argtypes = [C.c_void_p, C.c_char_p]
        L.SCIPsdpiSolverWriteSDP.argtypes = [C.c_void_p, C.c_char_p]
        assert L.SCIPsdpiSolverReadSDP(s.s, b"x.dat-s") == SCIP_LPERROR and L.SCIPsdpiSolverWriteSDP(s.s, b"x.dat-s") == SCIP_LPERROR
        it = C.c_int(-1)
        assert L.SCIPsdpiSolverGetIterations(s.s, C.byref(it)) in (sdpisolver_host.SCIP_OKAY, SCIP_LPERROR)
    finally:
        s.close()


def _resident_check_cases(solver, tol=1e-9):
    """sdpcuda_check_psd_resident against numpy's eigenvalues: the solution of the last solve, shifted copies of it and random
    vectors, on instances with one and with several blocks (shared by the CPU-oracle and the GPU suite)"""
    import os
    import numpy as np
    from scip_sdp_b200 import generators, misdp
    golden = os.path.join(os.path.dirname(__file__), "golden")
    rng = np.random.default_rng(5)
    for make in (lambda: misdp.read_sdpa(os.path.join(golden, "example_small.dat-s")).rows_to_bounds(),
                 lambda: misdp.read_sdpa(os.path.join(golden, "example_MkP.dat-s.gz")).rows_to_bounds(),
                 lambda: generators.cls(12, 9, 3, seed=5), lambda: generators.maxcut(150, 0.05, seed=11)):
        M = make()
        fp, _ = M.flatten()
        r = solver.solve(fp, gaptol=1e-7, feastol=1e-7)
        assert r["phase_name"] == "pdOPT"

        def lam_min(y):
            Cd = fp.dense_C()
            Z = [sum(y[j] * fp.dense_A(j)[k] for j in range(fp.m)) - Cd[k] for k in range(fp.nblocks)]
            return min(float(np.linalg.eigvalsh(z)[0]) for z in Z)

        lam = lam_min(r["y"])
        assert lam >= -1e-6                                     # the solution is (nearly) feasible
        assert solver.check_psd_resident(None, shift=1e-5)      # ... and the device says so for the y it holds
        assert solver.check_psd_resident(r["y"], shift=1e-5)
        for _ in range(4):
            y = r["y"] + rng.standard_normal(fp.m) * 0.3 * max(1.0, np.abs(r["y"]).max())
            lam = lam_min(y)
            for shift in (0.0, -lam + 1e-3 * max(1.0, abs(lam)), -lam - 1e-3 * max(1.0, abs(lam))):
                if abs(lam + shift) > tol:
                    assert solver.check_psd_resident(y, shift=shift) == (lam + shift > 0), (lam, shift)


def test_resident_psd_check_on_the_oracle():
    from scip_sdp_b200 import abi
    s = abi.Solver(abi.Lib(abi.ORACLE_LIB))
    import ctypes as C
    ok = C.c_int(0)
    assert s.L.lib.sdpcuda_check_psd_resident(s.h, None, 0.0, C.byref(ok)) == 4      # SDPCUDA_ERR_STATE: nothing loaded
    assert s.L.lib.sdpcuda_check_psd_resident(s.h, None, 0.0, None) == 1             # SDPCUDA_ERR_ARG
    _resident_check_cases(s)


def test_conflict_cut_reductions_on_the_resident_primal_solution():
    boundary_cases.run_primal_inner_products(sdpi_ref.LIB_ORACLE)
