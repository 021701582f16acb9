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
              "example_multaggr.cbf": -1.0, "example_diagzeroimpl.cbf": -1.0, "example_tightenmatrices.dat-s": -9.0}


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
