"""CPU suite, part 9: the host-side entry point of sdpi/lapack_cuda.c — SCIPlapackLinearSolve (one-sided Jacobi SVD) — against the
REFERENCE's own lapack_interface.c (DGELSD, lapack_interface.c:712-822; compiled unmodified into oracle/_ref/liblapack_ref.so)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "scip-sdp_b200", "lib", "libsdpisolver_cuda.so")
REF = os.path.join(ROOT, "oracle", "_ref", "liblapack_ref.so")
_dp = C.POINTER(C.c_double)
pytestmark = pytest.mark.skipif(not (os.path.exists(OURS) and os.path.exists(REF)), reason="libraries not built")


def _load(path):
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    L.BMScreateBufferMemory.restype = C.c_void_p
    L.BMScreateBufferMemory.argtypes = [C.c_double, C.c_int, C.c_uint]
    L.SCIPlapackLinearSolve.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
    return L, C.c_void_p(L.BMScreateBufferMemory(1.2, 4, 0))


def _solve(lib, A, b):
    L, buf = lib
    m, n = A.shape
    Af = np.asfortranarray(A).ravel(order="F").copy()          # column-major
    bb, x = b.astype(float).copy(), np.zeros(n)
    assert L.SCIPlapackLinearSolve(buf, m, n, Af.ctypes.data_as(_dp), bb.ctypes.data_as(_dp), x.ctypes.data_as(_dp)) == 1
    return x


@pytest.fixture(scope="module")
def libs():
    return _load(OURS), _load(REF)


@pytest.mark.parametrize("m,n", [(1, 1), (3, 3), (8, 5), (20, 20), (43, 12)])
def test_linear_solve_matches_dgelsd(libs, m, n):
    rng = np.random.default_rng(100 * m + n)
    A = rng.standard_normal((m, n)); b = rng.standard_normal(m)
    ours, ref = _solve(libs[0], A, b), _solve(libs[1], A, b)
    assert np.abs(ours - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    assert np.abs(ours - np.linalg.lstsq(A, b, rcond=None)[0]).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def test_linear_solve_rank_deficient_gives_the_minimum_norm_solution(libs):
    rng = np.random.default_rng(5)
    B = rng.standard_normal((9, 3))
    A = B @ rng.standard_normal((3, 6))                 # rank 3, six unknowns
    b = A @ rng.standard_normal(6)
    ours, ref = _solve(libs[0], A, b), _solve(libs[1], A, b)
    assert np.abs(ours - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    assert np.abs(A @ ours - b).max() <= 1e-9 and np.linalg.norm(ours) <= np.linalg.norm(ref) * (1 + 1e-9)


def test_linear_solve_ill_conditioned(libs):
    """cond(A) = 1e9: the normal equations (cond 1e18) lose everything, an SVD-grade method keeps 1e-7"""
    rng = np.random.default_rng(6)
    Q1, _ = np.linalg.qr(rng.standard_normal((12, 12)))
    Q2, _ = np.linalg.qr(rng.standard_normal((8, 8)))
    A = Q1[:, :8] @ np.diag(np.logspace(0, -9, 8)) @ Q2.T
    xtrue = rng.standard_normal(8)
    b = A @ xtrue
    ours, ref = _solve(libs[0], A, b), _solve(libs[1], A, b)
    assert np.abs(ours - xtrue).max() <= 1e-6 * np.abs(xtrue).max()
    assert np.abs(ours - ref).max() <= 1e-6 * np.abs(xtrue).max()
