"""GPU suite, part 4: the SCIPlapack* entry points of lib/libsdpisolver_cuda.so (sdpi/lapack_cuda.c, batched Jacobi kernel) against
the REFERENCE's own src/sdpi/lapack_interface.c (compiled unmodified into oracle/_ref/liblapack_ref.so) on the block sizes of the
shipped instances (10, 15, 43), and the reference's DGEMM known answer (unittests/src/checklapack.c:73-91)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "scip-sdp_b200", "lib", "libsdpisolver_cuda.so")
REF = os.path.join(ROOT, "oracle", "_ref", "liblapack_ref.so")
_dp = C.POINTER(C.c_double)


def _load(path):
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    L.BMScreateBufferMemory.restype = C.c_void_p
    L.BMScreateBufferMemory.argtypes = [C.c_double, C.c_int, C.c_uint]
    L.SCIPlapackComputeIthEigenvalue.argtypes = [C.c_void_p, C.c_uint, C.c_int, _dp, C.c_int, _dp, _dp]
    L.SCIPlapackComputeEigenvectorsNegative.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.POINTER(C.c_int), _dp, _dp]
    L.SCIPlapackComputeEigenvectorDecomposition.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    L.SCIPlapackMatrixMatrixMult.argtypes = [C.c_int, C.c_int, _dp, C.c_uint, C.c_int, C.c_int, _dp, C.c_uint, _dp]
    L.SCIPlapackMatrixVectorMult.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
    return L, C.c_void_p(L.BMScreateBufferMemory(1.2, 4, 0))


@pytest.fixture(scope="module")
def libs():
    return _load(OURS), _load(REF)


def _p(a):
    return a.ctypes.data_as(_dp)


@pytest.mark.parametrize("n", [2, 10, 15, 43, 64])
def test_eigen_entry_points_match_reference_lapack_interface(libs, n):
    (ours, bo), (ref, br) = libs
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)); A = A + A.T - 0.5 * np.eye(n)
    nrm = np.abs(np.linalg.eigvalsh(A)).max()
    # i-th eigenvalue (1-based) with eigenvector
    for i in (1, n):
        out = []
        for L, b in ((ours, bo), (ref, br)):
            Ac, ev, vec = A.copy(), C.c_double(0), np.zeros(n)
            assert L.SCIPlapackComputeIthEigenvalue(b, 1, n, _p(Ac), i, C.byref(ev), _p(vec)) == 1
            out.append((ev.value, vec))
        assert abs(out[0][0] - out[1][0]) <= 1e-10 * nrm
        assert np.linalg.norm(A @ out[0][1] - out[0][0] * out[0][1]) <= 1e-9 * nrm
        assert abs(abs(np.dot(out[0][1], out[1][1])) - 1.0) <= 1e-8            # same eigenvector up to sign
    # negative eigenpairs (the default cut path, cons_sdp.c:1699)
    res = []
    for L, b in ((ours, bo), (ref, br)):
        Ac, cnt, w, V = A.copy(), C.c_int(0), np.zeros(n), np.zeros(n * n)
        assert L.SCIPlapackComputeEigenvectorsNegative(b, n, _p(Ac), 1e-6, C.byref(cnt), _p(w), _p(V)) == 1
        res.append((cnt.value, w[:cnt.value].copy(), V.reshape(n, n)[:cnt.value].copy()))
    assert res[0][0] == res[1][0] and res[0][0] > 0
    assert np.abs(res[0][1] - res[1][1]).max() <= 1e-10 * nrm
    for k in range(res[0][0]):
        v = res[0][2][k]                                                        # eigenvector k is ROW k
        assert np.linalg.norm(A @ v - res[0][1][k] * v) <= 1e-9 * nrm
        assert v @ A @ v < 0                                                    # a violated cut direction
    # full decomposition
    full = []
    for L, b in ((ours, bo), (ref, br)):
        Ac, w, V = A.copy(), np.zeros(n), np.zeros(n * n)
        assert L.SCIPlapackComputeEigenvectorDecomposition(b, n, _p(Ac), _p(w), _p(V)) == 1
        full.append((w, V.reshape(n, n)))
    assert np.abs(full[0][0] - full[1][0]).max() <= 1e-10 * nrm
    assert np.abs(full[0][1] @ full[0][1].T - np.eye(n)).max() <= 1e-10 * n


def test_checklapack_known_answer(libs):
    (ours, _), (ref, _) = libs
    A = np.array([1.0, 2.0, 3.0, 4.0]); B = np.array([5.0, 6.0, 7.0, 8.0])
    for L in (ours, ref):
        out = np.zeros(4)
        assert L.SCIPlapackMatrixMatrixMult(2, 2, _p(A), 0, 2, 2, _p(B), 1, _p(out)) == 1
        assert np.allclose(out, [26.0, 38.0, 30.0, 44.0])
