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


@pytest.mark.parametrize("n", [96, 97, 120, 200])
def test_eigen_entry_points_on_larger_blocks(libs, n):
    """orders above the shared-memory limit of the Jacobi kernel (96) take its global-memory branch"""
    (ours, bo), (ref, br) = libs
    rng = np.random.default_rng(9000 + n)
    A = rng.standard_normal((n, n)); A = A + A.T
    nrm = np.abs(np.linalg.eigvalsh(A)).max()
    full = []
    for L, b in ((ours, bo), (ref, br)):
        Ac, w, V = A.copy(), np.zeros(n), np.zeros(n * n)
        assert L.SCIPlapackComputeEigenvectorDecomposition(b, n, _p(Ac), _p(w), _p(V)) == 1
        full.append((w, V.reshape(n, n)))
    assert np.abs(full[0][0] - full[1][0]).max() <= 1e-10 * nrm
    V = full[0][1]
    assert np.abs(V @ V.T - np.eye(n)).max() <= 1e-10 * n
    assert np.abs(A @ V.T - V.T * full[0][0][None, :]).max() <= 1e-10 * nrm * n


def _cuts(L, buf, Z, Aj, A0, tol, batch=None):
    """what separateSol / produceCutFromEigenvector form (cons_sdp.c:1612-1797, :896-1130) from the negative eigenpairs of Z(y):
    for every eigenvector v the cut  sum_j (v' A_j v) y_j >= v' A_0 v  -> (eigenvalues, coefficient rows, left-hand sides)"""
    n = Z.shape[0]
    if batch is None:
        Ac, cnt, w, V = Z.copy(), C.c_int(0), np.zeros(n), np.zeros(n * n)
        assert L.SCIPlapackComputeEigenvectorsNegative(buf, n, _p(Ac), tol, C.byref(cnt), _p(w), _p(V)) == 1
        k = cnt.value
    else:
        k, w, V = batch
    V = V.reshape(n, n)[:k]
    coef = np.array([[v @ A @ v for A in Aj] for v in V]).reshape(k, len(Aj))
    lhs = np.array([v @ A0 @ v for v in V])
    return w[:k].copy(), coef, lhs


def _dense_blocks(M, y):
    """(Z(y), [A_j], A_0) per SDP block of a Misdp, dense symmetric"""
    out = []
    for b, n in enumerate(M.blocksizes):
        def dense(ents):
            D = np.zeros((n, n))
            for r, c, v in ents:
                D[r, c] = v; D[c, r] = v
            return D
        Aj = [dense(M.A[b].get(j, [])) for j in range(M.nvars)]
        A0 = dense(M.C[b])
        out.append((sum(yj * A for yj, A in zip(y, Aj)) - A0, Aj, A0))
    return out


@pytest.mark.parametrize("name", ["example_small.dat-s", "example_TT.dat-s.gz", "example_MkP.dat-s.gz", "example_CLS.dat-s.gz"])
def test_eigenvector_cuts_match_the_reference_lapack_path(libs, name):
    """the cuts cons_sdp.c forms from our eigenvectors equal those from the reference's lapack_interface.c —
    coefficient vectors and left-hand sides to 1e-8 (sums over clusters of equal eigenvalues, which are rotation invariant),
    every cut violated by exactly its eigenvalue; and the batch entry point (one device call for all blocks of a separation
    round) returns the same eigenpairs as the per-constraint calls."""
    from scip_sdp_b200 import misdp
    (ours, bo), (ref, br) = libs
    M = misdp.read_instance(os.path.join(ROOT, "tests", "golden", name))
    rng = np.random.default_rng(17)
    tol = 1e-6
    ours.SCIPlapackComputeEigenvectorsNegativeBatch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(_dp), C.c_double,
                                                                 C.POINTER(C.c_int), C.POINTER(_dp), C.POINTER(_dp)]
    for trial in range(3):
        lo = np.where(M.lb > -1e19, M.lb, -1.0); hi = np.where(M.ub < 1e19, M.ub, 1.0)
        y = lo + (hi - lo) * rng.random(M.nvars) * (0.3 if trial else 0.0)          # an LP-relaxation-like point: Z(y) is indefinite
        blocks = _dense_blocks(M, y)
        # batch call over all blocks of the "separation round"
        nb = len(blocks)
        sizes = (C.c_int * nb)(*[Z.shape[0] for Z, _, _ in blocks])
        mats = [Z.copy() for Z, _, _ in blocks]
        ws = [np.zeros(Z.shape[0]) for Z, _, _ in blocks]
        Vs = [np.zeros(Z.shape[0] ** 2) for Z, _, _ in blocks]
        cnts = (C.c_int * nb)()
        assert ours.SCIPlapackComputeEigenvectorsNegativeBatch(bo, nb, sizes, (_dp * nb)(*[_p(m) for m in mats]), tol, cnts,
                                                                (_dp * nb)(*[_p(w) for w in ws]), (_dp * nb)(*[_p(v) for v in Vs])) == 1
        for k, (Z, Aj, A0) in enumerate(blocks):
            nrm = max(1.0, np.abs(Z).max())
            wo, co, lo_ = _cuts(ours, bo, Z, Aj, A0, tol)
            wr, cr, lr = _cuts(ref, br, Z, Aj, A0, tol)
            wb, cb, lb_ = _cuts(ours, bo, Z, Aj, A0, tol, batch=(cnts[k], ws[k], Vs[k]))
            assert len(wo) == len(wr) == len(wb)
            if len(wo) == 0:
                continue
            assert np.abs(wo - wr).max() <= 1e-10 * nrm and np.abs(wb - wo).max() <= 1e-12 * nrm
            # violation of cut i at y = coef_i . y - lhs_i = v' Z(y) v = eigenvalue i
            assert np.abs(co @ y - lo_ - wo).max() <= 1e-9 * nrm
            # clusters of (numerically) equal eigenvalues: the sum of their cuts does not depend on the basis chosen inside the cluster
            edges = [0] + [i for i in range(1, len(wr)) if wr[i] - wr[i - 1] > 1e-7 * nrm] + [len(wr)]
            for a, e in zip(edges[:-1], edges[1:]):
                scale = max(1.0, np.abs(cr[a:e].sum(0)).max())
                assert np.abs(co[a:e].sum(0) - cr[a:e].sum(0)).max() <= 1e-8 * scale
                assert np.abs(cb[a:e].sum(0) - cr[a:e].sum(0)).max() <= 1e-8 * scale
                assert abs(lo_[a:e].sum() - lr[a:e].sum()) <= 1e-8 * max(1.0, abs(lr[a:e].sum()))
