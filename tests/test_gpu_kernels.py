"""GPU suite, part 1: each hand-written kernel family through the C ABI against the CPU oracle (LAPACK/BLAS) on the
same seeded inputs.  FP64 tolerances are stated per test."""
import numpy as np
import pytest

from scip_sdp_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    return abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)


@pytest.fixture(scope="module")
def cpu():
    return abi.Solver(abi.Lib(abi.ORACLE_LIB))


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (64, 64, 64), (65, 63, 17), (130, 257, 100), (200, 100, 333), (512, 384, 640)])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_dgemm_matches_blas(gpu, cpu, m, n, k, ta, tb):
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    C0 = rng.standard_normal((m, n))
    got = gpu.dgemm(A, B, ta, tb, alpha=1.25, beta=-0.5, Cin=C0)
    ref = cpu.dgemm(A, B, ta, tb, alpha=1.25, beta=-0.5, Cin=C0)
    # tolerance: k * eps * |A||B| (different summation order on the tensor cores)
    assert np.abs(got - ref).max() <= 4e-16 * k * (np.abs(A).max() * np.abs(B).max()) * 8 + 1e-13


@pytest.mark.parametrize("m,n,k", [(1100, 900, 333), (2000, 1000, 517), (1025, 1027, 2)])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("variant", ["tma", "cpasync"])
def test_large_dgemm_tma_and_cp_async_variants(gpu, cpu, m, n, k, ta, tb, variant, monkeypatch):
    """products with at least 148 tiles of 64 x 64 take the large-tile kernel: fed by TMA tensor maps (default) or by the cp.async
    ring of round 1 (SDPCUDA_GEMM=cpasync); ragged edges are zero-filled by the TMA unit / by zero-size cp.async"""
    monkeypatch.setenv("SDPCUDA_GEMM", variant)
    rng = np.random.default_rng(m + 3 * n + 7 * k)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    C0 = rng.standard_normal((m, n))
    got = gpu.dgemm(A, B, ta, tb, alpha=-0.75, beta=0.5, Cin=C0)
    ref = cpu.dgemm(A, B, ta, tb, alpha=-0.75, beta=0.5, Cin=C0)
    assert np.abs(got - ref).max() <= 4e-16 * k * (np.abs(A).max() * np.abs(B).max()) * 8 + 1e-13


def test_dgemm_empty_and_degenerate(gpu):
    out = gpu.dgemm(np.zeros((3, 0)), np.zeros((0, 4)), Cin=np.ones((3, 4)), beta=2.0)
    assert np.allclose(out, 2.0)


@pytest.mark.parametrize("n", [1, 2, 17, 64, 65, 128, 200, 333, 1000])
def test_potrf_and_trtri_match_lapack(gpu, cpu, n):
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n))
    A = G @ G.T + n * np.eye(n)
    L, info = gpu.dpotrf(A)
    Lr, infor = cpu.dpotrf(A)
    assert info == 0 and infor == 0
    # relative 1e-12 as in SURVEY.md section 7 step 4
    assert np.abs(L - Lr).max() <= 1e-12 * np.abs(Lr).max()
    assert np.abs(L @ L.T - A).max() <= 1e-12 * np.abs(A).max()
    Li = gpu.dtrtri(Lr)
    Lir = cpu.dtrtri(Lr)
    assert np.abs(Li - Lir).max() <= 1e-11 * np.abs(Lir).max()
    assert np.abs(Li @ Lr - np.eye(n)).max() <= 1e-11


@pytest.mark.parametrize("n", [1, 3, 8, 63, 64, 65, 100, 127, 128, 129, 192, 250, 257, 700, 1501])
def test_potrf_with_inverse_factor(gpu, cpu, n):
    rng = np.random.default_rng(7000 + n)
    G = rng.standard_normal((n, n))
    A = G @ G.T + n * np.eye(n)
    L, Li, info = gpu.dpotrf_inv(A)
    Lr, Lir, infor = cpu.dpotrf_inv(A)
    assert info == 0 and infor == 0
    assert np.abs(L - Lr).max() <= 1e-12 * np.abs(Lr).max()
    assert np.abs(Li - Lir).max() <= 1e-11 * np.abs(Lir).max()
    assert np.abs(np.triu(Li, 1)).max() == 0.0
    assert np.abs(Li @ L - np.eye(n)).max() <= 1e-11


@pytest.mark.parametrize("n", [129, 192, 1000, 2000, 2048, 2049, 3000])
@pytest.mark.parametrize("variant", ["dag", "rec"])
def test_tile_dag_cholesky_matches_lapack_and_the_recursive_chain(gpu, cpu, n, variant, monkeypatch):
    """the one-kernel tile-DAG factorisation (default for n > 128) and the recursive kernel chain of round 1 (SDPCUDA_CHOL=rec)
    against LAPACK, with and without the inverse factor"""
    monkeypatch.setenv("SDPCUDA_CHOL", variant)
    rng = np.random.default_rng(31000 + n)
    G = rng.standard_normal((n, n))
    A = G @ G.T + n * np.eye(n)
    Lr, Lir, _ = cpu.dpotrf_inv(A)
    L, info = gpu.dpotrf(A)
    assert info == 0 and np.abs(L - Lr).max() <= 1e-12 * np.abs(Lr).max()
    L, Li, info = gpu.dpotrf_inv(A)
    assert info == 0 and np.abs(L - Lr).max() <= 1e-12 * np.abs(Lr).max()
    assert np.abs(Li - Lir).max() <= 1e-11 * np.abs(Lir).max() and np.abs(np.triu(Li, 1)).max() == 0.0


def test_tile_dag_cholesky_reports_the_first_bad_pivot(gpu):
    A = np.eye(700); A[450, 450] = -1.0; A[600, 600] = -2.0
    _, info = gpu.dpotrf(A)
    assert info == 451


def test_potrf_with_inverse_badly_scaled(gpu, cpu):
    # diagonal scaling over 12 orders of magnitude: pivots must stay accurate relative to their own size
    rng = np.random.default_rng(99)
    n = 128
    G = rng.standard_normal((n, n))
    D = np.diag(10.0 ** rng.uniform(-6, 6, n))
    A = D @ (G @ G.T + n * np.eye(n)) @ D
    L, Li, info = gpu.dpotrf_inv(A)
    Lr, Lir, _ = cpu.dpotrf_inv(A)
    assert info == 0
    assert (np.abs(L - Lr) <= 1e-11 * np.abs(np.diag(Lr))[:, None] + 1e-300).all()
    assert np.abs(Li @ L - np.eye(n)).max() <= 1e-9


@pytest.mark.parametrize("n,bad", [(64, 0), (100, 70), (128, 127), (300, 129), (300, 64)])
def test_potrf_pivot_index(gpu, n, bad):
    A = np.eye(n); A[bad, bad] = -1.0
    _, info = gpu.dpotrf(A)
    assert info == bad + 1
    _, _, info = gpu.dpotrf_inv(A)
    assert info == bad + 1


def test_potrf_reports_indefinite_matrix(gpu):
    A = np.eye(100); A[70, 70] = -1.0
    _, info = gpu.dpotrf(A)
    assert info == 71
    assert gpu.psd_check(A) is False and gpu.psd_check(A, shift=1.5) is True


@pytest.mark.parametrize("n", [1, 2, 3, 10, 15, 43, 64, 96, 120])
def test_jacobi_eigen_matches_lapack(gpu, cpu, n):
    rng = np.random.default_rng(100 + n)
    nb = 5
    A = rng.standard_normal((nb, n, n)); A = A + A.transpose(0, 2, 1)
    if n >= 4:
        A[1] = np.diag(np.arange(n, dtype=float))            # already diagonal
        A[2][:, :] = 1.0                                      # rank one: (n-1)-fold eigenvalue 0
    w, V = gpu.syev(A)
    wr, _ = cpu.syev(A)
    for b in range(nb):
        nrm = np.abs(wr[b]).max() + 1e-300
        assert np.abs(w[b] - wr[b]).max() <= 1e-10 * nrm                       # eigenvalues to 1e-10 |A|
        assert np.all(np.diff(w[b]) >= -1e-12 * nrm)                           # ascending
        # eigenvectors as ROWS (lapack_interface.c:507-603): residual and orthonormality instead of sign/rotation matching
        R = A[b] @ V[b].T - V[b].T * w[b][None, :]
        assert np.abs(R).max() <= 1e-10 * nrm * n
        assert np.abs(V[b] @ V[b].T - np.eye(n)).max() <= 1e-10 * n


def test_reference_lapack_probe_convention(gpu):
    """[[1,2],[2,4]] -> eigenvalues (0,5), eigenvectors as rows (-0.894,0.447), (0.447,0.894)"""
    w, V = gpu.syev(np.array([[1.0, 2.0], [2.0, 4.0]]))
    assert np.allclose(w, [0.0, 5.0], atol=1e-12)
    assert np.allclose(np.abs(V), [[0.894427191, 0.4472135955], [0.4472135955, 0.894427191]], atol=1e-9)
