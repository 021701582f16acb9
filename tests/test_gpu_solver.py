"""GPU suite, part 2: the interior-point solver (sdpcuda_solve) against the CPU oracle on the BASELINE instance shapes,
plus a-posteriori KKT residuals at sizes the oracle does not finish quickly.  Tolerance: relaxation objective within the
solver gap tolerance 1e-5 relative (north_star), y within 1e-4 of the oracle's where the solution is unique."""
import os

import numpy as np
import pytest

from scip_sdp_b200 import abi, generators, misdp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gpu():
    return abi.Solver(abi.Lib(abi.PRODUCT_LIB), device=0)


@pytest.fixture(scope="module")
def cpu():
    return abi.Solver(abi.Lib(abi.ORACLE_LIB))


def _instances():
    yield "example_small", lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s")).rows_to_bounds()
    yield "example_TT", lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")).rows_to_bounds()
    yield "example_CLS", lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_CLS.dat-s.gz")).rows_to_bounds()
    yield "example_MkP", lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_MkP.dat-s.gz")).rows_to_bounds()
    yield "maxcut-150", lambda: generators.maxcut(150, 0.05, seed=11)
    yield "mkp-24", lambda: generators.mkp(24, seed=12)
    yield "truss-60", lambda: generators.truss(4, 4, 60, seed=13)
    yield "cls-40", lambda: generators.cls(40, 25, 5, seed=14)
    yield "cls-39", lambda: generators.cls(39, 22, 5, seed=15)          # odd number of dense constraint matrices


@pytest.mark.parametrize("leaf,panel,nodes", [("128", "64", 24), ("64", "128", 40), ("128", "512", 24)])
def test_lookahead_factorisation_and_panel_substitution(gpu, cpu, leaf, panel, nodes, monkeypatch):
    """the path of large Schur complements (no explicit inverse factor): right-looking panels with look-ahead on a side stream and
    M dy = g by panel substitution, forced on small instances with small panels; also covers both leaf orders of the recursion"""
    fp, _ = generators.mkp(nodes, seed=11).flatten()
    monkeypatch.setenv("SDPCUDA_PATH", "m")
    monkeypatch.setenv("SDPCUDA_MINV_MAX", "0")
    monkeypatch.setenv("SDPCUDA_LEAF", leaf)
    monkeypatch.setenv("SDPCUDA_PANEL", panel)
    r = gpu.solve(fp, gaptol=1e-7, feastol=1e-7)
    ref = cpu.solve(fp, gaptol=1e-7, feastol=1e-7)
    assert ref["phase_name"] == "pdOPT" and r["phase_name"] == "pdOPT", (r["phase_name"], r["stop_name"])
    assert abs(r["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
    _check_kkt(fp, r, 1e-6)


@pytest.mark.parametrize("path", ["single-cta", "multi-kernel"])
@pytest.mark.parametrize("name,make", list(_instances()), ids=[n for n, _ in _instances()])
def test_relaxation_matches_oracle(gpu, cpu, name, make, path, monkeypatch):
    """both device paths: the one-launch kernel for small relaxations (ipm_small.cu) and the kernel-per-operation pipeline"""
    fp, _ = make().flatten()
    monkeypatch.setenv("SDPCUDA_PATH", "s" if path == "single-cta" else "m")
    if path == "single-cta" and (max(fp.blocksizes, default=0) > 64 or fp.m > 256):
        pytest.skip("outside the single-CTA limits")
    r = gpu.solve(fp, gaptol=1e-7, feastol=1e-7)
    if path == "single-cta":
        assert r["launches"] < 40                      # initial point + ONE solver kernel
    ref = cpu.solve(fp, gaptol=1e-7, feastol=1e-7)
    assert r["launches"] > 0
    assert ref["phase_name"] == "pdOPT" and r["phase_name"] == "pdOPT", (r["phase_name"], r["stop_name"])
    assert abs(r["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
    assert abs(r["pobj"] - ref["pobj"]) <= 1e-5 * max(1.0, abs(ref["pobj"]))
    _check_kkt(fp, r, 1e-6)


def _check_kkt(fp, r, tol):
    y = r["y"]
    C = fp.dense_C()
    for k in range(fp.nblocks):
        Z = -C[k].copy()
        for j in range(fp.m):
            Aj = fp.dense_A(j)[k]
            if np.any(Aj):
                Z += y[j] * Aj
        assert np.linalg.norm(Z - r["S"][k]) <= tol * (1 + np.linalg.norm(C[k]))
        assert np.linalg.eigvalsh(r["X"][k]).min() >= -1e-9 and np.linalg.eigvalsh(r["S"][k]).min() >= -1e-9
    if fp.nlp:
        D = fp.dense_D()
        assert np.abs(D @ y - fp.lprhs - r["slp"]).max() <= tol * (1 + np.abs(fp.lprhs).max())
        assert r["xlp"].min() >= 0 and r["slp"].min() >= 0


def test_infeasible_and_unbounded_certificates(gpu):
    # y-problem infeasible: [[y1, 1], [1, 0.75 y2]] psd with |y| <= 1 (checksdpi.c test9)
    M = misdp.Misdp(2, [-1.0, 0.0], [2])
    M.A[0][0] = [(0, 0, 1.0)]; M.A[0][1] = [(1, 1, 0.75)]; M.C[0] = [(1, 0, -1.0)]
    M.lb[:] = -1.0; M.ub[:] = 1.0
    r = gpu.solve(M.flatten()[0], gaptol=1e-6, feastol=1e-6)
    assert r["phase_name"] in ("pFEAS_dINF", "dINF")
    # y-problem unbounded: min -3 y1 - y2 s.t. 2 y1 + y2 <= 10, y1 + 3 y2 <= 15 (checksdpi.c test2)
    M = misdp.Misdp(2, [-3.0, -1.0], [])
    M.add_row({0: 2.0, 1: 1.0}, rhs=10.0); M.add_row({0: 1.0, 1: 3.0}, rhs=15.0)
    r = gpu.solve(M.flatten()[0], gaptol=1e-6, feastol=1e-6)
    assert r["phase_name"] == "pINF_DFEAS".replace("DFEAS", "dFEAS")


def test_maxcut_600_kkt_and_properties(gpu):
    """size-independent checks at a size where the Lanczos step-length path and the recursive Cholesky are active"""
    fp, _ = generators.maxcut(600, 0.02, seed=5).flatten()
    r = gpu.solve(fp, gaptol=1e-6, feastol=1e-6)
    assert r["phase_name"] == "pdOPT"
    y, X, S = r["y"], r["X"][0], r["S"][0]
    C = fp.dense_C()[0]
    assert np.linalg.norm(np.diag(y) - C - S) <= 1e-6 * (1 + np.linalg.norm(C))
    assert np.abs(np.diag(X) - 1.0).max() <= 1e-6                  # A_i . X = b_i = 1
    assert np.linalg.eigvalsh(X).min() >= -1e-9 and np.linalg.eigvalsh(S).min() >= -1e-9
    assert abs(np.vdot(X, S)) <= 1e-5 * max(1.0, abs(r["dobj"]))
    assert abs(y.sum() - np.vdot(C, X)) <= 1e-5 * abs(y.sum())


@pytest.mark.parametrize("name,make", list(_instances()), ids=[n for n, _ in _instances()])
def test_schur_shares_add_up_to_the_unsharded_complement(gpu, name, make, monkeypatch):
    """the partition of the Schur complement used by the multi-GPU path, emulated on one GPU: forming the shares of
    three ranks one after the other must give the same iterates (every entry belongs to exactly one share): bit-identical"""
    fp, _ = make().flatten()
    monkeypatch.setenv("SDPCUDA_PATH", "m")
    a = gpu.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)
    monkeypatch.setenv("SDPCUDA_SHARD_EMULATE", "3")
    b = gpu.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)
    assert a["phase_name"] == b["phase_name"] == "pdOPT"
    # every entry of M has exactly one writer on exactly one rank, also in the LP block (single-writer kernel, fixed summation order)
    assert a["iterations"] == b["iterations"] and a["dobj"] == b["dobj"] and a["pobj"] == b["pobj"]
    c = gpu.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)          # and a second run reproduces the first bit by bit
    assert c["iterations"] == b["iterations"] and c["dobj"] == b["dobj"] and c["pobj"] == b["pobj"]


def test_rank_one_schur_path_matches_the_entry_path(gpu, cpu, monkeypatch):
    """truss topology with 300 bars: the element matrices are rank one, their mutual Schur entries come from the two GEMM pairs
    (A' X A) o (A' S^-1 A) (SURVEY 7.5 ii); with SDPCUDA_RANK1=0 every pair goes through the entry lists.  Same optimum to 1e-9
    relative, the same iteration count, and both within 1e-5 of the oracle; the shares of three emulated ranks reproduce the
    unsharded iterates bit by bit with the path on"""
    fp, _ = generators.truss(5, 5, 300, seed=21).flatten()
    monkeypatch.setenv("SDPCUDA_PATH", "m")
    monkeypatch.setenv("SDPCUDA_RANK1", "force")
    on = gpu.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)
    monkeypatch.setenv("SDPCUDA_SHARD_EMULATE", "3")
    on3 = gpu.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)
    monkeypatch.delenv("SDPCUDA_SHARD_EMULATE")
    monkeypatch.setenv("SDPCUDA_RANK1", "0")
    off = gpu.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)
    ref = cpu.solve(fp, gaptol=1e-7, feastol=1e-7)
    assert on["phase_name"] == off["phase_name"] == ref["phase_name"] == "pdOPT"
    assert on["launches"] != off["launches"]                       # the path was taken (five more launches per iteration)
    assert on["iterations"] == off["iterations"]
    assert abs(on["dobj"] - off["dobj"]) <= 1e-9 * max(1.0, abs(off["dobj"]))
    assert abs(on["dobj"] - ref["dobj"]) <= 1e-5 * max(1.0, abs(ref["dobj"]))
    assert on3["iterations"] == on["iterations"] and on3["dobj"] == on["dobj"] and on3["pobj"] == on["pobj"]


def test_sharded_schur_on_two_gpus():
    """one SDP over two GPUs with the NCCL all-reduce of the Schur shares (needs two devices; tools/dist_check.py)"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(root, "tools", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("name", ["tt500", "cls", "mkp60", "mkp120", "maxcut2000"])
def test_full_size_shapes_match_pinned_oracle_objectives(gpu, name):
    """the synthetic BASELINE.json shapes at FULL size against root-relaxation objectives pinned with the CPU oracle
    (tests/golden/relaxation_values.json, made by tests/golden/make_relaxation_values.py; the oracle needs 2-40 s per shape, too
    slow to repeat in every GPU run).  Tolerance: 1e-5 relative, the north-star gap tolerance both sides were solved to."""
    import json
    from golden.make_relaxation_values import SHAPES
    with open(os.path.join(os.path.dirname(__file__), "golden", "relaxation_values.json")) as f:
        pinned = json.load(f)[name]
    fp, _ = SHAPES[name]().flatten()
    assert (fp.m, [int(b) for b in fp.blocksizes], fp.nlp) == (pinned["m"], pinned["blocks"], pinned["nlp"])
    r = gpu.solve(fp, gaptol=1e-5, feastol=1e-5, fetch=False)
    assert r["phase_name"] == "pdOPT", (r["phase_name"], r["stop_name"])
    scale = max(1.0, abs(pinned["dobj"]))
    assert abs(r["dobj"] - pinned["dobj"]) <= 1e-5 * scale
    # size-independent properties: small relative gap and residuals, X-side value (a lower bound) not above the y-side value
    assert r["relgap"] <= 1e-5 and r["pinf"] <= 1e-5 and r["dinf"] <= 1e-5
    assert r["pobj"] <= r["dobj"] + 2e-5 * scale
    assert abs(r["iterations"] - pinned["iterations"]) <= 5
