"""CPU suite: instance readers (the data formats in front of the hot path).
CBF semantics follow src/scipsdp/reader_cbf.c (file:line in scip_sdp_b200/misdp.py:read_cbf); SDPA semantics reader_sdpa.c."""
import os

import numpy as np
import pytest

from scip_sdp_b200 import abi, misdp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

CBF_PSDVAR = """
# min <C, X> + 3 x0  with C = [[1, 0.5], [0.5, 2]]  s.t.  X_01 = 1 (as <F, X> - 1 = 0, F_10 = 0.5),  x0 - 2 >= 0,  X psd
VER
2

OBJSENSE
MAX

VAR
1 1
L+ 1

PSDVAR
1
2

CON
2 2
L= 1
L+ 1

OBJFCOORD
3
0 0 0 -1.0
0 1 0 -0.5
0 1 1 -2.0

OBJACOORD
1
0 -3.0

OBJBCOORD
10.0

FCOORD
1
0 0 1 0 0.5

ACOORD
1
1 0 1.0

BCOORD
2
0 -1.0
1 -2.0
"""


def test_cbf_matrix_variable_and_max_sense(tmp_path):
    p = tmp_path / "tiny.cbf"
    p.write_text(CBF_PSDVAR)
    M = misdp.read_cbf(p)
    # one scalar variable + the lower triangle (0,0), (1,0), (1,1) of X
    assert M.nvars == 4 and M.blocksizes == [2] and M.objsense == -1 and M.objoffset == 10.0
    assert M.lb[0] == 0.0 and M.ub[0] >= 1e20 and (M.lb[1:] <= -1e20).all()
    # max of the negated file objective = min of: 3 x0 + X00 + 2*0.5 X10 + 2 X11 ; off-diagonals count twice
    assert np.allclose(M.obj, [3.0, 1.0, 1.0, 2.0])
    (c0, lhs0, rhs0), (c1, lhs1, rhs1) = M.rows
    assert c0 == {2: 1.0} and lhs0 == rhs0 == 1.0            # 2 * 0.5 * X10 = 1
    assert c1 == {0: 1.0} and lhs1 == 2.0 and rhs1 >= 1e20
    assert M.A[0] == {1: [(0, 0, 1.0)], 2: [(1, 0, 1.0)], 3: [(1, 1, 1.0)]} and M.C[0] == []
    # optimum: x0 = 2, X = [[a, 1], [1, 1/a]] minimising a + 2/a -> a = sqrt 2: value 6 + 1 + 2 sqrt 2; file sense: 10 - that
    if os.path.exists(abi.ORACLE_LIB):
        fp, info = M.rows_to_bounds().flatten()
        r = abi.Solver(abi.Lib(abi.ORACLE_LIB)).solve(fp, gaptol=1e-8, feastol=1e-8)
        assert r["phase_name"] == "pdOPT"
        val = M.file_objective(r["dobj"] + info["fixedobj"])
        assert abs(val - (10.0 - (7.0 + 2.0 * np.sqrt(2.0)))) <= 1e-5


def test_cbf_lmi_constant_sign_and_integrality():
    M = misdp.read_cbf(os.path.join(GOLDEN, "example_small_cbf.cbf"))
    S = misdp.read_sdpa(os.path.join(GOLDEN, "example_small.dat-s"))
    # the CBF twin of example_small: same LMIs (D = -A_0), same integrality, same objective
    assert M.blocksizes == S.blocksizes and (M.integer == S.integer).all() and np.allclose(M.obj, S.obj)
    y = np.array([0.3, -1.2, 2.5])
    for Zc, Zs in zip(M.dense_Z(y), S.dense_Z(y)):
        assert np.allclose(Zc, Zs)


def test_cbf_rejects_unsupported_cones(tmp_path):
    p = tmp_path / "soc.cbf"
    p.write_text("VER\n1\nOBJSENSE\nMIN\nVAR\n3 1\nQ 3\n")
    with pytest.raises(ValueError):
        misdp.read_cbf(p)


def test_read_instance_dispatch():
    assert misdp.read_instance(os.path.join(GOLDEN, "example_cbf_dual.cbf")).blocksizes == [2, 2]
    assert misdp.read_instance(os.path.join(GOLDEN, "example_TT.dat-s.gz")).nvars == 37


@pytest.mark.parametrize("make", [lambda: misdp.read_sdpa(os.path.join(GOLDEN, "example_TT.dat-s.gz")),
                                  lambda: __import__("scip_sdp_b200.generators", fromlist=["x"]).mkp(12, seed=3),
                                  lambda: __import__("scip_sdp_b200.generators", fromlist=["x"]).truss(3, 3, 20, seed=4),
                                  lambda: __import__("scip_sdp_b200.generators", fromlist=["x"]).maxcut(30, 0.2, seed=5)],
                         ids=["example_TT", "mkp-12", "truss-20", "maxcut-30"])
def test_sdpa_writer_reader_round_trip(tmp_path, make):
    """the synthetic generators write the reference's extended SDPA format (`.dat-s` files the reference readers
    could consume): writing and reading back gives the same model (LMIs, objective, integrality) and the same flattened problem"""
    M = make()
    path = tmp_path / "inst.dat-s.gz"
    M.write_sdpa(path)
    R = misdp.read_sdpa(path)
    assert R.nvars == M.nvars and R.blocksizes == M.blocksizes and np.allclose(R.obj, M.obj) and (R.integer == M.integer).all()
    rng = np.random.default_rng(1)
    y = rng.standard_normal(M.nvars)
    for Za, Zb in zip(M.dense_Z(y), R.dense_Z(y)):
        assert np.allclose(Za, Zb)
    fa, _ = M.rows_to_bounds().flatten()
    fb, _ = R.rows_to_bounds().flatten()
    assert fa.m == fb.m and fa.nlp == fb.nlp and list(fa.blocksizes) == list(fb.blocksizes)
    # the LP blocks describe the same polyhedron row by row (the writer emits one-sided rows, bounds last)
    assert np.allclose(np.sort(fa.lprhs), np.sort(fb.lprhs))


def test_generators_are_deterministic():
    from scip_sdp_b200 import generators
    for make in (lambda: generators.maxcut(40, 0.2, seed=9), lambda: generators.mkp(10, seed=9), lambda: generators.truss(3, 3, 15, seed=9),
                 lambda: generators.cls(12, 8, 3, seed=9)):
        a, _ = make().flatten()
        b, _ = make().flatten()
        assert np.array_equal(a.entval, b.entval) and np.array_equal(a.obj, b.obj) and np.array_equal(a.lprhs, b.lprhs)
    fa, fb = generators.dense_sdp_flat(5, 4, seed=1), generators.dense_sdp_flat(5, 4, seed=1)
    assert np.array_equal(fa.entval, fb.entval) and np.array_equal(fa.cval, fb.cval)


def test_sdpa_indicator_entries():
    """extended SDPA format: a negative variable index -k (k >= 2) in the LP block marks an indicator constraint
    (reader_sdpa.c:1147-1246): variable k-1 becomes binary and the row gets a slack variable with "variable = 1 => slack = 0" """
    M = misdp.read_sdpa(os.path.join(GOLDEN, "example_small_ind.dat-s"))
    assert M.nvars == 5 and M.indicators == [(4, 3)]
    assert M.integer.tolist() == [True, True, True, True, False]
    assert (M.lb[3], M.ub[3], M.lb[4]) == (0.0, 1.0, 0.0) and M.ub[4] >= 1e20
    rows_with_slack = [coefs for coefs, lhs, rhs in M.rows if 4 in coefs]
    assert len(rows_with_slack) == 1 and rows_with_slack[0][4] == 1.0


# the defect classes of unittests/src/readerrors.c (28 malformed files in unittests/instances, each must give SCIP_READERROR),
# re-created as edits of a small valid file: 3 variables, blocks (2, LP with 2 rows, 2), all integer, block 3 rank-1
_VALID = """3
3
2 -2 2
1 1 1
1 1 1 1 1
1 3 2 1 1
1 2 1 1 1
2 2 1 1 1
3 2 1 1 1
0 2 1 1 1
1 2 2 2 -1
2 2 2 2 -1
3 2 2 2 -1
0 2 2 2 -1
*INTEGER
*1
*2
*3
*RANK1
*3
"""


def _edit(old, new):
    assert old in _VALID
    return _VALID.replace(old, new, 1)


_DEFECTS = {
    "nvars_invalidsymb": _edit("3\n3\n", "?\n3\n"), "nvars_neg": _edit("3\n3\n", "-3\n3\n"),
    "nblocks_invalidsymb": _edit("3\n3\n", "3\n?\n"), "nblocks_neg": _edit("3\n3\n", "3\n-3\n"),
    "blocksizes_0": _edit("2 -2 2", "2 -2 0"), "blocksizes_invalidsymb": _edit("2 -2 2", "2 -2 ?"),
    "blocksizes_LPblocks": _edit("2 -2 2", "2 -2 -2"), "blocksizes_toofew": _edit("2 -2 2", "2 -2"),
    "objcoeff_invalidsymb": _edit("1 1 1\n1 1 1 1 1", "1 ? 1\n1 1 1 1 1"), "objcoeff_toofew": _edit("1 1 1\n1 1 1 1 1", "1 1\n1 1 1 1 1"),
    "blocks_col": _edit("1 1 1 1 1", "1 1 1 3 1"), "blocks_row": _edit("1 1 1 1 1", "1 1 -1 1 1"), "blocks_var": _edit("1 1 1 1 1", "4 1 1 1 1"),
    "blocks_sdpblock": _edit("1 1 1 1 1", "1 4 1 1 1"), "blocks_invalidline": _edit("1 1 1 1 1", "1 1 1 1"),
    "blocks_SDPnononz": _edit("1 1 1 1 1\n", ""), "blocks_LPnononz": _edit("1 2 1 1 1\n2 2 1 1 1\n3 2 1 1 1\n0 2 1 1 1\n", ""),
    "LPblock_LPcons": _edit("1 2 1 1 1", "1 2 3 3 1"), "LPblock_nondiag": _edit("1 2 1 1 1", "1 2 1 0 1"), "LPblock_var": _edit("1 2 1 1 1", "4 2 1 1 1"),
    "int_invalid": _edit("*INTEGER\n*1", "*INTEGER\n*"), "int_noast": _edit("*INTEGER\n*1", "*INTEGER\n1"), "int_var": _edit("*2\n*3\n*RANK1", "*2\n*4\n*RANK1"),
    "rnk1_before_int": "\n".join(_VALID.splitlines()[:14] + ["*RANK1", "*1", "*INTEGER", "*1", "*2", "*3"]) + "\n",
    "rnk1_block": _edit("*RANK1\n*3", "*RANK1\n*4"), "rnk1_forLP": _edit("*RANK1\n*3", "*RANK1\n*2"),
    "rnk1_invalid": _edit("*RANK1\n*3", "*RANK1\n*"), "rnk1_noast": _edit("*RANK1\n*3", "*RANK1\n3"),
}


def test_valid_base_file_of_the_defect_tests(tmp_path):
    p = tmp_path / "valid.dat-s"
    p.write_text(_VALID)
    M = misdp.read_sdpa(p)
    assert M.nvars == 3 and M.blocksizes == [2, 2] and len(M.rows) == 2 and M.integer.all() and M.rank1 == [1]


@pytest.mark.parametrize("defect", sorted(_DEFECTS))
def test_malformed_sdpa_files_are_rejected(tmp_path, defect):
    assert len(_DEFECTS) == 28
    p = tmp_path / (defect + ".dat-s")
    p.write_text(_DEFECTS[defect])
    with pytest.raises(misdp.SdpaFormatError):
        misdp.read_sdpa(p)


@pytest.mark.skipif(not os.path.isdir("/root/reference/unittests/instances"), reason="reference tree not present")
def test_reference_malformed_files_are_rejected():
    """the reference's own 28 malformed files (unittests/src/readerrors.c) where the reference tree is available"""
    import glob
    files = [f for f in sorted(glob.glob("/root/reference/unittests/instances/*.dat-s")) if "example_small" not in f]
    assert len(files) == 28
    for f in files:
        with pytest.raises(misdp.SdpaFormatError):
            misdp.read_sdpa(f)


def test_vectorised_node_marshalling_equals_the_loop_version():
    """Misdp.flatten_fast / node_problem_fast (numpy, used by the frontier drivers) produce exactly the arrays of flatten /
    node_problem (plain loops, the documented restatement of sdpisolver_sdpa.cpp:1015-1412 and sdpi.c's node presolve)"""
    import glob
    from scip_sdp_b200 import generators
    rng = np.random.default_rng(0)
    models = [misdp.read_instance(f) for f in sorted(glob.glob(os.path.join(GOLDEN, "example_*")))]
    models += [generators.truss(4, 4, 60, seed=13), generators.cls(20, 12, 4, seed=3), generators.mkp(12, seed=2), generators.maxcut(30, 0.2, seed=1)]
    fields = "obj blocksizes varbeg entblk entrow entcol entval cblk crow ccol cval lpbeg lpind lpval lprhs".split()
    for M in models:
        for to_bounds in (False, True):
            if to_bounds:
                M = M.rows_to_bounds()
            ints = np.flatnonzero(M.integer)
            for trial in range(6):
                lb, ub = M.lb.copy(), M.ub.copy()
                for j in rng.permutation(ints)[:rng.integers(0, len(ints) + 1)]:
                    lb[j] = ub[j] = float(np.clip(rng.integers(0, 2), max(lb[j], -5), min(ub[j], 5)))
                for compress, skip in ((False, False), (True, True)):
                    a, ia = M.flatten(lb, ub, compress=compress, skip_single_rows=skip)
                    b, ib = M.flatten_fast(lb, ub, compress=compress, skip_single_rows=skip)
                    assert all(np.array_equal(getattr(a, k), getattr(b, k)) for k in fields)
                    assert ia["fixedobj"] == ib["fixedobj"] and ia["rowmap"] == ib["rowmap"] and ia["boundmap"] == ib["boundmap"]
                s1, s2 = M.node_problem(lb, ub), M.node_problem_fast(lb, ub)
                assert s1[0] == s2[0]
                if s1[0] == "solve":
                    assert all(np.array_equal(getattr(s1[1], k), getattr(s2[1], k)) for k in fields)
                    assert np.array_equal(s1[2]["lb"], s2[2]["lb"]) and np.array_equal(s1[2]["ub"], s2[2]["ub"])


# unittests/src/readwrite.c:66-140 (runTests): read an instance, write it as CBF and as SDPA, read both files again, solve all three:
# the optimal values agree (the SDPA format has no objective sense, so its value is compared in the internal min form)
READWRITE = ["example_small.dat-s", "example_small_cbf.cbf", "example_inf.dat-s", "example_TT.dat-s.gz", "example_MkP.dat-s.gz",
             "example_cbf_primal.cbf", "example_cbf_mix.cbf", "example_cbf_dual.cbf", "example_multaggr.cbf", "example_diagzeroimpl.cbf",
             "example_tightenmatrices.dat-s"]


@pytest.mark.parametrize("name", READWRITE)
def test_read_write_read_gives_the_same_optimum(tmp_path, name):
    from scip_sdp_b200 import frontier
    lib = abi.Lib(abi.ORACLE_LIB)
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    M.write_cbf(tmp_path / "t.cbf")
    M.write_sdpa(tmp_path / "t.dat-s")
    Mc, Ms = misdp.read_cbf(tmp_path / "t.cbf"), misdp.read_sdpa(tmp_path / "t.dat-s")
    assert (Mc.nvars, Mc.blocksizes, Mc.objsense) == (M.nvars, M.blocksizes, M.objsense) and np.array_equal(Mc.integer, M.integer)
    assert np.allclose(Mc.obj, M.obj) and Mc.objoffset == M.objoffset
    runs = [frontier.branch_and_bound(abi.Solver(lib), X, mode="batch", width=64, timelimit=120) for X in (M, Mc, Ms)]
    assert len({r["status"] for r in runs}) == 1
    if runs[0]["status"] == "optimal":
        v = [runs[0]["objval"], runs[1]["objval"], runs[2]["objval"]]
        assert abs(M.file_objective(v[0]) - Mc.file_objective(v[1])) <= 1e-4 * max(1.0, abs(v[0]))
        assert abs(v[0] - v[2]) <= 1e-4 * max(1.0, abs(v[0]))


def test_cbf_mixed_and_dual_form_have_the_same_optimum():
    """unittests/src/mixcbf.c:70-93 (readCBFmixreadCBFdual): the same problem in mixed form (matrix variable + LMI) and in dual form;
    plus the primal form of the same family (example_cbf_primal has its own optimum 0.75 in short.solu)"""
    from scip_sdp_b200 import frontier
    lib = abi.Lib(abi.ORACLE_LIB)
    vals = []
    for f in ("example_cbf_mix.cbf", "example_cbf_dual.cbf"):
        M = misdp.read_instance(os.path.join(GOLDEN, f))
        r = frontier.branch_and_bound(abi.Solver(lib), M, mode="batch", width=16)
        assert r["status"] == "optimal"
        vals.append(M.file_objective(r["objval"]))
    assert abs(vals[0] - vals[1]) <= 1e-6 and abs(vals[0] - 4.0) <= 1e-4
