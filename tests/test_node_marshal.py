"""CPU suite, part 7: the native node marshalling (csrc/node_marshal.hpp, behind sdpcuda_solve_nodes) against its documented Python
restatement Misdp.node_problem / flatten: the same solver-form arrays for random nodes of every readable instance and of the
synthetic shapes, and the same frontier results through sdpcuda_solve_nodes as through the Python path."""
import glob
import os

import numpy as np
import pytest

from scip_sdp_b200 import abi, generators, misdp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIELDS = "obj blocksizes varbeg entblk entrow entcol entval cblk crow ccol cval lpbeg lpind lpval lprhs".split()


def _models():
    for f in sorted(glob.glob(os.path.join(GOLDEN, "example_*"))):
        yield os.path.basename(f), (lambda f=f: misdp.read_instance(f))
    yield "truss", lambda: generators.truss(4, 4, 60, seed=13)
    yield "cls", lambda: generators.cls(20, 12, 4, seed=3)
    yield "mkp", lambda: generators.mkp(12, seed=2)
    yield "maxcut", lambda: generators.maxcut(30, 0.2, seed=1)


def _random_nodes(M, rng, count):
    ints = np.flatnonzero(M.integer)
    for _ in range(count):
        lb, ub = M.lb.copy(), M.ub.copy()
        for j in rng.permutation(ints)[:rng.integers(0, len(ints) + 1)]:
            lb[j] = ub[j] = float(np.clip(rng.integers(0, 2), max(lb[j], -5), min(ub[j], 5)))
        yield lb, ub


@pytest.mark.parametrize("name,make", list(_models()), ids=[n for n, _ in _models()])
def test_native_node_problem_equals_the_python_restatement(name, make):
    lib = abi.Lib(abi.PRODUCT_LIB)            # pure host code: runs without a GPU
    rng = np.random.default_rng(7)
    for to_bounds in (False, True):
        M = make()
        if to_bounds:
            M = M.rows_to_bounds()
        model = abi.Model(lib, M)
        for lb, ub in _random_nodes(M, rng, 10):
            want = M.node_problem(lb, ub)
            st, fp, info = model.node_problem(lb, ub)
            assert st == {"solve": 0, "infeasible": 1, "allfixed": 2}[want[0]]
            if want[0] == "solve":
                assert all(np.array_equal(getattr(fp, k), getattr(want[1], k)) for k in FIELDS), [k for k in FIELDS if not np.array_equal(getattr(fp, k), getattr(want[1], k))]
                assert np.array_equal(info["active"], want[2]["active"]) and np.array_equal(info["lb"], want[2]["lb"]) and np.array_equal(info["ub"], want[2]["ub"])
                assert abs(info["fixedobj"] - want[2]["fixedobj"]) <= 1e-12 * max(1.0, abs(want[2]["fixedobj"]))
            elif want[0] == "allfixed":
                assert abs(info["fixedobj"] - want[2]["fixedobj"]) <= 1e-12 * max(1.0, abs(want[2]["fixedobj"]))
        model.close()


def test_solve_nodes_equals_the_python_path():
    """sdpcuda_solve_nodes on the checker back end: statuses, bounds, y in model variables and tightened bounds as from
    Misdp.node_problem + solve_batch"""
    lib = abi.Lib(abi.ORACLE_LIB)
    rng = np.random.default_rng(11)
    for f in ("example_TT.dat-s.gz", "example_MkP.dat-s.gz", "example_small.dat-s"):
        M = misdp.read_instance(os.path.join(GOLDEN, f))
        model, s = abi.Model(lib, M), abi.Solver(lib)
        nodes = list(_random_nodes(M, rng, 8))
        out = s.solve_nodes(model, np.array([n[0] for n in nodes]), np.array([n[1] for n in nodes]), gaptol=1e-6, feastol=1e-6)
        for i, (lb, ub) in enumerate(nodes):
            st, fp, info = M.node_problem(lb, ub)
            assert out["status"][i] == {"solve": 0, "infeasible": 1, "allfixed": 2}[st]
            if st == "solve":
                r = s.solve(fp, gaptol=1e-6, feastol=1e-6)
                assert out["results"][i]["phase_name"] == r["phase_name"]
                assert abs(out["bound"][i] - (r["dobj"] + info["fixedobj"])) <= 1e-9 * max(1.0, abs(r["dobj"]))
                y = info["lb"].copy(); y[info["active"]] = r["y"]
                assert np.allclose(out["y"][i], y, rtol=0, atol=1e-12) and np.array_equal(out["lb"][i], info["lb"]) and np.array_equal(out["ub"][i], info["ub"])
            elif st == "allfixed":
                assert abs(out["bound"][i] - info["fixedobj"]) <= 1e-12 * max(1.0, abs(info["fixedobj"]))
        # cutoffs: a node whose bound is above its cutoff stops with pUNBD
        solved = [i for i in range(len(nodes)) if out["status"][i] == 0 and out["results"][i]["phase_name"] == "pdOPT"]
        if solved:
            cut = np.full(len(nodes), 1e20)
            cut[solved[0]] = out["bound"][solved[0]] - 0.05 * max(1.0, abs(out["bound"][solved[0]]))
            again = s.solve_nodes(model, np.array([n[0] for n in nodes]), np.array([n[1] for n in nodes]), cutoff=cut, gaptol=1e-6, feastol=1e-6)
            assert again["results"][solved[0]]["phase_name"] == "pUNBD"
        model.close()


@pytest.mark.parametrize("name,want", [("example_small.dat-s", -8.0), ("example_inf.dat-s", None), ("example_TT.dat-s.gz", 2.11803),
                                       ("example_MkP.dat-s.gz", -95.0), ("example_small_ind.dat-s", -18.0), ("example_cbf_primal.cbf", 0.75)])
def test_branch_and_bound_with_native_node_marshalling(name, want):
    """frontier.branch_and_bound(native=True): the rounds go to sdpcuda_solve_nodes as bound vectors; same tree as the Python path"""
    from scip_sdp_b200 import frontier
    lib = abi.Lib(abi.ORACLE_LIB)
    M = misdp.read_instance(os.path.join(GOLDEN, name))
    a = frontier.branch_and_bound(abi.Solver(lib), M, mode="batch", width=64)
    b = frontier.branch_and_bound(abi.Solver(lib), M, mode="batch", width=64, native=True)
    assert (a["status"], a["nodes"], a["rounds"], a["unsolved"]) == (b["status"], b["nodes"], b["rounds"], b["unsolved"])
    if want is None:
        assert b["status"] == "infeasible"
    else:
        assert b["status"] == "optimal" and abs(M.file_objective(b["objval"]) - want) <= 1e-4 * max(1.0, abs(want))
        assert abs(a["objval"] - b["objval"]) <= 1e-9 * max(1.0, abs(a["objval"]))


def _fixing_chain(n=8):
    """binary y_0..y_{n-1}, y_0 = 1, y_j + y_{j+1} = 1: fixing y_0 fixes the whole chain, one variable per propagation pass
    (sdpi.c:3220-3225 repeats prepareLPData while a fixing was found); a 1x1 block keeps an SDP part in the problem"""
    M = misdp.Misdp(n, np.r_[np.zeros(5), -1.0, np.zeros(n - 6)], [1])
    M.lb[:] = 0.0; M.ub[:] = 1.0; M.integer[:] = True
    M.add_entry(n - 1, 0, 0, 0, 1.0); M.add_entry(-1, 0, 0, 0, -1.0)          # y_{n-1} + 1 >= 0
    M.add_row({0: 1.0}, 1.0, 1.0)
    for j in range(n - 1):
        M.add_row({j: 1.0, j + 1: 1.0}, 1.0, 1.0)
    return M


def test_fixing_chain_longer_than_four_passes():
    """the chain is propagated to its end by all three restatements of the node presolve (it used to stop after four passes and
    drop the rows that were left with one active variable)"""
    M = _fixing_chain(8)
    want = np.array([1.0, 0.0] * 4)
    for fn in (M.node_problem, M.node_problem_fast):
        st, fp, info = fn(M.lb, M.ub)
        assert st == "allfixed" and np.array_equal(info["y"], want) and info["fixedobj"] == 0.0
    lib = abi.Lib(abi.PRODUCT_LIB)
    model = abi.Model(lib, M)
    st, fp, info = model.node_problem(M.lb, M.ub)
    assert st == 2 and info["fixedobj"] == 0.0
    model.close()
    # the end of the chain contradicts a bound: infeasible, not "optimal with a violated row"
    ub = M.ub.copy(); ub[6] = 0.0
    assert M.node_problem(M.lb, ub)[0] == "infeasible" and M.node_problem_fast(M.lb, ub)[0] == "infeasible"
    from scip_sdp_b200 import frontier
    for native in (False, True):
        r = frontier.branch_and_bound(abi.Solver(abi.Lib(abi.ORACLE_LIB)), M, mode="batch", width=8, native=native)
        assert r["status"] == "optimal" and abs(r["objval"]) <= 1e-9
