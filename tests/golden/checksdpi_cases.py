"""Known-answer cases of the reference's SDPI unit tests, restated as data.
Source: /root/reference/unittests/src/checksdpi.c (test1 :537-566, test2 :577, test3 :616, test4 :656, test5 :698,
test6 :735, test7 :769, test9 :921-997, test10 :1022-1094, test11 :1113, test12 :1193; test8 is disabled upstream).
Tolerance of every comparison: EPS = 1e-6 (checksdpi.c:49) with SDPSOLVERFEASTOL = GAPTOL = 1e-6 (checksdpi.c:96-97).
Status vocabulary (checksdpi.c:52-58): feas / unbounded / ray / infeas for the primal (X) and dual (y) problem.
Tests 1-4, 9, 10 reach the solver boundary; 5-7 end in sdpi.c presolve; 11-12 in SCIPsolveOneVarSDP (sdpi.c:3301)."""
INF = 1e20

_LP_ROWS = [({0: 2.0, 1: 1.0}, -INF, 10.0), ({0: 1.0, 1: 3.0}, -INF, 15.0)]
_BOX_ROWS = [({0: 1.0}, -1.0, INF), ({0: -1.0}, -1.0, INF), ({1: 1.0}, -1.0, INF), ({1: -1.0}, -1.0, INF)]

CASES = {
    "test1": dict(nvars=2, obj=[-3, -1], lb=[0, 0], ub=[INF, INF], blocksizes=[], A=[], C=[], rows=_LP_ROWS,
                  primal="feas", dual="feas", dualsol=[5, 0], lbvals=[0, 0.5], rhsvals=[1.5, 0], reaches_solver=True),
    "test2": dict(nvars=2, obj=[-3, -1], lb=[-INF, -INF], ub=[INF, INF], blocksizes=[], A=[], C=[], rows=_LP_ROWS,
                  primal="infeas", dual="unbounded", reaches_solver=True),
    "test3": dict(nvars=2, obj=[10, 15], lb=[0, 0], ub=[INF, INF], blocksizes=[], A=[], C=[],
                  rows=[({0: 2.0, 1: 1.0}, 3.0, 3.0), ({0: 1.0, 1: 3.0}, 1.0, 1.0)],
                  primal="unbounded", dual="infeas", reaches_solver=True),
    "test4": dict(nvars=2, obj=[-1, -1], lb=[-INF, -INF], ub=[INF, INF], blocksizes=[], A=[], C=[],
                  rows=[({0: 1.0, 1: -1.0}, -INF, 0.0), ({0: -1.0, 1: 1.0}, -INF, -1.0)],
                  primal="infeas", dual="infeas", reaches_solver=True, skip_for=["DSDP"]),
    "test5": dict(nvars=2, obj=[-3, -1], lb=[0, 0], ub=[0, 0], blocksizes=[], A=[], C=[], rows=_LP_ROWS,
                  primal="feas", dual="feas", dualsol=[0, 0], reaches_solver=False),
    "test6": dict(nvars=2, obj=[-3, -1], lb=[4, 3], ub=[4, 3], blocksizes=[], A=[], C=[], rows=_LP_ROWS,
                  primal="unbounded", dual="infeas", reaches_solver=False),
    "test7": dict(nvars=2, obj=[-3, -1], lb=[1, 0], ub=[0, 10], blocksizes=[], A=[], C=[], rows=_LP_ROWS,
                  primal="unbounded", dual="infeas", reaches_solver=False),
    "test9": dict(nvars=2, obj=[-1, 0], lb=[-INF, -INF], ub=[INF, INF], blocksizes=[2],
                  A=[{0: [(0, 0, 1.0)], 1: [(1, 1, 0.75)]}], C=[[(1, 0, -1.0)]], rows=_BOX_ROWS,
                  primal="unbounded", dual="infeas", reaches_solver=True),
    "test10": dict(nvars=2, obj=[-1, -1], lb=[-INF, -INF], ub=[INF, INF], blocksizes=[2],
                   A=[{0: [(0, 0, 1.0)], 1: [(1, 1, 1.0)]}], C=[[]], rows=_BOX_ROWS,
                   primal="feas", dual="feas", dualsol=[1, 1], lhsvals=[0, 1, 0, 1], rhsvals=[0, 0, 0, 0],
                   primalmatrix=[0, 0, 0, 0], reaches_solver=True),
    "test11": dict(nvars=1, obj=[1], lb=[-INF], ub=[INF], blocksizes=[2], A=[{0: [(0, 0, 1.0), (1, 1, 1.0)]}],
                   C=[[(0, 0, 1.0), (1, 0, 2.0), (1, 1, 4.0)]], rows=[],
                   primal="feas", dual="feas", dualsol=[5], primalmatrix=[0.2, 0.4, 0.4, 0.8], reaches_solver=False),
    "test12": dict(nvars=1, obj=[1], lb=[0], ub=[0], blocksizes=[2], A=[{0: [(0, 0, 1.0), (1, 1, 1.0)]}],
                   C=[[(0, 0, 1.0), (1, 0, 2.0), (1, 1, 4.0)]], rows=[],
                   primal="unbounded", dual="infeas", primalmatrix=[0.2, 0.4, 0.4, 0.8], reaches_solver=False),
}
