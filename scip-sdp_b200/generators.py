"""Deterministic synthetic instances of the shapes named in BASELINE.json / SURVEY.md section 8(d)
(numpy.random.default_rng(seed); every instance can be written as .dat-s with Misdp.write_sdpa)."""
import itertools

import numpy as np

from .misdp import Misdp, INF


def maxcut(n=2000, p=0.01, seed=4004):
    """max-cut SDP relaxation in SCIP-SDP's dual form:  min sum y_i  s.t.  Diag(y) - L(G)/4 >= 0,  y free.
    One n x n block, n diagonal constraint matrices A_i = e_i e_i'; unit weights on G(n, p)."""
    rng = np.random.default_rng(seed)
    M = Misdp(n, np.ones(n), [n])
    iu = np.triu_indices(n, 1)
    mask = rng.random(len(iu[0])) < p
    rows, cols = iu[1][mask], iu[0][mask]          # row > col
    deg = np.zeros(n)
    np.add.at(deg, rows, 1.0); np.add.at(deg, cols, 1.0)
    for i in range(n):
        M.A[0][i] = [(i, i, 1.0)]
        if deg[i] > 0:
            M.C[0].append((i, i, 0.25 * deg[i]))
    for r, c in zip(rows.tolist(), cols.tolist()):
        M.C[0].append((r, c, -0.25))
    return M


def mkp(n=120, k=4, p=0.5, seed=3003):
    """min k-partitioning with the structure of instances/example_MkP.dat-s.gz: variables y_ab (a > b), A_ab = k/(k-1) at
    (a,b), A_0 = -1 on the diagonal and +1/(k-1) off the diagonal, objective -w_ab, 0 <= y <= 1 integer, per node two
    partition-size rows."""
    rng = np.random.default_rng(seed)
    pairs = [(a, b) for a in range(n) for b in range(a)]
    m = len(pairs)
    w = np.where(rng.random(m) < p, rng.integers(1, 11, m), 0).astype(float)
    M = Misdp(m, -w, [n])
    for j, (a, b) in enumerate(pairs):
        M.A[0][j] = [(a, b, k / (k - 1.0))]
    for a in range(n):
        M.C[0].append((a, a, -1.0))
        for b in range(a):
            M.C[0].append((a, b, 1.0 / (k - 1.0)))
    M.lb[:] = 0.0; M.ub[:] = 1.0; M.integer[:] = True
    idx = {pr: j for j, pr in enumerate(pairs)}
    # weighted partition-size rows as in the example (two per node): lo - t_a <= sum_b t_b y_ab <= hi - t_a
    t = rng.integers(1, 7, n).astype(float)
    W = t.sum()
    lo, hi = np.floor(0.5 * W / k), np.ceil(1.5 * W / k)
    for a in range(n):
        coefs = {idx[(max(a, b), min(a, b))]: t[b] for b in range(n) if b != a}
        M.add_row(coefs, lhs=lo - t[a], rhs=hi - t[a])
    return M


def truss(nx=6, ny=6, nbars=500, seed=1001):
    """truss topology design MISDP with the block structure of instances/example_TT.dat-s.gz: border of two rows carrying
    the compliance variable and the load, rank-1 bar stiffness matrices with three binary size levels per bar."""
    rng = np.random.default_rng(seed)
    nodes = [(i, j) for i in range(nx) for j in range(ny)]
    fixed = {0, ny - 1, 1, ny - 2}
    free = [v for v in range(len(nodes)) if v not in fixed]
    dof = {v: 2 + 2 * t for t, v in enumerate(free)}
    n = 2 + 2 * len(free)
    cand = [(u, v) for u, v in itertools.combinations(range(len(nodes)), 2)]
    length = np.array([np.hypot(nodes[u][0] - nodes[v][0], nodes[u][1] - nodes[v][1]) for u, v in cand])
    order = np.argsort(length + 1e-9 * rng.random(len(cand)))[:nbars]
    nvars = 1 + 3 * len(order)
    obj = np.zeros(nvars)
    M = Misdp(nvars, obj, [n])
    M.A[0][0] = [(0, 0, 2.0), (1, 1, 2.0)]          # compliance variable on the border
    M.lb[0] = 0.0
    for t, ci in enumerate(order):
        u, v = cand[ci]
        d = np.array(nodes[v], float) - np.array(nodes[u], float)
        d /= np.linalg.norm(d)
        gam = {}
        if u in dof:
            gam[dof[u]] = -d[0]; gam[dof[u] + 1] = -d[1]
        if v in dof:
            gam[dof[v]] = d[0]; gam[dof[v] + 1] = d[1]
        ks = sorted(gam)
        for lev in range(3):
            j = 1 + 3 * t + lev
            M.obj[j] = 0.5 * (lev + 1) * length[ci]
            stiff = 100.0 * (lev + 1) / length[ci]
            ents = [(a, b, stiff * gam[a] * gam[b]) for a in ks for b in ks if a >= b and gam[a] * gam[b] != 0.0]
            if ents:
                M.A[0][j] = ents
            M.lb[j] = 0.0; M.ub[j] = 1.0; M.integer[j] = True
        M.add_row({1 + 3 * t: 1.0, 2 + 3 * t: 1.0, 3 + 3 * t: 1.0}, rhs=1.0)
    load_node = free[len(free) // 2]
    M.C[0].append((dof[load_node] + 1, 0, -0.5))   # load vector in the border column
    M.obj[0] = 1.0
    return M


def cls(nfeat=199, nsamp=99, k=10, seed=2002):
    """cardinality-constrained least squares in the shape of instances/example_CLS.dat-s.gz: block of order nsamp + 1,
    nfeat binary indicator variables with dense constraint matrices plus one continuous epigraph variable."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((nsamp, nfeat))
    xs = np.zeros(nfeat); xs[rng.choice(nfeat, k, replace=False)] = rng.standard_normal(k)
    bvec = A @ xs + 0.1 * rng.standard_normal(nsamp)
    n = nsamp + 1
    nvars = nfeat + 1
    rho = 1.0
    M = Misdp(nvars, np.zeros(nvars), [n])
    M.obj[nfeat] = 1.0
    # [ I + (1/rho) sum_j z_j a_j a_j'   b ; b'  t ] >= 0   (Schur complement form of the regularised least-squares value)
    for j in range(nfeat):
        a = A[:, j]
        ents = [(r, c, a[r] * a[c] / rho) for r in range(nsamp) for c in range(r + 1) if abs(a[r] * a[c]) > 1e-3]
        M.A[0][j] = ents
        M.lb[j] = 0.0; M.ub[j] = 1.0; M.integer[j] = True
    M.A[0][nfeat] = [(n - 1, n - 1, 1.0)]
    for r in range(nsamp):
        M.C[0].append((r, r, -1.0))
        M.C[0].append((n - 1, r, -bvec[r]))
    M.add_row({j: 1.0 for j in range(nfeat)}, rhs=float(k))
    return M


def dense_sdp_flat(m, n, seed=5005):
    """Random SDP with m dense constraint matrices of order n in solver form (abi.FlatProblem), strictly feasible on both sides
    by construction: X0 = I, y0 random, S0 = I:  C = sum_j y0_j A_j - S0,  obj_j = <A_j, X0>.  This is the shape whose
    Schur complement costs O(m n^3 + m^2 n^2) flops per iteration (the dense route of SURVEY.md 8d), used for the
    one-SDP-over-several-GPUs measurement; built with numpy arrays directly (m n^2 / 2 entries)."""
    from .abi import FlatProblem
    rng = np.random.default_rng(seed)
    r, c = np.tril_indices(n)
    r, c = r.astype(np.int32), c.astype(np.int32)
    per = len(r)
    vals = rng.standard_normal((m, per)) / np.sqrt(n)
    y0 = rng.standard_normal(m)
    # C = sum y0_j A_j - I (lower triangle)
    cval = y0 @ vals
    cval[r == c] -= 1.0
    obj = vals[:, r == c].sum(axis=1)              # <A_j, I> = trace
    varbeg = np.arange(m + 1, dtype=np.int64) * per
    return FlatProblem(obj, [n], varbeg, np.zeros(m * per, dtype=np.int32), np.tile(r, m), np.tile(c, m), vals.reshape(-1),
                       np.zeros(per, dtype=np.int32), r, c, cval, [0], [], [], [])
