"""MISDP model in SCIP-SDP's "dual" form, reader/writer for the extended SDPA format (sdpa_format.txt of the
reference: 1-based, LP block = negative block size, `*INTEGER` / `*RANK1` sections; semantics of
src/scipsdp/reader_sdpa.c), and the flattening into the solver form of include/sdpcuda.h that mirrors what
sdpisolver_cuda.c does in C (sdpisolver_sdpa.cpp:1015-1412: fixed-variable elimination, lhs/rhs splitting, bounds as LP rows).

   min  obj'y   s.t.  sum_j A_j^(k) y_j - A_0^(k) >= 0 (psd),   lhs <= D y <= rhs,   lb <= y <= ub,   y_j integer (j in I)
"""
import gzip
import io

import numpy as np

from .abi import FlatProblem

INF = 1e20


class Misdp:
    def __init__(self, nvars, obj, blocksizes):
        self.nvars = nvars
        self.obj = np.asarray(obj, dtype=float)
        self.blocksizes = list(blocksizes)
        # A[b][j] = list of (row, col, val) with row >= col (0-based); C[b] likewise (this is A_0)
        self.A = [dict() for _ in self.blocksizes]
        self.C = [[] for _ in self.blocksizes]
        self.rows = []          # (dict var->coef, lhs, rhs)
        self.lb = np.full(nvars, -INF)
        self.ub = np.full(nvars, INF)
        self.integer = np.zeros(nvars, dtype=bool)
        self.rank1 = []

    def add_entry(self, j, b, r, c, v):
        if r < c:
            r, c = c, r
        if j < 0:
            self.C[b].append((r, c, v))
        else:
            self.A[b].setdefault(j, []).append((r, c, v))

    def add_row(self, coefs, lhs=-INF, rhs=INF):
        self.rows.append((dict(coefs), lhs, rhs))

    def rows_to_bounds(self):
        """single-variable rows become variable bounds (what sdpi.c:prepareLPData does for every solve, sdpi.c:1131)"""
        keep = []
        for coefs, lhs, rhs in self.rows:
            nz = [(j, a) for j, a in coefs.items() if a != 0.0]
            if len(nz) == 1:
                j, a = nz[0]
                if a > 0:
                    lo = lhs / a if lhs > -INF else -INF
                    hi = rhs / a if rhs < INF else INF
                else:
                    lo = rhs / a if rhs < INF else -INF
                    hi = lhs / a if lhs > -INF else INF
                self.lb[j] = max(self.lb[j], lo)
                self.ub[j] = min(self.ub[j], hi)
            elif len(nz) > 1:
                keep.append((dict(nz), lhs, rhs))
        self.rows = keep
        return self

    # ------------------------------------------------------------------ flattening
    def flatten(self, lb=None, ub=None, epsilon=1e-9, penalty=None):
        """-> (FlatProblem, info) for the given bounds; variables with ub-lb <= epsilon are fixed and eliminated.
        info: active (indices), fixedobj, rowmap [(input row, +1 lhs / -1 rhs)], boundmap [(var, +1 lb / -1 ub)]"""
        lb = self.lb if lb is None else np.asarray(lb, dtype=float)
        ub = self.ub if ub is None else np.asarray(ub, dtype=float)
        fixed = (ub - lb) <= epsilon
        active = np.flatnonzero(~fixed)
        amap = -np.ones(self.nvars, dtype=int)
        amap[active] = np.arange(len(active))
        fixedobj = float(np.dot(self.obj[fixed], lb[fixed]))
        m = len(active)
        per_var = [[] for _ in range(m)]
        cent = []
        for b in range(len(self.blocksizes)):
            for (r, c, v) in self.C[b]:
                cent.append((b, r, c, v))
            for j, ents in self.A[b].items():
                if fixed[j]:
                    for (r, c, v) in ents:      # moves into the constant part: A_0' = A_0 - y_j A_j
                        cent.append((b, r, c, -lb[j] * v))
                else:
                    per_var[amap[j]].extend((b, r, c, v) for (r, c, v) in ents)
        # merge duplicate constant entries
        cm = {}
        for (b, r, c, v) in cent:
            cm[(b, r, c)] = cm.get((b, r, c), 0.0) + v
        cent = [(b, r, c, v) for (b, r, c), v in sorted(cm.items()) if v != 0.0]
        varbeg = [0]
        eb, er, ec, ev = [], [], [], []
        for j in range(m):
            for (b, r, c, v) in sorted(per_var[j]):
                eb.append(b); er.append(r); ec.append(c); ev.append(v)
            varbeg.append(len(eb))
        lpbeg, lpind, lpval, lprhs = [0], [], [], []
        rowmap, boundmap = [], []
        for i, (coefs, lhs, rhs) in enumerate(self.rows):
            const = sum(a * lb[j] for j, a in coefs.items() if fixed[j])
            act = [(amap[j], a) for j, a in sorted(coefs.items()) if not fixed[j] and a != 0.0]
            if not act:
                continue
            if lhs > -INF:
                for j, a in act:
                    lpind.append(j); lpval.append(a)
                lpbeg.append(len(lpind)); lprhs.append(lhs - const); rowmap.append((i, +1))
            if rhs < INF:
                for j, a in act:
                    lpind.append(j); lpval.append(-a)
                lpbeg.append(len(lpind)); lprhs.append(-(rhs - const)); rowmap.append((i, -1))
        for j in active:
            if lb[j] > -INF:
                lpind.append(amap[j]); lpval.append(1.0); lpbeg.append(len(lpind)); lprhs.append(lb[j]); boundmap.append((j, +1))
            if ub[j] < INF:
                lpind.append(amap[j]); lpval.append(-1.0); lpbeg.append(len(lpind)); lprhs.append(-ub[j]); boundmap.append((j, -1))
        fp = FlatProblem(self.obj[active], self.blocksizes, varbeg, eb, er, ec, ev,
                         [t[0] for t in cent], [t[1] for t in cent], [t[2] for t in cent], [t[3] for t in cent],
                         lpbeg, lpind, lpval, lprhs)
        return fp, dict(active=active, fixedobj=fixedobj, rowmap=rowmap, boundmap=boundmap, fixed=fixed)

    # ------------------------------------------------------------------ dense views (tests)
    def dense_Z(self, y):
        out = []
        for b, n in enumerate(self.blocksizes):
            Z = np.zeros((n, n))
            for (r, c, v) in self.C[b]:
                Z[r, c] -= v
                if r != c:
                    Z[c, r] -= v
            for j, ents in self.A[b].items():
                for (r, c, v) in ents:
                    Z[r, c] += y[j] * v
                    if r != c:
                        Z[c, r] += y[j] * v
            out.append(Z)
        return out

    # ------------------------------------------------------------------ SDPA format
    def write_sdpa(self, path):
        """extended SDPA format; bounds are written as LP rows like reader_sdpa.c:2076"""
        rows = []
        for coefs, lhs, rhs in self.rows:
            if lhs > -INF:
                rows.append((coefs, lhs, 1.0))
            if rhs < INF:
                rows.append((coefs, rhs, -1.0))
        for j in range(self.nvars):
            if self.lb[j] > -INF:
                rows.append(({j: 1.0}, self.lb[j], 1.0))
            if self.ub[j] < INF:
                rows.append(({j: 1.0}, self.ub[j], -1.0))
        nb = len(self.blocksizes) + (1 if rows else 0)
        out = io.StringIO()
        out.write(f"{self.nvars}\n{nb}\n")
        out.write(" ".join([str(n) for n in self.blocksizes] + ([str(-len(rows))] if rows else [])) + "\n")
        out.write(" ".join(repr(float(v)) for v in self.obj) + "\n")
        for b in range(len(self.blocksizes)):
            for (r, c, v) in self.C[b]:
                out.write(f"0 {b + 1} {c + 1} {r + 1} {v!r}\n")
            for j in sorted(self.A[b]):
                for (r, c, v) in self.A[b][j]:
                    out.write(f"{j + 1} {b + 1} {c + 1} {r + 1} {v!r}\n")
        for i, (coefs, side, sgn) in enumerate(rows):
            for j in sorted(coefs):
                out.write(f"{j + 1} {nb} {i + 1} {i + 1} {sgn * coefs[j]!r}\n")
            if side != 0.0:
                out.write(f"0 {nb} {i + 1} {i + 1} {sgn * side!r}\n")
        if self.integer.any():
            out.write("*INTEGER\n")
            for j in np.flatnonzero(self.integer):
                out.write(f"*{j + 1}\n")
        if self.rank1:
            out.write("*RANK1\n")
            for b in self.rank1:
                out.write(f"*{b + 1}\n")
        data = out.getvalue()
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "wt") as f:
            f.write(data)


def _tokens(line):
    # SDPA files may separate numbers by blanks, commas, braces or parentheses
    for ch in ",(){}":
        line = line.replace(ch, " ")
    return line.split()


def read_sdpa(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        lines = f.read().splitlines()
    header = []
    k = 0
    # header: nvars, nblocks, block sizes, objective — comment lines start with '*' or '"'
    while len(header) < 2:
        s = lines[k].strip(); k += 1
        if not s or s[0] in '*"':
            continue
        header.append(int(_tokens(s)[0]))
    nvars, nblocks = header
    sizes = []
    while len(sizes) < nblocks:
        s = lines[k].strip(); k += 1
        if not s or s[0] in '*"':
            continue
        for t in _tokens(s):
            if len(sizes) < nblocks:
                try:
                    sizes.append(int(float(t)))
                except ValueError:
                    break
    obj = []
    while len(obj) < nvars:
        s = lines[k].strip(); k += 1
        if not s or s[0] in '*"':
            continue
        for t in _tokens(s):
            if len(obj) < nvars:
                try:
                    obj.append(float(t))
                except ValueError:
                    break
    sdpidx = [b for b, n in enumerate(sizes) if n > 0]
    lpidx = [b for b, n in enumerate(sizes) if n < 0]
    bmap = {b: i for i, b in enumerate(sdpidx)}
    M = Misdp(nvars, obj, [sizes[b] for b in sdpidx])
    lprows = {}
    section = None
    for s in lines[k:]:
        s = s.strip()
        if not s:
            continue
        if s[0] == '*':
            up = s.upper()
            if up.startswith("*INTEGER"):
                section = "int"
            elif up.startswith("*RANK1"):
                section = "rank1"
            elif section is not None and len(s) > 1 and s[1:].strip().split()[0].isdigit():
                idx = int(s[1:].strip().split()[0]) - 1
                if section == "int":
                    M.integer[idx] = True
                else:
                    M.rank1.append(bmap[idx])
            continue
        t = _tokens(s.split('*')[0])
        if len(t) < 5:
            continue
        j, b, r, c, v = int(t[0]) - 1, int(t[1]) - 1, int(t[2]) - 1, int(t[3]) - 1, float(t[4])
        if b in bmap:
            M.add_entry(j, bmap[b], r, c, v)
        else:
            assert b in lpidx and r == c, "LP block entries must be diagonal"
            row = lprows.setdefault((b, r), [dict(), 0.0])
            if j < 0:
                row[1] = v
            else:
                row[0][j] = row[0].get(j, 0.0) + v
    for key in sorted(lprows):
        coefs, rhs = lprows[key]
        M.add_row(coefs, lhs=rhs)      # all LP-block inequalities are  a'y - a_0 >= 0
    return M
