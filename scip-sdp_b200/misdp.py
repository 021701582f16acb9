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
        self.indicators = []       # (slack variable, binary variable): binary = 1  =>  slack = 0   (reader_sdpa.c:1195-1246)
        self.objsense = 1          # +1: the file asked for min obj'y; -1: it asked for max (obj is stored negated)
        self.objoffset = 0.0       # constant term of the file's objective (in the file's sense)

    def file_objective(self, value):
        """objective value in the sense and with the constant of the input file, from the internal min-form value"""
        return self.objsense * value + self.objoffset

    def add_entry(self, j, b, r, c, v):
        self._fast = None
        if r < c:
            r, c = c, r
        if j < 0:
            self.C[b].append((r, c, v))
        else:
            self.A[b].setdefault(j, []).append((r, c, v))

    def add_row(self, coefs, lhs=-INF, rhs=INF):
        self._fast = None
        self.rows.append((dict(coefs), lhs, rhs))

    def add_variable(self, obj=0.0, lb=-INF, ub=INF, integer=False):
        """appends a scalar variable (no SDP entries) and returns its index"""
        self.obj = np.append(self.obj, float(obj))
        self.lb = np.append(self.lb, float(lb)); self.ub = np.append(self.ub, float(ub))
        self.integer = np.append(self.integer, bool(integer))
        self.nvars += 1
        self._fast = None
        return self.nvars - 1

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
        self._fast = None
        return self

    # ------------------------------------------------------------------ flattening
    def node_problem(self, lb, ub, epsilon=1e-9, feastol=1e-6):
        """What sdpi.c does to a branch-and-bound node before the solver sees it (SCIPsdpiSolve, sdpi.c:3123-3399), for drivers that
        talk to the C ABI directly: rows without active variables are checked and dropped, rows with one active variable tighten
        its bounds (prepareLPData, sdpi.c:1131-1290), empty rows/columns of the SDP blocks and empty blocks are removed
        (findEmptyRowColsSDP, sdpi.c:691-810), and a node whose variables are all fixed is decided on the spot (sdpi.c:3219-3290).
        -> (status, FlatProblem or None, info); status: "solve" | "infeasible" | "allfixed" (feasible, info["fixedobj"] is its value)"""
        lb, ub = np.array(lb, dtype=float), np.array(ub, dtype=float)
        while True:                                                   # until no bound moves (sdpi.c:3220-3225: while fixingfound)
            if np.any(lb > ub + epsilon):
                return "infeasible", None, {}
            fixed = (ub - lb) <= epsilon
            changed = False
            for coefs, lhs, rhs in self.rows:
                const = sum(a * lb[j] for j, a in coefs.items() if fixed[j])
                act = [(j, a) for j, a in coefs.items() if not fixed[j] and a != 0.0]
                if not act:
                    if const < lhs - feastol or const > rhs + feastol:
                        return "infeasible", None, {}
                elif len(act) == 1:
                    j, a = act[0]
                    lo = (lhs - const) / a if lhs > -INF else -INF
                    hi = (rhs - const) / a if rhs < INF else INF
                    if a < 0:
                        lo, hi = (hi if hi < INF else -INF), (lo if lo > -INF else INF)
                    if lo > lb[j] + epsilon:
                        lb[j] = lo; changed = True
                    if hi < ub[j] - epsilon:
                        ub[j] = hi; changed = True
            if not changed:
                break
        if np.any(lb > ub + epsilon):
            return "infeasible", None, {}
        fixed = (ub - lb) <= epsilon
        if fixed.all():
            Z = self.dense_Z(lb)
            ok = all(np.linalg.eigvalsh(z)[0] >= -feastol for z in Z if z.size)
            return ("allfixed" if ok else "infeasible"), None, dict(fixedobj=float(np.dot(self.obj, lb)), y=lb.copy())
        fp, info = self.flatten(lb, ub, epsilon=epsilon, compress=True, skip_single_rows=True)
        info["lb"], info["ub"] = lb, ub
        return "solve", fp, info

    # ------------------------------------------------------------------ vectorised node marshalling (same results as flatten)
    def _arrays(self):
        """flat numpy views of the model, cached (rebuilt when the entry counts change); orders follow flatten's loops so that
        sums are formed in the same order"""
        sig = (self.nvars, len(self.rows), tuple(map(len, self.C)), tuple(map(len, self.A)))
        cache = getattr(self, "_fast", None)
        if cache is not None and cache["sig"] == sig:
            return cache
        ev, eb, er, ec, ex = [], [], [], [], []            # block-major, constants (var -1) first: the order flatten builds `cent` in
        for b in range(len(self.blocksizes)):
            for (r, c, v) in self.C[b]:
                ev.append(-1); eb.append(b); er.append(r); ec.append(c); ex.append(v)
            for j, ents in self.A[b].items():
                for (r, c, v) in ents:
                    ev.append(j); eb.append(b); er.append(r); ec.append(c); ex.append(v)
        ev, eb, er, ec = (np.asarray(a, dtype=np.int64) for a in (ev, eb, er, ec))
        ex = np.asarray(ex, dtype=float)
        var_order = np.lexsort((ex, ec, er, eb, ev))       # per variable sorted by (block, row, col, value) like sorted(per_var[j])
        var_order = var_order[ev[var_order] >= 0]
        rid, rj, ra = [], [], []                           # row nonzeros in dict order (constants of fixed variables)
        for i, (coefs, lhs, rhs) in enumerate(self.rows):
            for j, a in coefs.items():
                rid.append(i); rj.append(j); ra.append(a)
        rid, rj = np.asarray(rid, dtype=np.int64), np.asarray(rj, dtype=np.int64)
        ra = np.asarray(ra, dtype=float)
        srt = np.lexsort((rj, rid))                        # per row sorted by variable (emission order)
        cache = dict(sig=sig, ev=ev, eb=eb, er=er, ec=ec, ex=ex, var_order=var_order, rid=rid, rj=rj, ra=ra, srt=srt,
                     lhs=np.array([r[1] for r in self.rows], dtype=float), rhs=np.array([r[2] for r in self.rows], dtype=float),
                     maxn=max(self.blocksizes, default=1))
        self._fast = cache
        return cache

    def flatten_fast(self, lb=None, ub=None, epsilon=1e-9, compress=False, skip_single_rows=False, maps=True):
        """flatten(...) without Python loops over entries (identical arrays; tests/test_readers_cpu.py compares them)"""
        F = self._arrays()
        lb = self.lb if lb is None else np.asarray(lb, dtype=float)
        ub = self.ub if ub is None else np.asarray(ub, dtype=float)
        fixed = (ub - lb) <= epsilon
        active = np.flatnonzero(~fixed)
        m = len(active)
        amap = -np.ones(self.nvars + 1, dtype=np.int64)
        amap[active] = np.arange(m)
        fixedobj = float(np.dot(self.obj[fixed], lb[fixed]))
        ev, eb, er, ec, ex, maxn = F["ev"], F["eb"], F["er"], F["ec"], F["ex"], F["maxn"]
        # entries of active variables
        vo = F["var_order"]
        vo = vo[~fixed[ev[vo]]]
        aj, ab, arow, acol, aval = amap[ev[vo]], eb[vo], er[vo], ec[vo], ex[vo]
        varbeg = np.concatenate([[0], np.cumsum(np.bincount(aj, minlength=m))]).astype(np.int32) if m else np.zeros(1, dtype=np.int32)
        # constant part: A_0 and the fixed variables, duplicates summed in input order
        isc = ev < 0
        sel = np.flatnonzero(isc | fixed[np.where(isc, 0, ev)])
        cval = np.where(isc[sel], ex[sel], -lb[np.where(isc[sel], 0, ev[sel])] * ex[sel])
        key = (eb[sel] * maxn + er[sel]) * maxn + ec[sel]
        ukey, inv = np.unique(key, return_inverse=True)
        csum = np.zeros(len(ukey))
        np.add.at(csum, inv, cval)
        keepc = csum != 0.0
        if compress:
            keepc &= np.abs(csum) > epsilon
        ukey, csum = ukey[keepc], csum[keepc]
        cb, cr, cc = ukey // (maxn * maxn), (ukey // maxn) % maxn, ukey % maxn
        blocksizes = list(self.blocksizes)
        if compress:
            nb = len(blocksizes)
            used = np.zeros((nb, maxn), dtype=bool)
            used[ab, arow] = True; used[ab, acol] = True
            used[cb, cr] = True; used[cb, cc] = True
            newidx = np.cumsum(used, axis=1) - 1
            keepb = used.any(axis=1)
            bmap = np.cumsum(keepb) - 1
            blocksizes = [int(x) for x in used.sum(axis=1)[keepb]]
            arow, acol, ab = newidx[ab, arow], newidx[ab, acol], bmap[ab]
            cr, cc, cb = newidx[cb, cr], newidx[cb, cc], bmap[cb]
        # LP rows: constants of the fixed variables (dict order), active entries sorted by variable
        rid, rj, ra, srt = F["rid"], F["rj"], F["ra"], F["srt"]
        nrows = len(self.rows)
        fx = fixed[rj] if len(rj) else np.zeros(0, dtype=bool)
        const = np.bincount(rid[fx], weights=ra[fx] * lb[rj[fx]], minlength=nrows) if nrows else np.zeros(0)
        srid, srj, sra = rid[srt], rj[srt], ra[srt]
        actm = ~fixed[srj] & (sra != 0.0) if len(srj) else np.zeros(0, dtype=bool)
        srid, srj, sra = srid[actm], amap[srj[actm]], sra[actm]
        nact = np.bincount(srid, minlength=nrows) if nrows else np.zeros(0, dtype=np.int64)
        rbeg = np.concatenate([[0], np.cumsum(nact)])
        keep = nact >= (2 if skip_single_rows else 1)
        el = np.flatnonzero(keep & (F["lhs"] > -INF))
        eh = np.flatnonzero(keep & (F["rhs"] < INF))
        erow = np.concatenate([el, eh])
        esgn = np.concatenate([np.ones(len(el)), -np.ones(len(eh))])
        order = np.lexsort((-esgn, erow))                  # per row: lhs part first
        erow, esgn = erow[order], esgn[order]
        cnt = nact[erow]
        start = np.repeat(rbeg[erow], cnt)
        within = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        src = start + within
        lpind = srj[src]
        lpval = sra[src] * np.repeat(esgn, cnt)
        lprhs = np.where(esgn > 0, F["lhs"][erow] - const[erow], -(F["rhs"][erow] - const[erow]))
        rowmap = [(int(i), int(sg)) for i, sg in zip(erow, esgn)] if maps else None
        # variable bounds as rows: per active variable lb first, then ub
        hl = lb[active] > -INF
        hu = ub[active] < INF
        bj = np.concatenate([np.flatnonzero(hl), np.flatnonzero(hu)])
        bs = np.concatenate([np.ones(int(hl.sum())), -np.ones(int(hu.sum()))])
        order = np.lexsort((-bs, bj))
        bj, bs = bj[order], bs[order]
        brhs = np.where(bs > 0, lb[active][bj], -ub[active][bj])
        lpbeg = np.concatenate([[0], np.cumsum(cnt), int(cnt.sum()) + 1 + np.arange(len(bj))]).astype(np.int32)
        boundmap = [(int(active[j]), int(sg)) for j, sg in zip(bj, bs)] if maps else None
        fp = FlatProblem(self.obj[active], blocksizes, varbeg, ab, arow, acol, aval, cb, cr, cc, csum,
                         lpbeg, np.concatenate([lpind, bj]), np.concatenate([lpval, bs]), np.concatenate([lprhs, brhs]))
        return fp, dict(active=active, fixedobj=fixedobj, rowmap=rowmap, boundmap=boundmap, fixed=fixed)

    def node_problem_fast(self, lb, ub, epsilon=1e-9, feastol=1e-6):
        """node_problem(...) on the cached arrays (identical results)"""
        F = self._arrays()
        lb, ub = np.array(lb, dtype=float), np.array(ub, dtype=float)
        rid, rj, ra = F["rid"], F["rj"], F["ra"]
        nrows = len(self.rows)
        while True:                                                   # until no bound moves (sdpi.c:3220-3225: while fixingfound)
            if np.any(lb > ub + epsilon):
                return "infeasible", None, {}
            if nrows == 0:
                break
            fixed = (ub - lb) <= epsilon
            fx = fixed[rj]
            const = np.bincount(rid[fx], weights=ra[fx] * lb[rj[fx]], minlength=nrows)
            am = ~fx & (ra != 0.0)
            nact = np.bincount(rid[am], minlength=nrows)
            none = nact == 0
            if np.any(none & ((const < F["lhs"] - feastol) | (const > F["rhs"] + feastol))):
                return "infeasible", None, {}
            one = np.flatnonzero(am & (nact[rid] == 1))
            if len(one) == 0:
                break
            # rows that can tighten anything at all (bounds only move inwards during the pass), then sequentially like node_problem
            i1, j1, a1 = rid[one], rj[one], ra[one]
            with np.errstate(over="ignore", invalid="ignore"):
                lo1 = np.where(F["lhs"][i1] > -INF, (F["lhs"][i1] - const[i1]) / a1, -INF)
                hi1 = np.where(F["rhs"][i1] < INF, (F["rhs"][i1] - const[i1]) / a1, INF)
            neg = a1 < 0
            lo2 = np.where(neg, np.where(hi1 < INF, hi1, -INF), lo1)
            hi2 = np.where(neg, np.where(lo1 > -INF, lo1, INF), hi1)
            one = one[(lo2 > lb[j1] + epsilon) | (hi2 < ub[j1] - epsilon)]
            changed = False
            for t in one:
                i, j, a = rid[t], rj[t], ra[t]
                lhs, rhs = F["lhs"][i], F["rhs"][i]
                lo = (lhs - const[i]) / a if lhs > -INF else -INF
                hi = (rhs - const[i]) / a if rhs < INF else INF
                if a < 0:
                    lo, hi = (hi if hi < INF else -INF), (lo if lo > -INF else INF)
                if lo > lb[j] + epsilon:
                    lb[j] = lo; changed = True
                if hi < ub[j] - epsilon:
                    ub[j] = hi; changed = True
            if not changed:
                break
        if np.any(lb > ub + epsilon):
            return "infeasible", None, {}
        fixed = (ub - lb) <= epsilon
        if fixed.all():
            Z = self.dense_Z(lb)
            ok = all(np.linalg.eigvalsh(z)[0] >= -feastol for z in Z if z.size)
            return ("allfixed" if ok else "infeasible"), None, dict(fixedobj=float(np.dot(self.obj, lb)), y=lb.copy())
        fp, info = self.flatten_fast(lb, ub, epsilon=epsilon, compress=True, skip_single_rows=True, maps=False)
        info["lb"], info["ub"] = lb, ub
        return "solve", fp, info

    def flatten(self, lb=None, ub=None, epsilon=1e-9, penalty=None, compress=False, skip_single_rows=False):
        """-> (FlatProblem, info) for the given bounds; variables with ub-lb <= epsilon are fixed and eliminated.
        info: active (indices), fixedobj, rowmap [(input row, +1 lhs / -1 rhs)], boundmap [(var, +1 lb / -1 ub)].
        compress: rows/columns of a block that carry no entry of an active variable and no constant entry are removed, blocks
        without entries too (sdpi.c:691-810); skip_single_rows: rows with one active variable are left out (node_problem has
        turned them into bounds)"""
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
        blocksizes = list(self.blocksizes)
        if compress:
            used = [set() for _ in blocksizes]
            for ents in per_var:
                for (b, r, c, v) in ents:
                    used[b].update((r, c))
            for (b, r, c, v) in cent:
                if abs(v) > epsilon:
                    used[b].update((r, c))
            bmap, imap, blocksizes = {}, {}, []
            for b, u in enumerate(used):
                if u:
                    bmap[b] = len(blocksizes)
                    imap[b] = {i: k for k, i in enumerate(sorted(u))}
                    blocksizes.append(len(u))
            per_var = [[(bmap[b], imap[b][r], imap[b][c], v) for (b, r, c, v) in ents] for ents in per_var]
            cent = [(bmap[b], imap[b][r], imap[b][c], v) for (b, r, c, v) in cent if abs(v) > epsilon]
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
            if not act or (skip_single_rows and len(act) == 1):
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
        objv = self.obj[active]
        if penalty is not None:
            # penalty formulation (SCIPsdpiSolverLoadAndSolveWithPenalty, sdpisolver.h:258-322; sdpisolver_sdpa.cpp:1232-1239,
            # 1338-1357,1405-1410): one more variable r with objective gamma, + r I on every block and + r in every LP row (not in
            # the variable bounds); withobj=False drops the original objective, rbound adds r >= 0
            gamma, withobj, rbound = penalty
            nrows = len(rowmap)
            for b, n in enumerate(blocksizes):
                for i in range(n):
                    eb.append(b); er.append(i); ec.append(i); ev.append(1.0)
            varbeg.append(len(eb))
            newbeg, newind, newval = [0], [], []
            for l in range(len(lprhs)):
                newind += lpind[lpbeg[l]:lpbeg[l + 1]]; newval += lpval[lpbeg[l]:lpbeg[l + 1]]
                if l < nrows:
                    newind.append(m); newval.append(1.0)
                newbeg.append(len(newind))
            lpbeg, lpind, lpval = newbeg, newind, newval
            if rbound:
                lpind.append(m); lpval.append(1.0); lpbeg.append(len(lpind)); lprhs.append(0.0)
            objv = np.concatenate([objv if withobj else np.zeros(m), [gamma]])
            if not withobj:
                fixedobj = 0.0
        fp = FlatProblem(objv, blocksizes, varbeg, eb, er, ec, ev,
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
                out.write(f"0 {b + 1} {c + 1} {r + 1} {float(v)!r}\n")
            for j in sorted(self.A[b]):
                for (r, c, v) in self.A[b][j]:
                    out.write(f"{j + 1} {b + 1} {c + 1} {r + 1} {float(v)!r}\n")
        for i, (coefs, side, sgn) in enumerate(rows):
            for j in sorted(coefs):
                out.write(f"{j + 1} {nb} {i + 1} {i + 1} {float(sgn * coefs[j])!r}\n")
            if side != 0.0:
                out.write(f"0 {nb} {i + 1} {i + 1} {float(sgn * side)!r}\n")
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


    def write_cbf(self, path):
        """CBF version 2 in the dual form that reader_cbf.c writes (reader_cbf.c:2780-3440): scalar variables with their sign cones
        (L+ / L- / L= / F; other finite bounds become linear constraints), INT, linear constraints grouped as L= , L+ , L-,
        one PSDCON per block with HCOORD (the A_j) and DCOORD (= -A_0), PSDCONRANK1, objective in the file's own sense"""
        if self.indicators:
            raise ValueError("indicator constraints cannot be written in CBF")
        cones, extra = [], []
        for j in range(self.nvars):
            lo, hi = self.lb[j], self.ub[j]
            if lo == 0.0 and hi == 0.0:
                cone = "L="
            elif lo == 0.0:
                cone = "L+"
            elif hi == 0.0:
                cone = "L-"
            else:
                cone = "F"
            if lo > -INF and lo != 0.0:
                extra.append(({j: 1.0}, lo, INF))
            if hi < INF and hi != 0.0:
                extra.append(({j: 1.0}, -INF, hi))
            if cones and cones[-1][0] == cone:
                cones[-1][1] += 1
            else:
                cones.append([cone, 1])
        eq, ge, le = [], [], []              # (coefficients, constant b) of  a'x + b  in  L= / L+ / L-
        for coefs, lhs, rhs in list(self.rows) + extra:
            if lhs > -INF and rhs < INF and lhs == rhs:
                eq.append((coefs, -lhs))
            else:
                if lhs > -INF:
                    ge.append((coefs, -lhs))
                if rhs < INF:
                    le.append((coefs, -rhs))
        cons = eq + ge + le
        out = io.StringIO()
        out.write("VER\n2\n\nOBJSENSE\n" + ("MIN" if self.objsense == 1 else "MAX") + "\n\n")
        out.write(f"VAR\n{self.nvars} {len(cones)}\n" + "".join(f"{c} {n}\n" for c, n in cones) + "\n")
        ints = np.flatnonzero(self.integer)
        if len(ints):
            out.write(f"INT\n{len(ints)}\n" + "".join(f"{j}\n" for j in ints) + "\n")
        if cons:
            groups = [(c, len(g)) for c, g in (("L=", eq), ("L+", ge), ("L-", le)) if g]
            out.write(f"CON\n{len(cons)} {len(groups)}\n" + "".join(f"{c} {n}\n" for c, n in groups) + "\n")
        if self.blocksizes:
            out.write(f"PSDCON\n{len(self.blocksizes)}\n" + "".join(f"{n}\n" for n in self.blocksizes) + "\n")
        if self.rank1:
            out.write(f"PSDCONRANK1\n{len(self.rank1)}\n" + "".join(f"{b}\n" for b in self.rank1) + "\n")
        objfile = self.objsense * self.obj
        nz = np.flatnonzero(objfile)
        out.write(f"OBJACOORD\n{len(nz)}\n" + "".join(f"{j} {float(objfile[j])!r}\n" for j in nz) + "\n")
        if self.objoffset != 0.0:
            out.write(f"OBJBCOORD\n{float(self.objoffset)!r}\n\n")
        if cons:
            ac = [(i, j, a) for i, (coefs, b) in enumerate(cons) for j, a in sorted(coefs.items()) if a != 0.0]
            out.write(f"ACOORD\n{len(ac)}\n" + "".join(f"{i} {j} {float(a)!r}\n" for i, j, a in ac) + "\n")
            bc = [(i, b) for i, (coefs, b) in enumerate(cons) if b != 0.0]
            if bc:
                out.write(f"BCOORD\n{len(bc)}\n" + "".join(f"{i} {float(b)!r}\n" for i, b in bc) + "\n")
        hc = [(b, j, r, c, v) for b in range(len(self.blocksizes)) for j in sorted(self.A[b]) for (r, c, v) in self.A[b][j]]
        if hc:
            out.write(f"HCOORD\n{len(hc)}\n" + "".join(f"{b} {j} {r} {c} {float(v)!r}\n" for b, j, r, c, v in hc) + "\n")
        dc = [(b, r, c, -v) for b in range(len(self.blocksizes)) for (r, c, v) in self.C[b]]
        if dc:
            out.write(f"DCOORD\n{len(dc)}\n" + "".join(f"{b} {r} {c} {float(v)!r}\n" for b, r, c, v in dc) + "\n")
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "wt") as f:
            f.write(out.getvalue())


def _tokens(line):
    # SDPA files may separate numbers by blanks, commas, braces or parentheses
    for ch in ",(){}":
        line = line.replace(ch, " ")
    return line.split()


class SdpaFormatError(ValueError):
    """malformed extended-SDPA file (what reader_sdpa.c answers with SCIP_READERROR)"""


def _numbers(line, what, lineno, count=None, integer=False):
    """leading numbers of a header line; the rest of the line (after '=', '*' or '"') is a comment like in the shipped instances"""
    for ch in '=*"':
        line = line.split(ch)[0]
    out = []
    for t in _tokens(line):
        try:
            out.append(int(t) if integer else float(t))
        except ValueError:
            raise SdpaFormatError(f"line {lineno}: invalid symbol '{t}' in {what}") from None
    if count is not None and len(out) != count:
        raise SdpaFormatError(f"line {lineno}: expected {count} value(s) for {what}, found {len(out)}")
    return out


def read_sdpa(path):
    """Reader for the extended SDPA format of SCIP-SDP (sdpa_format.txt; checks as in src/scipsdp/reader_sdpa.c:500-1660, whose
    malformed unit-test files unittests/instances/*.dat-s must be rejected): 1-based indices, one LP block given by a negative
    block size with the linear constraints on its diagonal, `*INTEGER` then `*RANK1` sections, indicator entries (variable -k)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        lines = f.read().splitlines()
    k = 0

    def next_data_line(what):
        nonlocal k
        while k < len(lines):
            s = lines[k].strip(); k += 1
            if s and s[0] not in '*"':
                return s, k
        raise SdpaFormatError(f"unexpected end of file while reading {what}")

    s, ln = next_data_line("the number of variables")
    nvars = _numbers(s, "the number of variables", ln, integer=True)[:1]
    if not nvars or nvars[0] < 0:
        raise SdpaFormatError(f"line {ln}: the number of variables must be a non-negative integer")
    nvars = nvars[0]
    s, ln = next_data_line("the number of blocks")
    nblocks = _numbers(s, "the number of blocks", ln, integer=True)[:1]
    if not nblocks or nblocks[0] < 0:
        raise SdpaFormatError(f"line {ln}: the number of blocks must be a non-negative integer")
    nblocks = nblocks[0]
    s, ln = next_data_line("the block sizes")
    sizes = _numbers(s, "the block sizes", ln, count=nblocks, integer=True)
    if any(n == 0 for n in sizes):
        raise SdpaFormatError(f"line {ln}: a block size of 0 is not valid")
    if sum(1 for n in sizes if n < 0) > 1:
        raise SdpaFormatError(f"line {ln}: only one LP block can be defined")
    s, ln = next_data_line("the objective")
    obj = _numbers(s, "the objective coefficients", ln, count=nvars)
    sdpidx = [b for b, n in enumerate(sizes) if n > 0]
    lpidx = [b for b, n in enumerate(sizes) if n < 0]
    nlin = -sizes[lpidx[0]] if lpidx else 0
    bmap = {b: i for i, b in enumerate(sdpidx)}
    M = Misdp(nvars, obj, [sizes[b] for b in sdpidx])
    lprows = {}
    lpcoefs = [0] * nlin
    blocknnz = [0] * len(sdpidx)
    indrows = []
    section = None
    for ln, raw in enumerate(lines[k:], start=k + 1):
        s = raw.strip()
        if not s:
            continue
        up = s.upper()
        if up.startswith("*INTEGER"):
            if section == "rank1":
                raise SdpaFormatError(f"line {ln}: the integer section has to be in front of the rank-1 section")
            section = "int"
            continue
        if up.startswith("*RANK1"):
            section = "rank1"
            continue
        if section is not None:
            # inside a section every entry is '*<index>'
            if s[0] != '*':
                raise SdpaFormatError(f"line {ln}: expected '*' at the beginning of the line in the {section.upper()} section")
            tok = s[1:].split()
            if not tok or not tok[0].lstrip("+-").isdigit():
                raise SdpaFormatError(f"line {ln}: could not read the index in the {section.upper()} section")
            idx = int(tok[0]) - 1
            if section == "int":
                if idx < 0 or idx >= nvars:
                    raise SdpaFormatError(f"line {ln}: integrality given for variable {idx + 1} which does not exist")
                M.integer[idx] = True
            else:
                if idx in lpidx:
                    raise SdpaFormatError(f"line {ln}: rank-1 given for the LP block")
                if idx not in bmap:
                    raise SdpaFormatError(f"line {ln}: rank-1 given for SDP block {idx + 1} which does not exist")
                M.rank1.append(bmap[idx])
            continue
        if s[0] in '*"':
            continue                                  # comment line
        t = _numbers(s, "a block entry", ln)
        if len(t) < 5:
            raise SdpaFormatError(f"line {ln}: could not read block entry (variable block row column value)")
        if any(x != int(x) for x in t[:4]):
            raise SdpaFormatError(f"line {ln}: indices of a block entry must be integers")
        j, b, r, c, v = int(t[0]) - 1, int(t[1]) - 1, int(t[2]) - 1, int(t[3]) - 1, float(t[4])
        if b in bmap:
            n = sizes[b]
            if j < -1 or j >= nvars:
                raise SdpaFormatError(f"line {ln}: coefficient for variable {j + 1} which does not exist")
            if r < 0 or r >= n or c < 0 or c >= n:
                raise SdpaFormatError(f"line {ln}: row/column index outside the block of size {n}")
            M.add_entry(j, bmap[b], r, c, v)
            if j >= 0:
                blocknnz[bmap[b]] += 1
        elif b in lpidx:
            if j >= nvars:
                raise SdpaFormatError(f"line {ln}: linear coefficient for variable {j + 1} which does not exist")
            if r != c:
                raise SdpaFormatError(f"line {ln}: linear coefficient is not located on the diagonal of the LP block")
            if r < 0 or r >= nlin:
                raise SdpaFormatError(f"line {ln}: linear constraint {r + 1} does not exist")
            row = lprows.setdefault((b, r), [dict(), 0.0])
            if j < -1:
                # indicator constraint (file index -k, k >= 2): variable k-1 becomes binary, the row gets a slack variable s >= 0
                # and "variable = 1 => s = 0" (reader_sdpa.c:1195-1246; the value of the entry is not used there either)
                if -j - 2 >= nvars:
                    raise SdpaFormatError(f"line {ln}: indicator variable {-j - 1} does not exist")
                indrows.append(((b, r), -j - 2))
            elif j < 0:
                row[1] = v
            else:
                row[0][j] = row[0].get(j, 0.0) + v
                lpcoefs[r] += 1
        else:
            raise SdpaFormatError(f"line {ln}: coefficient for block {b + 1} which does not exist")
    for i, cnt in enumerate(blocknnz):
        if cnt == 0:
            raise SdpaFormatError(f"SDP block {sdpidx[i] + 1} does not contain any nonzero entries")
    for r, cnt in enumerate(lpcoefs):
        if cnt == 0:
            raise SdpaFormatError(f"linear constraint {r + 1} does not contain nonzero entries")
    for key, z in indrows:
        sl = M.add_variable(obj=0.0, lb=0.0)
        lprows[key][0][sl] = 1.0
        M.lb[z], M.ub[z], M.integer[z] = max(M.lb[z], 0.0), min(M.ub[z], 1.0), True
        M.indicators.append((sl, z))
    for key in sorted(lprows):
        coefs, rhs = lprows[key]
        M.add_row(coefs, lhs=rhs)      # all LP-block inequalities are  a'y - a_0 >= 0
    return M


# ---------------------------------------------------------------------- CBF (conic benchmark format)
def read_cbf(path):
    """Reader for the subset of CBF that src/scipsdp/reader_cbf.c accepts (VER 1-3): scalar variables in the cones F/L+/L-/L=,
    matrix variables PSDVAR, scalar constraints CON in L=/L+/L-, LMIs PSDCON, INT, and the coordinate sections
    OBJFCOORD/OBJACOORD/OBJBCOORD/FCOORD/ACOORD/BCOORD/HCOORD/DCOORD.  Conventions as in the reference:
      * constraint i:  sum_j a_ij x_j + sum_v <F_iv, X_v> + b_i  in  its cone   (reader_cbf.c:720-853, 1512-1744)
      * LMI k:         sum_j H_kj x_j + D_k  psd                                   (reader_cbf.c:1744-2160)  => A_0 = -D_k
      * a matrix variable X_v of order n becomes n(n+1)/2 scalar variables (lower triangle) and an LMI sum x_ij E_ij psd;
        off-diagonal coefficients of F / the objective count twice (reader_cbf.c:521-720, 1195-1211, 1479-1492)
      * OBJSENSE MAX is turned into min of the negated objective (Misdp.objsense = -1).
    Second-order cones and the RANK1 sections are recorded but not supported by the B&B harness."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        raw = [ln.split('#')[0].strip() for ln in f.read().splitlines()]
    lines = [ln for ln in raw if ln]
    pos = 0

    def take():
        nonlocal pos
        ln = lines[pos]; pos += 1
        return ln.split()

    sense = 1
    nscal = 0
    varcones, concones = [], []
    psdvar, psdcon = [], []
    ints = []
    sec = {k: [] for k in ("OBJFCOORD", "OBJACOORD", "OBJBCOORD", "FCOORD", "ACOORD", "BCOORD", "HCOORD", "DCOORD")}
    rank1con, rank1var = [], []
    ncons = 0
    while pos < len(lines):
        key = take()[0].upper()
        if key == "VER":
            ver = int(take()[0])
            if ver not in (1, 2, 3):
                raise ValueError(f"unsupported CBF version {ver}")
        elif key == "OBJSENSE":
            sense = -1 if take()[0].upper() == "MAX" else 1
        elif key == "VAR":
            nscal, nc = (int(t) for t in take())
            for _ in range(nc):
                cone, cnt = take()
                varcones.append((cone.upper(), int(cnt)))
        elif key == "CON":
            ncons, nc = (int(t) for t in take())
            for _ in range(nc):
                cone, cnt = take()
                concones.append((cone.upper(), int(cnt)))
        elif key in ("PSDVAR", "PSDCON"):
            cnt = int(take()[0])
            (psdvar if key == "PSDVAR" else psdcon).extend(int(take()[0]) for _ in range(cnt))
        elif key == "INT":
            cnt = int(take()[0])
            ints.extend(int(take()[0]) for _ in range(cnt))
        elif key in ("PSDCONRANK1", "PSDVARRANK1"):
            cnt = int(take()[0])
            (rank1con if key == "PSDCONRANK1" else rank1var).extend(int(take()[0]) for _ in range(cnt))
        elif key == "OBJBCOORD":
            sec[key].append(take())
        elif key in sec:
            cnt = int(take()[0])
            sec[key].extend(take() for _ in range(cnt))
        else:
            raise ValueError(f"CBF section {key} is not supported")

    # scalar variables of the file, then the entries of the matrix variables
    psdoff, nvars = [], nscal
    for n in psdvar:
        psdoff.append(nvars)
        nvars += n * (n + 1) // 2

    def xidx(v, r, c):
        if r < c:
            r, c = c, r
        return psdoff[v] + r * (r + 1) // 2 + c

    M = Misdp(nvars, np.zeros(nvars), list(psdcon) + list(psdvar))
    M.objsense = sense
    j = 0
    for cone, cnt in varcones:
        for _ in range(cnt):
            if cone == "L+":
                M.lb[j] = 0.0
            elif cone == "L-":
                M.ub[j] = 0.0
            elif cone == "L=":
                M.lb[j] = M.ub[j] = 0.0
            elif cone != "F":
                raise ValueError(f"variable cone {cone} is not supported")
            j += 1
    assert j == nscal, "VAR cone sizes do not add up"
    for i in ints:
        M.integer[i] = True
    obj = np.zeros(nvars)
    for v, r, c, val in sec["OBJFCOORD"]:
        v, r, c, val = int(v), int(r), int(c), float(val)
        obj[xidx(v, r, c)] += val if r == c else 2.0 * val
    for jv, val in sec["OBJACOORD"]:
        obj[int(jv)] += float(val)
    for t in sec["OBJBCOORD"]:
        M.objoffset += float(t[0])
    M.obj = sense * obj
    rows = [dict() for _ in range(ncons)]
    bconst = np.zeros(ncons)
    for i, v, r, c, val in sec["FCOORD"]:
        i, v, r, c, val = int(i), int(v), int(r), int(c), float(val)
        k = xidx(v, r, c)
        rows[i][k] = rows[i].get(k, 0.0) + (val if r == c else 2.0 * val)
    for i, jv, val in sec["ACOORD"]:
        i, jv = int(i), int(jv)
        rows[i][jv] = rows[i].get(jv, 0.0) + float(val)
    for i, val in sec["BCOORD"]:
        bconst[int(i)] += float(val)
    i = 0
    for cone, cnt in concones:
        for _ in range(cnt):
            if cone == "L+":
                M.add_row(rows[i], lhs=-bconst[i])
            elif cone == "L-":
                M.add_row(rows[i], rhs=-bconst[i])
            elif cone == "L=":
                M.add_row(rows[i], lhs=-bconst[i], rhs=-bconst[i])
            else:
                raise ValueError(f"constraint cone {cone} is not supported")
            i += 1
    assert i == ncons, "CON cone sizes do not add up"
    for k, jv, r, c, val in sec["HCOORD"]:
        M.add_entry(int(jv), int(k), int(r), int(c), float(val))
    for k, r, c, val in sec["DCOORD"]:
        M.add_entry(-1, int(k), int(r), int(c), -float(val))
    for v, n in enumerate(psdvar):
        b = len(psdcon) + v
        for r in range(n):
            for c in range(r + 1):
                M.add_entry(xidx(v, r, c), b, r, c, 1.0)
    M.rank1 = list(rank1con) + [len(psdcon) + v for v in rank1var]
    return M


def read_instance(path):
    """dispatch on the file name like the reference's reader plugins (reader_sdpa.c: dat-s, reader_cbf.c: cbf)"""
    name = str(path)
    if name.endswith(".gz"):
        name = name[:-3]
    return read_cbf(path) if name.endswith(".cbf") else read_sdpa(path)
