"""ctypes view of the C ABI in include/sdpcuda.h (test/bench harness side; the product host code is C:
scip-sdp_b200/sdpi/sdpisolver_cuda.c).  The same wrapper can load the product library (lib/libsdpcuda.so) or —
from tests and bench.py's cpu_baseline only — the CPU oracle (oracle/liboracle_sdp.so)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.path.join(ROOT, "scip-sdp_b200", "lib", "libsdpcuda.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle_sdp.so")

PHASES = ["noINFO", "pFEAS", "dFEAS", "pdFEAS", "pdINF", "pFEAS_dINF", "pINF_dFEAS", "pdOPT", "pUNBD", "dUNBD", "dINF"]
STOPS = ["converged", "infeascert", "numerics", "objlimit", "iterlimit", "timelimit"]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Problem(C.Structure):
    _fields_ = [("m", C.c_int), ("obj", _dp), ("nblocks", C.c_int), ("blocksizes", _ip), ("varbeg", _ip),
                ("entblk", _ip), ("entrow", _ip), ("entcol", _ip), ("entval", _dp),
                ("cnnz", C.c_int), ("cblk", _ip), ("crow", _ip), ("ccol", _ip), ("cval", _dp),
                ("nlp", C.c_int), ("lpbeg", _ip), ("lpind", _ip), ("lpval", _dp), ("lprhs", _dp)]


class Params(C.Structure):
    _fields_ = [("gaptol", C.c_double), ("feastol", C.c_double), ("objlimit", C.c_double), ("lambdastar", C.c_double),
                ("timelimit", C.c_double), ("absgaptol", C.c_double), ("maxiter", C.c_int), ("setting", C.c_int),
                ("verbose", C.c_int), ("reserved", C.c_int), ("preoptgap", C.c_double)]


class Result(C.Structure):
    _fields_ = [("phase", C.c_int), ("stop", C.c_int), ("iterations", C.c_int), ("launches", C.c_int),
                ("pobj", C.c_double), ("dobj", C.c_double), ("relgap", C.c_double), ("pinf", C.c_double),
                ("dinf", C.c_double), ("mu", C.c_double), ("seconds", C.c_double), ("device_ms", C.c_double),
                ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double)]


RESULT_DTYPE = np.dtype([(name, np.int32 if ctype is C.c_int else np.float64) for name, ctype in Result._fields_], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class FlatProblem:
    """Solver-form problem (see include/sdpcuda.h): numpy arrays kept alive next to the ctypes struct."""

    def __init__(self, obj, blocksizes, varbeg, entblk, entrow, entcol, entval, cblk, crow, ccol, cval, lpbeg, lpind,
                 lpval, lprhs):
        self.obj, self.blocksizes, self.varbeg = _d(obj), _i(blocksizes), _i(varbeg)
        self.entblk, self.entrow, self.entcol, self.entval = _i(entblk), _i(entrow), _i(entcol), _d(entval)
        self.cblk, self.crow, self.ccol, self.cval = _i(cblk), _i(crow), _i(ccol), _d(cval)
        self.lpbeg, self.lpind, self.lpval, self.lprhs = _i(lpbeg), _i(lpind), _d(lpval), _d(lprhs)
        self.m, self.nblocks, self.nlp = len(self.obj), len(self.blocksizes), len(self.lprhs)
        assert len(self.varbeg) == self.m + 1 and len(self.lpbeg) == self.nlp + 1

    def struct(self):
        if getattr(self, "_struct", None) is not None:      # the arrays are never modified after construction
            return self._struct
        p = Problem()
        p.m, p.nblocks, p.nlp, p.cnnz = self.m, self.nblocks, self.nlp, len(self.cval)
        for name in ("obj", "entval", "cval", "lpval", "lprhs"):
            setattr(p, name, getattr(self, name).ctypes.data_as(_dp))
        for name in ("blocksizes", "varbeg", "entblk", "entrow", "entcol", "cblk", "crow", "ccol", "lpbeg", "lpind"):
            setattr(p, name, getattr(self, name).ctypes.data_as(_ip))
        self._struct = p
        return p

    # dense helpers used by the tests for a-posteriori KKT checks
    def dense_A(self, j):
        mats = [np.zeros((n, n)) for n in self.blocksizes]
        for e in range(self.varbeg[j], self.varbeg[j + 1]):
            b, r, c, v = self.entblk[e], self.entrow[e], self.entcol[e], self.entval[e]
            mats[b][r, c] += v
            if r != c:
                mats[b][c, r] += v
        return mats

    def dense_C(self):
        mats = [np.zeros((n, n)) for n in self.blocksizes]
        for b, r, c, v in zip(self.cblk, self.crow, self.ccol, self.cval):
            mats[b][r, c] += v
            if r != c:
                mats[b][c, r] += v
        return mats

    def dense_D(self):
        D = np.zeros((self.nlp, self.m))
        for l in range(self.nlp):
            for p in range(self.lpbeg[l], self.lpbeg[l + 1]):
                D[l, self.lpind[p]] += self.lpval[p]
        return D


class Lib:
    """One loaded implementation of the C ABI."""

    def __init__(self, path=PRODUCT_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.lib = L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.sdpcuda_backend_name.restype = C.c_char_p
        L.sdpcuda_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.sdpcuda_destroy.argtypes = [C.c_void_p]
        L.sdpcuda_default_params.argtypes = [C.POINTER(Params)]
        L.sdpcuda_default_params.restype = None
        L.sdpcuda_solve.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Params), _dp, C.POINTER(Result)]
        L.sdpcuda_solve_resident.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Result)]
        L.sdpcuda_solve_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(Problem)), C.POINTER(Params), C.POINTER(Result), C.POINTER(_dp), _dp]
        L.sdpcuda_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.sdpcuda_get_profile.argtypes = [C.c_void_p, _dp]
        for f in ("sdpcuda_get_y", "sdpcuda_get_xlp", "sdpcuda_get_slp"):
            getattr(L, f).argtypes = [C.c_void_p, _dp]
        for f in ("sdpcuda_get_X", "sdpcuda_get_S"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, _dp]
        L.sdpcuda_syev_batched.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
        L.sdpcuda_dgemm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _dp, C.c_int,
                                    _dp, C.c_int, C.c_double, _dp, C.c_int]
        L.sdpcuda_dpotrf.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _ip]
        L.sdpcuda_dtrtri.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int]
        L.sdpcuda_set_start_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp]
        L.sdpcuda_set_start_lp.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.sdpcuda_get_preopt.argtypes = [C.c_void_p, _ip, _dp, _dp]
        L.sdpcuda_get_preopt_X.argtypes = [C.c_void_p, C.c_int, _dp]
        L.sdpcuda_dist_unique_id.argtypes = [C.c_void_p]
        L.sdpcuda_dist_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.sdpcuda_dist_finalize.argtypes = [C.c_void_p]
        L.sdpcuda_dpotrf_inv.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _dp, C.c_int, _ip]
        L.sdpcuda_psd_check.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_double, _ip]
        L.sdpcuda_time_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.sdpcuda_check_psd_resident.argtypes = [C.c_void_p, _dp, C.c_double, _ip]
        L.sdpcuda_model_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, _dp, C.c_int, _ip, C.c_int, _ip, _ip, _ip, _ip, _dp, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.sdpcuda_model_destroy.argtypes = [C.c_void_p]
        L.sdpcuda_solve_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, _dp, C.POINTER(Params), _dp, _ip, C.POINTER(Result), _dp, _dp, _dp, _dp]
        L.sdpcuda_debug_node_problem.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _ip, _dp, _ip, _ip, C.c_size_t, _dp, C.c_size_t, _dp, _dp]

    def backend(self):
        return self.lib.sdpcuda_backend_name().decode()

    def default_params(self, **kw):
        p = Params()
        self.lib.sdpcuda_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class Model:
    """sdpcuda_model: a mixed-integer SDP handed over once (scip_sdp_b200.misdp.Misdp), so that whole frontiers of its nodes can be
    given to Solver.solve_nodes as bound vectors only; node presolve and marshalling then run in C++ (csrc/node_marshal.hpp)"""

    def __init__(self, lib, misdp_model):
        F = misdp_model._arrays()                  # entries block-major with the constant part first, rows in input order
        self.L, self.nvars = lib, misdp_model.nvars
        rowbeg = np.concatenate([[0], np.cumsum(np.bincount(F["rid"], minlength=len(misdp_model.rows)))]) if len(misdp_model.rows) else np.zeros(1)
        keep = [_d(misdp_model.obj), _i(misdp_model.blocksizes), _i(F["ev"]), _i(F["eb"]), _i(F["er"]), _i(F["ec"]), _d(F["ex"]),
                _i(rowbeg), _i(F["rj"]), _d(F["ra"]), _d(F["lhs"]), _d(F["rhs"])]
        self.h = C.c_void_p()
        ip = lambda a: a.ctypes.data_as(_ip)       # noqa: E731
        dp = lambda a: a.ctypes.data_as(_dp)       # noqa: E731
        rc = lib.lib.sdpcuda_model_create(C.byref(self.h), self.nvars, dp(keep[0]), len(keep[1]), ip(keep[1]), len(keep[2]), ip(keep[2]), ip(keep[3]),
                                          ip(keep[4]), ip(keep[5]), dp(keep[6]), len(misdp_model.rows), ip(keep[7]), ip(keep[8]), dp(keep[9]),
                                          dp(keep[10]), dp(keep[11]))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_model_create failed with code {rc}")

    def close(self):
        if self.h:
            self.L.lib.sdpcuda_model_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def node_problem(self, lb, ub, feastol=1e-6):
        """test hook: -> (status 0 solve / 1 infeasible / 2 all fixed, FlatProblem or None, dict(fixedobj, active, lb, ub))"""
        lb, ub = _d(lb), _d(ub)
        st, fo = C.c_int(-1), C.c_double(0)
        sizes = (C.c_int * 6)()
        lbo, ubo = np.zeros(self.nvars), np.zeros(self.nvars)
        F = self.L.lib.sdpcuda_debug_node_problem
        args = (self.h, lb.ctypes.data_as(_dp), ub.ctypes.data_as(_dp), feastol, C.byref(st), C.byref(fo), sizes)
        assert F(*args, None, 0, None, 0, lbo.ctypes.data_as(_dp), ubo.ctypes.data_as(_dp)) == 0
        info = dict(fixedobj=fo.value, lb=lbo, ub=ubo)
        if st.value != 0:
            return st.value, None, info
        m, nb, nnz, cn, nlp, lnz = [int(v) for v in sizes]
        ib = np.zeros(nb + (m + 1) + 3 * nnz + 3 * cn + (nlp + 1) + lnz + m, dtype=np.int32)
        db = np.zeros(m + nnz + cn + lnz + nlp)
        assert F(*args, ib.ctypes.data_as(_ip), ib.size, db.ctypes.data_as(_dp), db.size, lbo.ctypes.data_as(_dp), ubo.ctypes.data_as(_dp)) == 0
        it, dt = iter(np.split(ib, np.cumsum([nb, m + 1, nnz, nnz, nnz, cn, cn, cn, nlp + 1, lnz]))), iter(np.split(db, np.cumsum([m, nnz, cn, lnz])))
        bs, varbeg, eb, er, ec, cb, cr, cc, lpbeg, lpind, active = [next(it) for _ in range(11)]
        obj, ev, cv, lpval, lprhs = [next(dt) for _ in range(5)]
        info["active"] = active
        return 0, FlatProblem(obj, bs, varbeg, eb, er, ec, ev, cb, cr, cc, cv, lpbeg, lpind, lpval, lprhs), info


class Solver:
    def __init__(self, lib, device=-1):
        self.L = lib
        self.h = C.c_void_p()
        rc = lib.lib.sdpcuda_create(C.byref(self.h), device)
        if rc != 0:
            raise RuntimeError(f"sdpcuda_create failed with code {rc} ({lib.path})")
        self.prob = None

    def close(self):
        if self.h:
            self.L.lib.sdpcuda_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dist_unique_id(self):
        """rank 0: the 128-byte NCCL id that the other ranks need for dist_init"""
        buf = C.create_string_buffer(128)
        rc = self.L.lib.sdpcuda_dist_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"sdpcuda_dist_unique_id failed with code {rc} (NCCL missing?)")
        return buf.raw

    def dist_init(self, nranks, rank, idbytes):
        rc = self.L.lib.sdpcuda_dist_init(self.h, nranks, rank, C.create_string_buffer(idbytes, 128) if idbytes is not None else None)
        if rc != 0:
            raise RuntimeError(f"sdpcuda_dist_init failed with code {rc}")

    def dist_finalize(self):
        self.L.lib.sdpcuda_dist_finalize(self.h)

    def set_start(self, X, S, xlp=None, slp=None):
        """stage a warm start (dense blocks, LP multipliers/slacks) for the next solve(); it is used only together with start_y"""
        for which, blocks in ((0, X), (1, S)):
            for b, A in enumerate(blocks):
                Af = np.ascontiguousarray(A, dtype=np.float64)
                rc = self.L.lib.sdpcuda_set_start_block(self.h, which, b, Af.shape[0], Af.ctypes.data_as(_dp))
                if rc != 0:
                    raise RuntimeError(f"sdpcuda_set_start_block failed with code {rc}")
        if xlp is not None:
            xf, sf = _d(xlp), _d(slp)
            rc = self.L.lib.sdpcuda_set_start_lp(self.h, len(xf), xf.ctypes.data_as(_dp), sf.ctypes.data_as(_dp))
            if rc != 0:
                raise RuntimeError(f"sdpcuda_set_start_lp failed with code {rc}")

    def get_preopt(self):
        """-> None or dict(y, X, xlp): the iterate saved when the gap first dropped below params.preoptgap"""
        ex = C.c_int(0)
        y = np.zeros(max(self.prob.m, 1)); x = np.zeros(max(self.prob.nlp, 1))
        rc = self.L.lib.sdpcuda_get_preopt(self.h, C.byref(ex), y.ctypes.data_as(_dp), x.ctypes.data_as(_dp))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_get_preopt failed with code {rc}")
        if not ex.value:
            return None
        X = []
        for b, n in enumerate(self.prob.blocksizes):
            A = np.zeros((int(n), int(n)))
            rc = self.L.lib.sdpcuda_get_preopt_X(self.h, b, A.ctypes.data_as(_dp))
            if rc != 0:
                raise RuntimeError(f"sdpcuda_get_preopt_X failed with code {rc}")
            X.append(A)
        return dict(y=y[:self.prob.m], X=X, xlp=x[:self.prob.nlp])

    def solve(self, prob, params=None, start_y=None, fetch=True, **kw):
        params = params if params is not None else self.L.default_params(**kw)
        self.prob = prob
        st = prob.struct()
        res = Result()
        sy = _d(start_y).ctypes.data_as(_dp) if start_y is not None else None
        rc = self.L.lib.sdpcuda_solve(self.h, C.byref(st), C.byref(params), sy, C.byref(res))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_solve failed with code {rc}")
        out = {f[0]: getattr(res, f[0]) for f in Result._fields_}
        out["phase_name"], out["stop_name"] = PHASES[res.phase], STOPS[res.stop]
        if fetch:
            out["y"] = self.get_y()
            out["X"] = [self.get_X(b) for b in range(prob.nblocks)]
            out["S"] = [self.get_S(b) for b in range(prob.nblocks)]
            out["xlp"], out["slp"] = self.get_xlp(), self.get_slp()
        return out

    def solve_batch(self, probs, params=None, fetch=True, objlimits=None, **kw):
        """sdpcuda_solve_batch: all relaxations in `probs` that fit the single-CTA kernel in ONE launch (one CTA per node), the others
        one by one on this handle.  Returns one result dict per node (with "y" unless fetch=False).  The matrix getters of the
        handle do not refer to batched nodes.  objlimits: per-node stop value for the lower bound (phase pUNBD when exceeded)."""
        n = len(probs)
        if n == 0:
            return []
        params = params if params is not None else self.L.default_params(**kw)
        structs = [p.struct() for p in probs]
        ps = (C.POINTER(Problem) * n)(*[C.pointer(st) for st in structs])
        res = (Result * n)()
        ys = [np.zeros(max(p.m, 1)) for p in probs] if fetch else None
        yp = (_dp * n)(*[y.ctypes.data_as(_dp) for y in ys]) if fetch else None
        ol = _d(objlimits) if objlimits is not None else None
        assert ol is None or len(ol) == n
        rc = self.L.lib.sdpcuda_solve_batch(self.h, n, ps, C.byref(params), res, yp, ol.ctypes.data_as(_dp) if ol is not None else None)
        if rc != 0:
            raise RuntimeError(f"sdpcuda_solve_batch failed with code {rc}")
        self.prob = probs[-1]
        out = []
        for i, r in enumerate(res):
            d = {f[0]: getattr(r, f[0]) for f in Result._fields_}
            d["phase_name"], d["stop_name"] = PHASES[r.phase], STOPS[r.stop]
            if fetch:
                d["y"] = ys[i][:probs[i].m]
            out.append(d)
        return out

    def solve_nodes(self, model, lbs, ubs, params=None, cutoff=None, lean=False, **kw):
        """sdpcuda_solve_nodes: nodes of `model` (abi.Model) given by their bound vectors (arrays count x nvars); presolve, marshalling
        and the batch launch happen in the library.  -> dict(status [count] (0 solved, 1 infeasible by presolve, 2 all fixed),
        results (list of dicts), bound, y (count x nvars, model variables), lb, ub (tightened))"""
        lbs, ubs = _d(lbs), _d(ubs)
        n, nv = lbs.shape
        assert nv == model.nvars and ubs.shape == lbs.shape
        params = params if params is not None else self.L.default_params(**kw)
        status = np.zeros(n, dtype=np.int32)
        res = (Result * n)()
        bound, y, lbo, ubo = np.zeros(n), np.zeros((n, nv)), np.zeros((n, nv)), np.zeros((n, nv))
        co = _d(cutoff) if cutoff is not None else None
        dp = lambda a: a.ctypes.data_as(_dp) if a is not None else None      # noqa: E731
        rc = self.L.lib.sdpcuda_solve_nodes(self.h, model.h, n, dp(lbs), dp(ubs), C.byref(params), dp(co), status.ctypes.data_as(_ip), res,
                                            dp(bound), dp(y), dp(lbo), dp(ubo))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_solve_nodes failed with code {rc}")
        if lean:          # the result structs as one numpy record array (fields of sdpcuda_result), no per-node Python objects
            return dict(status=status, results=np.frombuffer(res, dtype=RESULT_DTYPE, count=n), bound=bound, y=y, lb=lbo, ub=ubo, keep=res)
        results = []
        for r in res:
            d = {f[0]: getattr(r, f[0]) for f in Result._fields_}
            d["phase_name"], d["stop_name"] = PHASES[r.phase], STOPS[r.stop]
            results.append(d)
        return dict(status=status, results=results, bound=bound, y=y, lb=lbo, ub=ubo)

    def solve_resident(self, params=None, **kw):
        params = params if params is not None else self.L.default_params(**kw)
        res = Result()
        rc = self.L.lib.sdpcuda_solve_resident(self.h, C.byref(params), C.byref(res))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_solve_resident failed with code {rc}")
        out = {f[0]: getattr(res, f[0]) for f in Result._fields_}
        out["phase_name"], out["stop_name"] = PHASES[res.phase], STOPS[res.stop]
        return out

    PROF_CLASSES = ["gemm_dmma", "diag_block", "schur", "eig", "trsv", "elementwise", "gemm_dmma_small"]

    def set_profiling(self, on):
        self.L.lib.sdpcuda_set_profiling(self.h, int(on))

    def get_profile(self):
        a = np.zeros(3 * len(self.PROF_CLASSES))
        self.L.lib.sdpcuda_get_profile(self.h, a.ctypes.data_as(_dp))
        return {c: dict(launches=int(a[3 * i]), ms=float(a[3 * i + 1]), work=float(a[3 * i + 2])) for i, c in enumerate(self.PROF_CLASSES)}

    def _vec(self, fn, n):
        a = np.zeros(max(n, 1))
        rc = fn(self.h, a.ctypes.data_as(_dp))
        if rc != 0:
            raise RuntimeError(f"getter failed with code {rc}")
        return a[:n]

    def get_y(self):
        return self._vec(self.L.lib.sdpcuda_get_y, self.prob.m)

    def get_xlp(self):
        return self._vec(self.L.lib.sdpcuda_get_xlp, self.prob.nlp)

    def get_slp(self):
        return self._vec(self.L.lib.sdpcuda_get_slp, self.prob.nlp)

    def _mat(self, fn, b):
        n = int(self.prob.blocksizes[b])
        a = np.zeros((n, n))
        rc = fn(self.h, b, a.ctypes.data_as(_dp))
        if rc != 0:
            raise RuntimeError(f"getter failed with code {rc}")
        return a

    def get_X(self, b):
        return self._mat(self.L.lib.sdpcuda_get_X, b)

    def get_S(self, b):
        return self._mat(self.L.lib.sdpcuda_get_S, b)

    # ---- kernel-level entry points ----
    def syev(self, A, vectors=True):
        A = _d(A)
        single = A.ndim == 2
        A3 = A[None] if single else A
        nb, n = A3.shape[0], A3.shape[1]
        w = np.zeros((nb, n))
        V = np.zeros((nb, n, n)) if vectors else None
        rc = self.L.lib.sdpcuda_syev_batched(self.h, n, nb, A3.ctypes.data_as(_dp), w.ctypes.data_as(_dp),
                                             V.ctypes.data_as(_dp) if vectors else None)
        if rc != 0:
            raise RuntimeError(f"sdpcuda_syev_batched failed with code {rc}")
        if single:
            return w[0], (V[0] if vectors else None)
        return w, V

    def dgemm(self, A, B, transa=False, transb=False, alpha=1.0, beta=0.0, Cin=None):
        """column-major semantics; A, B given as numpy (row-major) arrays are passed as their Fortran copies"""
        Af, Bf = np.asfortranarray(A, dtype=np.float64), np.asfortranarray(B, dtype=np.float64)
        m = Af.shape[1] if transa else Af.shape[0]
        k = Af.shape[0] if transa else Af.shape[1]
        n = Bf.shape[0] if transb else Bf.shape[1]
        Cf = np.asfortranarray(np.zeros((m, n)) if Cin is None else Cin, dtype=np.float64).copy(order="F")
        rc = self.L.lib.sdpcuda_dgemm(self.h, int(transa), int(transb), m, n, k, alpha, Af.ctypes.data_as(_dp),
                                      Af.shape[0], Bf.ctypes.data_as(_dp), Bf.shape[0], beta, Cf.ctypes.data_as(_dp), m)
        if rc != 0:
            raise RuntimeError(f"sdpcuda_dgemm failed with code {rc}")
        return Cf

    def dpotrf(self, A):
        Af = np.asfortranarray(A, dtype=np.float64).copy(order="F")
        info = C.c_int(0)
        rc = self.L.lib.sdpcuda_dpotrf(self.h, Af.shape[0], Af.ctypes.data_as(_dp), Af.shape[0], C.byref(info))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_dpotrf failed with code {rc}")
        return np.tril(Af), info.value

    def dpotrf_inv(self, A):
        Af = np.asfortranarray(A, dtype=np.float64).copy(order="F")
        Li = np.zeros_like(Af, order="F")
        info = C.c_int(0)
        n = Af.shape[0]
        rc = self.L.lib.sdpcuda_dpotrf_inv(self.h, n, Af.ctypes.data_as(_dp), n, Li.ctypes.data_as(_dp), n, C.byref(info))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_dpotrf_inv failed with code {rc}")
        return np.tril(Af), Li, info.value

    def dtrtri(self, Lm):
        Lf = np.asfortranarray(Lm, dtype=np.float64).copy(order="F")
        rc = self.L.lib.sdpcuda_dtrtri(self.h, Lf.shape[0], Lf.ctypes.data_as(_dp), Lf.shape[0])
        if rc != 0:
            raise RuntimeError(f"sdpcuda_dtrtri failed with code {rc}")
        return np.tril(Lf)

    def psd_check(self, A, shift=0.0):
        Af = np.asfortranarray(A, dtype=np.float64)
        ok = C.c_int(0)
        rc = self.L.lib.sdpcuda_psd_check(self.h, Af.shape[0], Af.ctypes.data_as(_dp), Af.shape[0], shift, C.byref(ok))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_psd_check failed with code {rc}")
        return bool(ok.value)

    def check_psd_resident(self, y=None, shift=0.0):
        """is sum_j y_j A_j - C + shift I positive definite for every block of the problem resident from the last solve?
        y = None: the solution of the last solve (already on the device)"""
        ok = C.c_int(0)
        yp = _d(y).ctypes.data_as(_dp) if y is not None else None
        rc = self.L.lib.sdpcuda_check_psd_resident(self.h, yp, shift, C.byref(ok))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_check_psd_resident failed with code {rc}")
        return bool(ok.value)

    def time_kernel(self, kind, n, reps=10):
        ms, work = C.c_double(0), C.c_double(0)
        rc = self.L.lib.sdpcuda_time_kernel(self.h, kind, n, reps, C.byref(ms), C.byref(work))
        if rc != 0:
            raise RuntimeError(f"sdpcuda_time_kernel failed with code {rc}")
        return ms.value, work.value
