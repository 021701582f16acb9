"""ctypes driver for the REFERENCE's solver-independent SDP interface (src/sdpi/sdpi.h, ~70 SCIPsdpi* functions),
compiled unmodified from /root/reference by oracle/Makefile into oracle/_ref/libsdpi_{oracle,cuda}.so together with
our sdpisolver_cuda.c binding.  Test infrastructure: lets the ported reference unit tests (unittests/src/checksdpi.c)
and the B&B harness drive the binding exactly the way relax_sdp.c does (SCIPsdpiLoadSDP / ChgBounds / Solve / getters)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_ORACLE = os.path.join(ROOT, "oracle", "_ref", "libsdpi_oracle.so")
LIB_CUDA = os.path.join(ROOT, "oracle", "_ref", "libsdpi_cuda.so")

INF = 1e20
SCIP_OKAY = 1
# SCIP_SDPPARAM (type_sdpi.h:46-66)
PAR = dict(EPSILON=0, GAPTOL=1, FEASTOL=2, SDPSOLVERFEASTOL=3, OBJLIMIT=4, SDPINFO=5, SLATERCHECK=6, PENALTYPARAM=7,
           MAXPENALTYPARAM=8, NPENALTYINCR=9, LAMBDASTAR=10, NTHREADS=11, WARMSTARTPOGAP=12, PENINFEASADJUST=13)
SETTING_UNSOLVED = -1

_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
_dpp, _ipp = C.POINTER(_dp), C.POINTER(_ip)
_dppp, _ippp = C.POINTER(_dpp), C.POINTER(_ipp)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _Keep:
    """pointer-of-pointer arrays with the numpy buffers kept alive"""

    def __init__(self):
        self.refs = []

    def d(self, a):
        a = _d(a); self.refs.append(a)
        return a.ctypes.data_as(_dp)

    def i(self, a):
        a = _i(a); self.refs.append(a)
        return a.ctypes.data_as(_ip)

    def pp(self, ptrs, typ):
        arr = (typ * max(len(ptrs), 1))(*ptrs)
        self.refs.append(arr)
        return arr


class SdpiLib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.SCIPsdpiGetSolverName.restype = C.c_char_p
        L.BMSgetMemoryUsed.restype = C.c_longlong
        L.BMScreateBlockMemory.restype = C.c_void_p
        L.BMScreateBufferMemory.restype = C.c_void_p
        L.BMScreateBufferMemory.argtypes = [C.c_double, C.c_int, C.c_uint]
        L.SCIPsdpiCreate.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p]
        L.SCIPsdpiFree.argtypes = [C.POINTER(C.c_void_p)]
        L.SCIPsdpiLoadSDP.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_uint), C.c_int, _ip, _ip, C.c_int, _ip,
                                      _ipp, _ipp, _dpp, C.c_int, _ipp, _ipp, _ippp, _ippp, _dppp, C.c_int, _dp, _dp, C.c_int,
                                      _ip, _ip, _dp, C.c_uint]
        L.SCIPsdpiChgBounds.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp]
        L.SCIPsdpiChgObj.argtypes = [C.c_void_p, C.c_int, _ip, _dp]
        L.SCIPsdpiSolve.argtypes = [C.c_void_p, _dp, _ip, _ipp, _ipp, _dpp, _ip, _ipp, _ipp, _dpp, C.c_int, C.c_uint, C.c_double]
        for f in ("WasSolved", "SolvedOrig", "FeasibilityKnown", "IsPrimalUnbounded", "IsPrimalInfeasible", "IsPrimalFeasible",
                  "IsDualUnbounded", "IsDualInfeasible", "IsDualFeasible", "IsConverged", "IsObjlimExc", "IsIterlimExc",
                  "IsTimelimExc", "IsOptimal", "IsAcceptable", "HavePrimalSol"):
            fn = getattr(L, "SCIPsdpi" + f)
            fn.argtypes = [C.c_void_p]
            fn.restype = C.c_uint
        L.SCIPsdpiGetInternalStatus.argtypes = [C.c_void_p]
        L.SCIPsdpiGetSolFeasibility.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.SCIPsdpiGetObjval.argtypes = [C.c_void_p, _dp]
        L.SCIPsdpiGetDualSol.argtypes = [C.c_void_p, _dp, _dp]
        L.SCIPsdpiGetPrimalBoundVars.argtypes = [C.c_void_p, _dp, _dp, C.POINTER(C.c_uint)]
        L.SCIPsdpiGetPrimalLPSides.argtypes = [C.c_void_p, _dp, _dp, C.POINTER(C.c_uint)]
        L.SCIPsdpiGetPrimalSolutionMatrix.argtypes = [C.c_void_p, _dpp, C.POINTER(C.c_uint)]
        L.SCIPsdpiGetIterations.argtypes = [C.c_void_p, _ip]
        L.SCIPsdpiGetSdpCalls.argtypes = [C.c_void_p, _ip]
        L.SCIPsdpiSettingsUsed.argtypes = [C.c_void_p, _ip]
        L.SCIPsdpiGetTime.argtypes = [C.c_void_p, _dp]
        L.SCIPsdpiSetRealpar.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.SCIPsdpiSetIntpar.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.SCIPsdpiGetPreoptimalPrimalNonzeros.argtypes = [C.c_void_p, C.c_int, _ip]
        L.SCIPsdpiGetPreoptimalSol.argtypes = [C.c_void_p, C.POINTER(C.c_uint), _dp, C.c_int, _ip, _ipp, _ipp, _dpp]
        L.SCIPsdpiGetPrimalNonzeros.argtypes = [C.c_void_p, C.c_int, _ip]
        L.SCIPsdpiGetPrimalMatrix.argtypes = [C.c_void_p, C.c_int, _ip, _ipp, _ipp, _dpp]
        L.SCIPsdpiGetNLPRows.argtypes = [C.c_void_p, _ip]
        L.SCIPsdpiGetNVars.argtypes = [C.c_void_p, _ip]

    def solver_name(self):
        return self.lib.SCIPsdpiGetSolverName().decode()

    def memory_used(self):
        return int(self.lib.BMSgetMemoryUsed())


def _ok(rc, what):
    if rc != SCIP_OKAY:
        raise RuntimeError(f"{what} returned SCIP_RETCODE {rc}")


class Sdpi:
    """one SCIP_SDPI object (what relax_sdp.c creates with SCIPsdpiCreate, relax_sdp.c:5387)"""

    def __init__(self, lib, gaptol=1e-6, sdpsolverfeastol=1e-6, feastol=None, sdpinfo=False):
        self.L = lib
        l = lib.lib
        self.blk = C.c_void_p(l.BMScreateBlockMemory(1, 10))
        self.buf = C.c_void_p(l.BMScreateBufferMemory(1.2, 4, 0))
        self.sdpi = C.c_void_p()
        _ok(l.SCIPsdpiCreate(C.byref(self.sdpi), None, self.blk, self.buf), "SCIPsdpiCreate")
        self.set_real("SDPSOLVERFEASTOL", sdpsolverfeastol)
        self.set_real("GAPTOL", gaptol)
        if feastol is not None:
            self.set_real("FEASTOL", feastol)
        if sdpinfo:
            l.SCIPsdpiSetIntpar(self.sdpi, PAR["SDPINFO"], 1)
        self.nvars = self.nrows = 0
        self.blocksizes = []

    def set_real(self, name, val):
        rc = self.L.lib.SCIPsdpiSetRealpar(self.sdpi, PAR[name], float(val))
        if rc not in (SCIP_OKAY, -12):     # SCIP_PARAMETERUNKNOWN is tolerated like SCIP_CALL_PARAM (checksdpi.c:68-78)
            raise RuntimeError(f"SCIPsdpiSetRealpar({name}) returned {rc}")

    def close(self):
        l = self.L.lib
        if self.sdpi:
            _ok(l.SCIPsdpiFree(C.byref(self.sdpi)), "SCIPsdpiFree")
            l.BMSdestroyBufferMemory(C.byref(self.buf))
            l.BMSdestroyBlockMemory(C.byref(self.blk))
            self.sdpi = C.c_void_p()

    # -------------------------------------------------------------- loading
    def load(self, nvars, obj, lb, ub, blocksizes, A, Cmat, rows, allfixedprimalray=True):
        """A[b] = {var: [(row, col, val)...]}, Cmat[b] = [(row, col, val)...], rows = [(coefs dict, lhs, rhs)...]"""
        k = _Keep()
        nb = len(blocksizes)
        nblockvars, constn = [], []
        rowpp, colpp, valpp, varp, nnzp = [], [], [], [], []
        crow, ccol, cval = [], [], []
        sdpnnonz = constnnonz = 0
        for b in range(nb):
            vs = sorted(A[b])
            nblockvars.append(len(vs))
            varp.append(k.i(vs if vs else [0]))
            nnzp.append(k.i([len(A[b][v]) for v in vs] if vs else [0]))
            rp = [k.i([t[0] for t in A[b][v]]) for v in vs]
            cp = [k.i([t[1] for t in A[b][v]]) for v in vs]
            vp = [k.d([t[2] for t in A[b][v]]) for v in vs]
            rowpp.append(C.cast(k.pp(rp, _ip), _ipp)); colpp.append(C.cast(k.pp(cp, _ip), _ipp))
            valpp.append(C.cast(k.pp(vp, _dp), _dpp))
            sdpnnonz += sum(len(A[b][v]) for v in vs)
            constn.append(len(Cmat[b])); constnnonz += len(Cmat[b])
            crow.append(k.i([t[0] for t in Cmat[b]] if Cmat[b] else [0]))
            ccol.append(k.i([t[1] for t in Cmat[b]] if Cmat[b] else [0]))
            cval.append(k.d([t[2] for t in Cmat[b]] if Cmat[b] else [0.0]))
        beg, ind, val, lhs, rhs = [], [], [], [], []
        for coefs, lo, hi in rows:
            beg.append(len(ind))
            for j in sorted(coefs):
                ind.append(j); val.append(coefs[j])
            lhs.append(lo); rhs.append(hi)
        self.nvars, self.nrows, self.blocksizes = nvars, len(rows), list(blocksizes)
        rc = self.L.lib.SCIPsdpiLoadSDP(
            self.sdpi, nvars, k.d(obj), k.d(lb), k.d(ub), None, nb, k.i(blocksizes if nb else [0]),
            k.i(nblockvars if nb else [0]), constnnonz, k.i(constn if nb else [0]),
            C.cast(k.pp(crow, _ip), _ipp), C.cast(k.pp(ccol, _ip), _ipp), C.cast(k.pp(cval, _dp), _dpp),
            sdpnnonz, C.cast(k.pp(nnzp, _ip), _ipp), C.cast(k.pp(varp, _ip), _ipp),
            C.cast(k.pp(rowpp, _ipp), _ippp), C.cast(k.pp(colpp, _ipp), _ippp), C.cast(k.pp(valpp, _dpp), _dppp),
            len(rows), k.d(lhs if rows else [0.0]), k.d(rhs if rows else [0.0]), len(ind),
            k.i(beg if rows else [0]), k.i(ind if ind else [0]), k.d(val if val else [0.0]), int(allfixedprimalray))
        _ok(rc, "SCIPsdpiLoadSDP")

    def load_model(self, M, lb=None, ub=None):
        self.load(M.nvars, M.obj, M.lb if lb is None else lb, M.ub if ub is None else ub, M.blocksizes, M.A, M.C, M.rows)

    def chg_bounds(self, idx, lb, ub):
        k = _Keep()
        _ok(self.L.lib.SCIPsdpiChgBounds(self.sdpi, len(idx), k.i(idx), k.d(lb), k.d(ub)), "SCIPsdpiChgBounds")

    # -------------------------------------------------------------- solving
    def solve(self, starty=None, startsettings=SETTING_UNSOLVED, enforceslater=False, timelimit=1e20, startZ=None, startX=None):
        """startZ / startX: lists of (rows, cols, vals) per SDP block + the LP block last (sdpi.h: SCIPsdpiSolve start point)"""
        sy = _d(starty).ctypes.data_as(_dp) if starty is not None else None
        k = _Keep()
        za = [None] * 4
        xa = [None] * 4
        if startZ is not None and startX is not None:
            for arr, blocks in ((za, startZ), (xa, startX)):
                arr[0] = k.i([len(t[0]) for t in blocks])
                arr[1] = k.pp([k.i(t[0] if len(t[0]) else [0]) for t in blocks], _ip)
                arr[2] = k.pp([k.i(t[1] if len(t[1]) else [0]) for t in blocks], _ip)
                arr[3] = k.pp([k.d(t[2] if len(t[2]) else [0.0]) for t in blocks], _dp)
        rc = self.L.lib.SCIPsdpiSolve(self.sdpi, sy, za[0], za[1], za[2], za[3], xa[0], xa[1], xa[2], xa[3], startsettings,
                                      int(enforceslater), timelimit)
        _ok(rc, "SCIPsdpiSolve")

    def _sparse_blocks(self, counter, getter, what):
        nb = len(self.blocksizes) + 1
        cnt = (C.c_int * nb)()
        _ok(counter(self.sdpi, nb, cnt), what + "Nonzeros")
        if cnt[0] == -1:
            return None, None
        rows = [np.zeros(max(c, 1), dtype=np.int32) for c in cnt]
        cols = [np.zeros(max(c, 1), dtype=np.int32) for c in cnt]
        vals = [np.zeros(max(c, 1)) for c in cnt]
        rp = (_ip * nb)(*[r.ctypes.data_as(_ip) for r in rows])
        cp = (_ip * nb)(*[c.ctypes.data_as(_ip) for c in cols])
        vp = (_dp * nb)(*[v.ctypes.data_as(_dp) for v in vals])
        extra = getter(nb, cnt, C.cast(rp, _ipp), C.cast(cp, _ipp), C.cast(vp, _dpp))
        return extra, [(rows[b][:cnt[b]], cols[b][:cnt[b]], vals[b][:cnt[b]]) for b in range(nb)]

    def primal_matrix_sparse(self):
        """SCIPsdpiGetPrimalNonzeros + SCIPsdpiGetPrimalMatrix"""
        def get(nb, cnt, rp, cp, vp):
            _ok(self.L.lib.SCIPsdpiGetPrimalMatrix(self.sdpi, nb, cnt, rp, cp, vp), "SCIPsdpiGetPrimalMatrix")
            return True
        return self._sparse_blocks(self.L.lib.SCIPsdpiGetPrimalNonzeros, get, "SCIPsdpiGetPrimal")[1]

    def preoptimal_y_only(self):
        """SCIPsdpiGetPreoptimalSol with nblocks = -1 (only the dual vector is wanted, relax_sdp.c:3907)"""
        y = np.zeros(self.nvars)
        ok = C.c_uint(0)
        _ok(self.L.lib.SCIPsdpiGetPreoptimalSol(self.sdpi, C.byref(ok), y.ctypes.data_as(_dp), -1, None, None, None, None), "SCIPsdpiGetPreoptimalSol")
        return y if ok.value else None

    def preoptimal_sol(self):
        """SCIPsdpiGetPreoptimalPrimalNonzeros + SCIPsdpiGetPreoptimalSol -> None or (y, blocks)"""
        y = np.zeros(self.nvars)

        def get(nb, cnt, rp, cp, vp):
            ok = C.c_uint(0)
            _ok(self.L.lib.SCIPsdpiGetPreoptimalSol(self.sdpi, C.byref(ok), y.ctypes.data_as(_dp), nb, cnt, rp, cp, vp), "SCIPsdpiGetPreoptimalSol")
            return bool(ok.value)
        ok, blocks = self._sparse_blocks(self.L.lib.SCIPsdpiGetPreoptimalPrimalNonzeros, get, "SCIPsdpiGetPreoptimalPrimal")
        return (y, blocks) if ok else None

    def flag(self, name):
        return bool(getattr(self.L.lib, "SCIPsdpi" + name)(self.sdpi))

    def sol_feasibility(self):
        p, d = C.c_uint(0), C.c_uint(0)
        _ok(self.L.lib.SCIPsdpiGetSolFeasibility(self.sdpi, C.byref(p), C.byref(d)), "SCIPsdpiGetSolFeasibility")
        return bool(p.value), bool(d.value)

    def dual_sol(self):
        obj = C.c_double(0)
        y = np.zeros(self.nvars)
        _ok(self.L.lib.SCIPsdpiGetDualSol(self.sdpi, C.byref(obj), y.ctypes.data_as(_dp)), "SCIPsdpiGetDualSol")
        return obj.value, y

    def objval(self):
        obj = C.c_double(0)
        _ok(self.L.lib.SCIPsdpiGetObjval(self.sdpi, C.byref(obj)), "SCIPsdpiGetObjval")
        return obj.value

    def primal_bound_vars(self):
        lbv, ubv, ok = np.zeros(self.nvars), np.zeros(self.nvars), C.c_uint(0)
        _ok(self.L.lib.SCIPsdpiGetPrimalBoundVars(self.sdpi, lbv.ctypes.data_as(_dp), ubv.ctypes.data_as(_dp), C.byref(ok)),
            "SCIPsdpiGetPrimalBoundVars")
        return lbv, ubv, bool(ok.value)

    def primal_lp_sides(self):
        n = max(self.nrows, 1)
        l, r, ok = np.zeros(n), np.zeros(n), C.c_uint(0)
        _ok(self.L.lib.SCIPsdpiGetPrimalLPSides(self.sdpi, l.ctypes.data_as(_dp), r.ctypes.data_as(_dp), C.byref(ok)),
            "SCIPsdpiGetPrimalLPSides")
        return l[:self.nrows], r[:self.nrows], bool(ok.value)

    def primal_matrices(self):
        mats = [np.zeros((n, n)) for n in self.blocksizes]
        arr = (_dp * max(len(mats), 1))(*[m.ctypes.data_as(_dp) for m in mats])
        ok = C.c_uint(0)
        _ok(self.L.lib.SCIPsdpiGetPrimalSolutionMatrix(self.sdpi, C.cast(arr, _dpp), C.byref(ok)), "SCIPsdpiGetPrimalSolutionMatrix")
        return mats, bool(ok.value)

    def stats(self):
        it, calls, st, t = C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0)
        self.L.lib.SCIPsdpiGetIterations(self.sdpi, C.byref(it))
        self.L.lib.SCIPsdpiGetSdpCalls(self.sdpi, C.byref(calls))
        self.L.lib.SCIPsdpiSettingsUsed(self.sdpi, C.byref(st))
        self.L.lib.SCIPsdpiGetTime(self.sdpi, C.byref(t))
        return dict(iterations=it.value, sdpcalls=calls.value, setting=st.value, time=t.value)
