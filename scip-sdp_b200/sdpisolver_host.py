"""ctypes view of the SCIP-SDP solver boundary (src/sdpi/sdpisolver.h) as implemented by sdpi/sdpisolver_cuda.c in
lib/libsdpisolver_cuda.so.  It passes HOST buffers in exactly the layout sdpi.c hands to SCIPsdpiSolverLoadAndSolve
(pointer-of-pointer sparse triplets, lower triangles, CSR LP rows without sentinel, indchanges/blockindchanges maps), so
bench.py's end-to-end number and the tests exercise the reference-facing call itself."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINDING_LIB = os.path.join(ROOT, "scip-sdp_b200", "lib", "libsdpisolver_cuda.so")

_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
_dpp, _ipp = C.POINTER(_dp), C.POINTER(_ip)
_dppp, _ippp = C.POINTER(_dpp), C.POINTER(_ipp)
SCIP_OKAY = 1
INF = 1e20


class _Keep:
    def __init__(self):
        self.refs = []

    def d(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64); self.refs.append(a)
        return a.ctypes.data_as(_dp)

    def i(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32); self.refs.append(a)
        return a.ctypes.data_as(_ip)

    def pp(self, ptrs, typ, cast):
        arr = (typ * max(len(ptrs), 1))(*ptrs)
        self.refs.append(arr)
        return C.cast(arr, cast)


class BoundaryProblem:
    """the argument list of SCIPsdpiSolverLoadAndSolve for one Misdp with given bounds, marshalled once"""

    def __init__(self, M, lb=None, ub=None):
        k = self.keep = _Keep()
        lb = M.lb if lb is None else lb
        ub = M.ub if ub is None else ub
        nb = len(M.blocksizes)
        self.nvars, self.nblocks, self.blocksizes = M.nvars, nb, list(M.blocksizes)
        nblockvars, constn, crow, ccol, cval, nnzp, varp, rowpp, colpp, valpp, indch = [], [], [], [], [], [], [], [], [], [], []
        sdpnnonz = constnnonz = 0
        for b in range(nb):
            vs = sorted(M.A[b])
            nblockvars.append(len(vs))
            varp.append(k.i(vs if vs else [0]))
            nnzp.append(k.i([len(M.A[b][v]) for v in vs] if vs else [0]))
            rowpp.append(k.pp([k.i([t[0] for t in M.A[b][v]]) for v in vs], _ip, _ipp))
            colpp.append(k.pp([k.i([t[1] for t in M.A[b][v]]) for v in vs], _ip, _ipp))
            valpp.append(k.pp([k.d([t[2] for t in M.A[b][v]]) for v in vs], _dp, _dpp))
            sdpnnonz += sum(len(M.A[b][v]) for v in vs)
            constn.append(len(M.C[b])); constnnonz += len(M.C[b])
            crow.append(k.i([t[0] for t in M.C[b]] if M.C[b] else [0]))
            ccol.append(k.i([t[1] for t in M.C[b]] if M.C[b] else [0]))
            cval.append(k.d([t[2] for t in M.C[b]] if M.C[b] else [0.0]))
            indch.append(k.i(np.zeros(M.blocksizes[b], dtype=np.int32)))
        beg, ind, val, lhs, rhs = [], [], [], [], []
        for coefs, lo, hi in M.rows:
            beg.append(len(ind))
            for j in sorted(coefs):
                ind.append(j); val.append(coefs[j])
            lhs.append(lo); rhs.append(hi)
        self.nlpcons = len(M.rows)
        self.args = [
            M.nvars, k.d(M.obj), k.d(lb), k.d(ub), nb, k.i(M.blocksizes if nb else [0]), k.i(nblockvars if nb else [0]),
            constnnonz, k.i(constn if nb else [0]), k.pp(crow, _ip, _ipp), k.pp(ccol, _ip, _ipp), k.pp(cval, _dp, _dpp),
            sdpnnonz, k.pp(nnzp, _ip, _ipp), k.pp(varp, _ip, _ipp), k.pp(rowpp, _ipp, _ippp), k.pp(colpp, _ipp, _ippp),
            k.pp(valpp, _dpp, _dppp), k.pp(indch, _ip, _ipp), k.i(np.zeros(max(nb, 1), dtype=np.int32)),
            k.i(np.zeros(max(nb, 1), dtype=np.int32)), 0,
            self.nlpcons, k.i(np.zeros(max(self.nlpcons, 1), dtype=np.int32)), k.d(lhs if lhs else [0.0]), k.d(rhs if rhs else [0.0]),
            len(ind), k.i(beg if beg else [0]), k.i(ind if ind else [0]), k.d(val if val else [0.0]),
            None, None, None, None, None, None, None, None, None,      # starty, startZ*, startX*
            -1, 1e20, None]                                            # startsettings UNSOLVED, no time limit, no clock
        self.host_bytes = sum(a.nbytes for a in k.refs if isinstance(a, np.ndarray))


class SdpiSolver:
    """one SCIP_SDPISOLVER object"""

    def __init__(self, path=BINDING_LIB, gaptol=1e-5, feastol=1e-5):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built: run __graft_entry__.build()")
        self.lib = L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.BMScreateBlockMemory.restype = C.c_void_p
        L.BMScreateBufferMemory.restype = C.c_void_p
        L.BMScreateBufferMemory.argtypes = [C.c_double, C.c_int, C.c_uint]
        L.SCIPsdpiSolverGetSolverName.restype = C.c_char_p
        L.SCIPsdpiSolverCreate.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p]
        L.SCIPsdpiSolverFree.argtypes = [C.POINTER(C.c_void_p)]
        L.SCIPsdpiSolverLoadAndSolve.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, C.c_int, _ip, _ipp, _ipp, _dpp,
                                                 C.c_int, _ipp, _ipp, _ippp, _ippp, _dppp, _ipp, _ip, _ip, C.c_int,
                                                 C.c_int, _ip, _dp, _dp, C.c_int, _ip, _ip, _dp,
                                                 _dp, _ip, _ipp, _ipp, _dpp, _ip, _ipp, _ipp, _dpp, C.c_int, C.c_double, C.c_void_p]
        L.SCIPsdpiSolverLoadAndSolveWithPenalty.argtypes = [C.c_void_p, C.c_double, C.c_uint, C.c_uint] + L.SCIPsdpiSolverLoadAndSolve.argtypes[1:] + \
            [C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.SCIPsdpiSolverGetObjval.argtypes = [C.c_void_p, _dp]
        L.SCIPsdpiSolverGetPrimalBoundVars.argtypes = [C.c_void_p, _dp, _dp]
        L.SCIPsdpiSolverGetPrimalNonzeros.argtypes = [C.c_void_p, C.c_int, _ip]
        L.SCIPsdpiSolverGetPrimalMatrix.argtypes = [C.c_void_p, C.c_int, _ip, _ipp, _ipp, _dpp]
        L.SCIPsdpiSolverSettingsUsed.argtypes = [C.c_void_p, _ip]
        L.SCIPsdpiSolverSetRealpar.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.SCIPsdpiSolverGetDualSol.argtypes = [C.c_void_p, _dp, _dp]
        L.SCIPsdpiSolverGetIterations.argtypes = [C.c_void_p, _ip]
        L.SCIPsdpiSolverGetSdpCalls.argtypes = [C.c_void_p, _ip]
        L.SCIPsdpiSolverGetPrimalSolutionMatrix.argtypes = [C.c_void_p, C.c_int, _ip, _ipp, _ip, _ip, _dpp]
        L.SCIPsdpiSolverGetPreoptimalPrimalNonzeros.argtypes = [C.c_void_p, C.c_int, _ip]
        L.SCIPsdpiSolverGetPreoptimalSol.argtypes = [C.c_void_p, C.POINTER(C.c_uint), _dp, C.c_int, _ip, _ipp, _ipp, _dpp]
        L.SCIPsdpiSolverDoesWarmstartNeedPrimal.restype = C.c_uint
        for f in ("WasSolved", "IsAcceptable", "IsOptimal", "IsDualInfeasible", "IsDualFeasible", "IsPrimalFeasible", "IsConverged"):
            fn = getattr(L, "SCIPsdpiSolver" + f)
            fn.argtypes = [C.c_void_p]; fn.restype = C.c_uint
        self.blk = C.c_void_p(L.BMScreateBlockMemory(1, 10))
        self.buf = C.c_void_p(L.BMScreateBufferMemory(1.2, 4, 0))
        self.s = C.c_void_p()
        rc = L.SCIPsdpiSolverCreate(C.byref(self.s), None, self.blk, self.buf)
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverCreate returned {rc} (no CUDA device?)")
        for par, val in ((1, gaptol), (2, feastol), (3, feastol)):      # GAPTOL, FEASTOL, SDPSOLVERFEASTOL
            L.SCIPsdpiSolverSetRealpar(self.s, par, val)
        self.nvars = 0

    def name(self):
        return self.lib.SCIPsdpiSolverGetSolverName().decode()

    def load_and_solve(self, bp, start=None):
        """start = None or dict(y=[nvars], Z=[(rows, cols, vals)] * (nblocks + 1), X=likewise): the starting point of
        SCIPsdpiSolverLoadAndSolve (sdpisolver.h:160-175; sparse lower triangles, LP block last, indices 2i / 2i+1 for lhs / rhs of
        row i and 2 nlpcons + 2j (+1) for lb (ub) of variable j)"""
        self.nvars = bp.nvars
        args = list(bp.args)
        if start is not None:
            k = _Keep()
            args[30] = k.d(start["y"])
            for base, key in ((31, "Z"), (35, "X")):
                blocks = start[key]
                assert len(blocks) == bp.nblocks + 1
                args[base] = k.i([len(t[0]) for t in blocks])
                args[base + 1] = k.pp([k.i(t[0] if len(t[0]) else [0]) for t in blocks], _ip, _ipp)
                args[base + 2] = k.pp([k.i(t[1] if len(t[1]) else [0]) for t in blocks], _ip, _ipp)
                args[base + 3] = k.pp([k.d(t[2] if len(t[2]) else [0.0]) for t in blocks], _dp, _dpp)
        rc = self.lib.SCIPsdpiSolverLoadAndSolve(self.s, *args)
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverLoadAndSolve returned SCIP_RETCODE {rc}")

    def set_warmstart_pogap(self, gap):
        rc = self.lib.SCIPsdpiSolverSetRealpar(self.s, 12, float(gap))      # SCIP_SDPPAR_WARMSTARTPOGAP
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverSetRealpar(WARMSTARTPOGAP) returned {rc}")

    def preoptimal_sol(self, bp):
        """SCIPsdpiSolverGetPreoptimalPrimalNonzeros + GetPreoptimalSol -> None or (y, [(rows, cols, vals)] per block, LP block last)"""
        nb = bp.nblocks + 1
        cnt = (C.c_int * nb)()
        rc = self.lib.SCIPsdpiSolverGetPreoptimalPrimalNonzeros(self.s, nb, cnt)
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetPreoptimalPrimalNonzeros returned {rc}")
        if cnt[0] == -1:
            return None
        rows = [np.zeros(max(c, 1), dtype=np.int32) for c in cnt]
        cols = [np.zeros(max(c, 1), dtype=np.int32) for c in cnt]
        vals = [np.zeros(max(c, 1)) for c in cnt]
        rp = (_ip * nb)(*[r.ctypes.data_as(_ip) for r in rows])
        cp = (_ip * nb)(*[c.ctypes.data_as(_ip) for c in cols])
        vp = (_dp * nb)(*[v.ctypes.data_as(_dp) for v in vals])
        ok = C.c_uint(0)
        y = np.zeros(self.nvars)
        rc = self.lib.SCIPsdpiSolverGetPreoptimalSol(self.s, C.byref(ok), y.ctypes.data_as(_dp), nb, cnt, C.cast(rp, _ipp), C.cast(cp, _ipp), C.cast(vp, _dpp))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetPreoptimalSol returned {rc}")
        if not ok.value:
            return None
        return y, [(rows[b][:cnt[b]], cols[b][:cnt[b]], vals[b][:cnt[b]]) for b in range(nb)]

    def load_and_solve_with_penalty(self, bp, penaltyparam, withobj, rbound):
        """-> (feasorig, penaltybound) of SCIPsdpiSolverLoadAndSolveWithPenalty (sdpisolver.h:258-322)"""
        self.nvars = bp.nvars
        self.bp = bp
        fo, pb = C.c_uint(0), C.c_uint(0)
        rc = self.lib.SCIPsdpiSolverLoadAndSolveWithPenalty(self.s, float(penaltyparam), int(withobj), int(rbound), *bp.args, C.byref(fo), C.byref(pb))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverLoadAndSolveWithPenalty returned SCIP_RETCODE {rc}")
        return bool(fo.value), bool(pb.value)

    def objval(self):
        v = C.c_double(0)
        rc = self.lib.SCIPsdpiSolverGetObjval(self.s, C.byref(v))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetObjval returned {rc}")
        return v.value

    def primal_matrix_sparse(self, bp):
        """SCIPsdpiSolverGetPrimalNonzeros + GetPrimalMatrix -> list of (rows, cols, vals) per block, LP block last"""
        nb = bp.nblocks + 1
        cnt = (C.c_int * nb)()
        rc = self.lib.SCIPsdpiSolverGetPrimalNonzeros(self.s, nb, cnt)
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetPrimalNonzeros returned {rc}")
        rows = [np.zeros(max(c, 1), dtype=np.int32) for c in cnt]
        cols = [np.zeros(max(c, 1), dtype=np.int32) for c in cnt]
        vals = [np.zeros(max(c, 1)) for c in cnt]
        rp = (_ip * nb)(*[r.ctypes.data_as(_ip) for r in rows])
        cp = (_ip * nb)(*[c.ctypes.data_as(_ip) for c in cols])
        vp = (_dp * nb)(*[v.ctypes.data_as(_dp) for v in vals])
        rc = self.lib.SCIPsdpiSolverGetPrimalMatrix(self.s, nb, cnt, C.cast(rp, _ipp), C.cast(cp, _ipp), C.cast(vp, _dpp))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetPrimalMatrix returned {rc}")
        return [(rows[b][:cnt[b]], cols[b][:cnt[b]], vals[b][:cnt[b]]) for b in range(nb)]

    def primal_matrix_dense(self, bp):
        mats = [np.zeros((n, n)) for n in bp.blocksizes]
        k = _Keep()
        arr = (_dp * max(len(mats), 1))(*[m.ctypes.data_as(_dp) for m in mats])
        ind = k.pp([k.i(np.zeros(n, dtype=np.int32)) for n in bp.blocksizes], _ip, _ipp)
        rc = self.lib.SCIPsdpiSolverGetPrimalSolutionMatrix(self.s, bp.nblocks, k.i(bp.blocksizes if bp.blocksizes else [0]), ind,
                                                           k.i(np.zeros(max(bp.nblocks, 1), dtype=np.int32)),
                                                           k.i(np.zeros(max(bp.nblocks, 1), dtype=np.int32)), C.cast(arr, _dpp))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetPrimalSolutionMatrix returned {rc}")
        return mats

    def primal_inner_products(self, bp):
        """SCIPsdpiSolverGetPrimalInnerProducts (not part of sdpisolver.h): per block the products <A_v, X> of its block
        variables (in the order of the block's variable list), <A_0, X> and a lower bound of min(lambda_min(X), 0), formed on the device"""
        nb = bp.nblocks
        nbv = [int(bp.args[6][b]) for b in range(nb)]
        prods = [np.zeros(max(n, 1)) for n in nbv]
        const, mineig = np.zeros(max(nb, 1)), np.zeros(max(nb, 1))
        arr = (_dp * max(nb, 1))(*[p.ctypes.data_as(_dp) for p in prods])
        F = self.lib.SCIPsdpiSolverGetPrimalInnerProducts
        F.argtypes = [C.c_void_p, C.c_int, _ip, _ipp, _ippp, _ippp, _dppp, _ip, _ipp, _ipp, _dpp, _dpp, _dp, _dp]
        a = bp.args
        rc = F(self.s, nb, a[6], a[13], a[15], a[16], a[17], a[8], a[9], a[10], a[11], C.cast(arr, _dpp),
               const.ctypes.data_as(_dp), mineig.ctypes.data_as(_dp))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetPrimalInnerProducts returned {rc}")
        return [p[:n] for p, n in zip(prods, nbv)], const[:nb], mineig[:nb]

    def bound_multipliers(self):
        lbv, ubv = np.zeros(self.nvars), np.zeros(self.nvars)
        rc = self.lib.SCIPsdpiSolverGetPrimalBoundVars(self.s, lbv.ctypes.data_as(_dp), ubv.ctypes.data_as(_dp))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetPrimalBoundVars returned {rc}")
        return lbv, ubv

    def flag(self, name):
        return bool(getattr(self.lib, "SCIPsdpiSolver" + name)(self.s))

    def dual_sol(self):
        obj = C.c_double(0)
        y = np.zeros(self.nvars)
        rc = self.lib.SCIPsdpiSolverGetDualSol(self.s, C.byref(obj), y.ctypes.data_as(_dp))
        if rc != SCIP_OKAY:
            raise RuntimeError(f"SCIPsdpiSolverGetDualSol returned {rc}")
        return obj.value, y

    def iterations(self):
        it, calls = C.c_int(0), C.c_int(0)
        self.lib.SCIPsdpiSolverGetIterations(self.s, C.byref(it))
        self.lib.SCIPsdpiSolverGetSdpCalls(self.s, C.byref(calls))
        return it.value, calls.value

    def transfer_stats(self):
        """(full uploads, re-solves on the resident problem, host->device bytes) of this solver object since creation"""
        up, pa, by = C.c_int(0), C.c_int(0), C.c_double(0)
        self.lib.SCIPsdpiSolverCudaGetTransferStats.argtypes = [C.c_void_p, _ip, _ip, _dp]
        self.lib.SCIPsdpiSolverCudaGetTransferStats.restype = None
        self.lib.SCIPsdpiSolverCudaGetTransferStats(self.s, C.byref(up), C.byref(pa), C.byref(by))
        return up.value, pa.value, by.value

    def close(self):
        if self.s:
            self.lib.SCIPsdpiSolverFree(C.byref(self.s))
            self.lib.BMSdestroyBufferMemory(C.byref(self.buf))
            self.lib.BMSdestroyBlockMemory(C.byref(self.blk))
            self.s = C.c_void_p()
