"""CPU suite, part 5: the host half of sdpcuda_solve_batch.  sdpcuda_debug_pack_node (no device needed) packs one node exactly like
the batch call; this test reads the image through the kernel descriptor the way the device code does (assemble / applyA / lprows /
lpcols of csrc/ipm_small.cu, restated in numpy) and compares with dense operators built straight from the problem."""
import ctypes as C
import os

import numpy as np
import pytest

from scip_sdp_b200 import abi, generators, misdp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MAXB, MAXG = 16, 8
P_ = C.c_uint64          # device pointers are read back as plain addresses


class SmallBlock(C.Structure):
    _fields_ = [("n", C.c_int), ("ld", C.c_int), ("off", C.c_longlong), ("lzoff", C.c_longlong)]


class SmallArgs(C.Structure):       # mirror of sdpk::SmallArgs (csrc/ipm_small.cuh); its size is checked against the library's
    _fields_ = ([(k, C.c_int) for k in "m nb nlp N ldm npos cnnz ndense ngroups maxiter setting verbose".split()]
                + [("arena", C.c_longlong), ("blk", SmallBlock * MAXB)]
                + [(k, P_) for k in "varbeg erow ecol eld eoff eval cls posbeg pos mirror posvar posval posc cpos cmirror cval "
                                     "lpbeg lpind lpval lprhs colbeg colrow colval b denselist Adense".split()]
                + [("gblk", C.c_int * MAXG), ("gfirst", C.c_int * MAXG), ("gcount", C.c_int * MAXG), ("gaoff", C.c_longlong * MAXG)]
                + [(k, P_) for k in "X S Sinv L Linv LX LXinv dX dS dXa dSa K T1 T2 Rd y dy g rp AX DTx tm1 tm2 "
                                     "x s dx ds dxa dsa klp rdlp Dy Ddy M Mfac Hd Ud lz".split()]
                + [(k, C.c_double) for k in "gaptol feastol absgaptol objlimit normb normC normCsdp2 gammabase".split()]
                + [("out", P_), ("selfinit", C.c_int), ("stage_doubles", C.c_longlong), ("workbase", P_), ("copyback", C.c_int), ("adense_total", C.c_longlong), ("mpk_off", C.c_longlong), ("xil", C.c_double), ("etal", C.c_double),
                   ("xi", C.c_double * MAXB), ("eta", C.c_double * MAXB)])


IMG_BASE, WORK_BASE, Y_BASE = 0x10000000, 0x40000000, 0x70000000


@pytest.fixture(scope="module")
def lib():
    L = abi.Lib(abi.PRODUCT_LIB)          # loads without a GPU; the hook makes no CUDA call
    L.lib.sdpcuda_debug_pack_node.argtypes = [C.POINTER(abi.Problem), C.POINTER(abi.Params), C.c_ulonglong, C.c_ulonglong, C.c_ulonglong,
                                              C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                              C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    return L


def pack(lib, fp, **kw):
    par = lib.default_params(**kw)
    st = fp.struct()
    nimg, nwork, ndesc, fits = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_int(-1)
    args = (C.byref(st), C.byref(par), IMG_BASE, WORK_BASE, Y_BASE)
    assert lib.lib.sdpcuda_debug_pack_node(*args, None, 0, C.byref(nimg), C.byref(nwork), None, 0, C.byref(ndesc), C.byref(fits)) == 0
    if not fits.value:
        return None
    assert ndesc.value == C.sizeof(SmallArgs), "tests/test_batch_pack.py:SmallArgs is out of step with csrc/ipm_small.cuh"
    img = np.zeros(nimg.value, dtype=np.uint8)
    a = SmallArgs()
    assert lib.lib.sdpcuda_debug_pack_node(*args, img.ctypes.data, img.size, C.byref(nimg), C.byref(nwork), C.byref(a), C.sizeof(a),
                                           C.byref(ndesc), C.byref(fits)) == 0
    return a, img, nwork.value


def view(img, addr, dtype, count):
    off = addr - IMG_BASE
    assert 0 <= off and off % 16 == 0 and off + count * np.dtype(dtype).itemsize <= img.size, "array outside the image"
    return img[off:off + count * np.dtype(dtype).itemsize].view(dtype)


def _cases():
    for f in ("example_small.dat-s", "example_TT.dat-s.gz", "example_CLS.dat-s.gz", "example_MkP.dat-s.gz"):
        yield f, (lambda f=f: misdp.read_sdpa(os.path.join(GOLDEN, f)).rows_to_bounds().flatten()[0])
    yield "maxcut-40 (no LP block)", lambda: generators.maxcut(40, 0.2, seed=3).flatten()[0]
    yield "cls-12 (dense matrices)", lambda: generators.cls(12, 9, 3, seed=5).flatten()[0]
    yield "mkp-10", lambda: generators.mkp(10, seed=6).flatten()[0]


@pytest.mark.parametrize("name,make", list(_cases()), ids=[n for n, _ in _cases()])
def test_packed_node_reproduces_the_operators(lib, name, make):
    fp = make()
    packed = pack(lib, fp, gaptol=1e-5, feastol=1e-5)
    assert packed is not None, "all test cases are inside the single-CTA limits"
    a, img, nwork = packed
    m, nb, nlp = fp.m, fp.nblocks, fp.nlp
    assert (a.m, a.nb, a.nlp, a.cnnz, a.selfinit) == (m, nb, nlp, len(fp.cval), 1)
    assert a.ldm >= m and a.ldm % 4 == 0 and a.N == nlp + int(fp.blocksizes.sum())
    # ---- block table: 128-byte aligned blocks that do not overlap ----
    end = 0
    for k in range(nb):
        b = a.blk[k]
        assert b.n == fp.blocksizes[k] and b.ld >= b.n and b.ld % 4 == 0 and b.off % 16 == 0 and b.off >= end
        end = b.off + b.ld * b.n
    assert end <= a.arena
    nnz = int(fp.varbeg[m])
    varbeg, erow, ecol = view(img, a.varbeg, np.int32, m + 1), view(img, a.erow, np.int32, nnz), view(img, a.ecol, np.int32, nnz)
    eld, eoff, eval_ = view(img, a.eld, np.int32, nnz), view(img, a.eoff, np.int64, nnz), view(img, a.eval, np.float64, nnz)
    assert np.array_equal(varbeg, fp.varbeg)
    rng = np.random.default_rng(1)
    # ---- assemble: T = sum_j v_j A_j - C through the position-major lists (ipm_small.cu: assemble) ----
    posbeg = view(img, a.posbeg, np.int32, a.npos + 1)
    pos, mirror = view(img, a.pos, np.int64, a.npos), view(img, a.mirror, np.int64, a.npos)
    posvar, posval = view(img, a.posvar, np.int32, posbeg[-1]), view(img, a.posval, np.float64, posbeg[-1])
    posc = view(img, a.posc, np.float64, a.npos)
    v = rng.standard_normal(m)
    T = np.zeros(a.arena)
    for p in range(a.npos):
        s = -posc[p] + sum(v[posvar[e]] * posval[e] for e in range(posbeg[p], posbeg[p + 1]))
        T[pos[p]] = s
        T[mirror[p]] = s
    Cd = fp.dense_C()
    want = [sum(v[j] * fp.dense_A(j)[k] for j in range(m)) - Cd[k] for k in range(nb)]
    for k in range(nb):
        b = a.blk[k]
        got = T[b.off:b.off + b.ld * b.n].reshape(b.n, b.ld)[:, :b.n].T        # column-major with leading dimension ld
        assert np.allclose(got, want[k], atol=1e-12), f"block {k}"
    # ---- applyA: out_j = A_j . X through the entry lists (ipm_small.cu: applyA) ----
    Xs = [rng.standard_normal((n, n)) for n in fp.blocksizes]
    Xs = [x + x.T for x in Xs]
    Xar = np.zeros(a.arena)
    for k in range(nb):
        b = a.blk[k]
        Xar[b.off:b.off + b.ld * b.n].reshape(b.n, b.ld)[:, :b.n] = Xs[k].T
    out = np.zeros(m)
    for j in range(m):
        for e in range(varbeg[j], varbeg[j + 1]):
            r, c, ld, off = erow[e], ecol[e], eld[e], eoff[e]
            u = Xar[off + c * ld + r] + (Xar[off + r * ld + c] if r != c else 0.0)
            out[j] += eval_[e] * u
    want = np.array([sum(float((fp.dense_A(j)[k] * Xs[k]).sum()) for k in range(nb)) for j in range(m)])
    assert np.allclose(out, want, atol=1e-10)
    # ---- constant entries (const_dots) ----
    cpos, cmir, cval = view(img, a.cpos, np.int64, a.cnnz), view(img, a.cmirror, np.int64, a.cnnz), view(img, a.cval, np.float64, a.cnnz)
    cx = sum(cval[e] * (Xar[cpos[e]] + (Xar[cmir[e]] if cpos[e] != cmir[e] else 0.0)) for e in range(a.cnnz))
    assert np.isclose(cx, sum(float((Cd[k] * Xs[k]).sum()) for k in range(nb)), atol=1e-10)
    # ---- LP block: rows (CSR) and columns (CSC) ----
    D = fp.dense_D()
    lpbeg = view(img, a.lpbeg, np.int32, nlp + 1)
    lnz = int(lpbeg[-1])
    lpind, lpval = view(img, a.lpind, np.int32, lnz), view(img, a.lpval, np.float64, lnz)
    colbeg = view(img, a.colbeg, np.int32, m + 1)
    colrow, colval = view(img, a.colrow, np.int32, lnz), view(img, a.colval, np.float64, lnz)
    Dy = np.array([sum(lpval[p] * v[lpind[p]] for p in range(lpbeg[l], lpbeg[l + 1])) for l in range(nlp)])
    xv = rng.standard_normal(nlp)
    DTx = np.array([sum(colval[p] * xv[colrow[p]] for p in range(colbeg[j], colbeg[j + 1])) for j in range(m)])
    assert np.allclose(Dy, D @ v, atol=1e-12) and np.allclose(DTx, D.T @ xv, atol=1e-12)
    assert np.array_equal(view(img, a.lprhs, np.float64, nlp), fp.lprhs) and np.array_equal(view(img, a.b, np.float64, m), fp.obj)
    # ---- variable classes and dense groups ----
    cls = view(img, a.cls, np.int32, m)
    dl = view(img, a.denselist, np.int32, a.ndense)
    assert sorted(dl) == sorted(np.flatnonzero(cls == 2)) and sum(a.gcount[g] for g in range(a.ngroups)) == a.ndense
    tot = 0
    for g in range(a.ngroups):
        b = a.blk[a.gblk[g]]
        assert a.gaoff[g] == tot and a.gfirst[g] == sum(a.gcount[q] for q in range(g))
        for d in range(a.gcount[g]):
            j = dl[a.gfirst[g] + d]
            assert all(eoff[e] == b.off for e in range(varbeg[j], varbeg[j + 1]))
        tot += a.gcount[g] * b.ld * b.n
    assert tot == a.adense_total
    # ---- work space: every array inside [0, nwork), 128-byte aligned, pairwise disjoint ----
    ar, mv, lv, mm = a.arena, m + 1, nlp + 1, a.ldm * m
    hd = max((a.gcount[g] for g in range(a.ngroups)), default=0) * max((a.blk[a.gblk[g]].ld * a.blk[a.gblk[g]].n for g in range(a.ngroups)), default=0)
    lz = sum(34 * a.blk[k].n for k in range(nb)) + 16
    need = dict(X=ar, S=ar, Sinv=ar, L=ar, Linv=ar, LX=ar, LXinv=ar, dX=ar, dS=ar, dXa=ar, dSa=ar, K=ar, T1=ar, T2=ar, Rd=ar,
                dy=mv, g=mv, rp=mv, AX=mv, DTx=mv, tm1=mv, tm2=mv, x=lv, s=lv, dx=lv, ds=lv, dxa=lv, dsa=lv, klp=lv, rdlp=lv, Dy=lv, Ddy=lv,
                M=mm, Mfac=mm, Adense=a.adense_total, Hd=hd, Ud=hd, lz=lz)
    spans = []
    for k, n in need.items():
        addr = getattr(a, k)
        assert addr % 128 == 0 and (addr - WORK_BASE) % 8 == 0, k
        lo = (addr - WORK_BASE) // 8
        assert 0 <= lo and lo + n <= nwork, k
        if n > 0:
            spans.append((lo, lo + n, k))
    spans.sort()
    for (l0, h0, k0), (l1, h1, k1) in zip(spans, spans[1:]):
        assert h0 <= l1, f"{k0} overlaps {k1}"
    assert a.y == Y_BASE and a.out == 0
    # ---- cold start and tolerances ----
    assert all(a.xi[k] >= 10.0 and a.eta[k] >= 10.0 for k in range(nb)) and a.xil >= 10.0 and a.etal >= 10.0
    assert (a.gaptol, a.feastol, a.maxiter, a.gammabase) == (1e-5, 1e-5, 100, 0.9)
    assert np.isclose(a.normb, np.linalg.norm(fp.obj)) and np.isclose(a.normC ** 2, sum(float((c * c).sum()) for c in Cd) + float(fp.lprhs @ fp.lprhs))


def test_relaxations_outside_the_single_cta_limits_do_not_fit(lib):
    assert pack(lib, generators.maxcut(96, 0.1, seed=7).flatten()[0]) is None          # block of order 96 > 64
    assert pack(lib, generators.mkp(24, seed=12).flatten()[0]) is None                 # m = 276 > 256
    assert pack(lib, generators.mkp(20, seed=12).flatten()[0]) is not None             # m = 190
