"""Eigen path of the cut separation (cons_sdp.c:1612-1797): all SDP blocks of a separation round through
SCIPlapackComputeEigenvectorsNegativeBatch (one device call per distinct order: one H2D, one launch of the batched Jacobi kernel, one
D2H) against the per-constraint calls of our library and against the REFERENCE's own lapack_interface.c (DSYEVR, one thread, as SCIP
calls it; oracle/_ref/liblapack_ref.so).  Prints matrices/s and the achieved GB/s on the algorithmic bytes 16 n^2 + 8 n per matrix.
   python tools/eig_bench.py [matrices per round]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def load(path):
    L = C.CDLL(path, mode=C.RTLD_LOCAL)
    L.BMScreateBufferMemory.restype = C.c_void_p
    L.BMScreateBufferMemory.argtypes = [C.c_double, C.c_int, C.c_uint]
    L.SCIPlapackComputeEigenvectorsNegative.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.POINTER(C.c_int), _dp, _dp]
    return L, C.c_void_p(L.BMScreateBufferMemory(1.2, 4, 0))


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ours, bo = load(os.path.join(ROOT, "scip-sdp_b200", "lib", "libsdpisolver_cuda.so"))
    ref, br = load(os.path.join(ROOT, "oracle", "_ref", "liblapack_ref.so"))
    ours.SCIPlapackComputeEigenvectorsNegativeBatch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(_dp), C.c_double,
                                                                 C.POINTER(C.c_int), C.POINTER(_dp), C.POINTER(_dp)]
    rng = np.random.default_rng(1)
    p = lambda a: a.ctypes.data_as(_dp)          # noqa: E731
    print(f"{'order':>6s} {'matrices':>9s} {'reference lapack_interface.c':>30s} {'ours, one call per matrix':>28s} {'ours, one batch call':>24s} {'GB/s batch':>11s}")
    for n in (10, 15, 43, 64, 100):
        mats = []
        for _ in range(count):
            A = rng.standard_normal((n, n)); mats.append(A + A.T)
        w, V, cnt = np.zeros(n), np.zeros(n * n), C.c_int(0)
        t = {}
        for key, (L, b) in (("ref", (ref, br)), ("single", (ours, bo))):
            for rep in range(2):                                   # the second pass is timed (first call creates the device handle)
                t0 = time.perf_counter()
                for A in mats:
                    Ac = A.copy()
                    assert L.SCIPlapackComputeEigenvectorsNegative(b, n, p(Ac), 1e-6, C.byref(cnt), p(w), p(V)) == 1
                t[key] = time.perf_counter() - t0
        sizes = (C.c_int * count)(*([n] * count))
        cp = [A.copy() for A in mats]
        ws, Vs, cnts = [np.zeros(n) for _ in mats], [np.zeros(n * n) for _ in mats], (C.c_int * count)()
        args = (bo, count, sizes, (_dp * count)(*[p(a) for a in cp]), 1e-6, cnts, (_dp * count)(*[p(a) for a in ws]), (_dp * count)(*[p(a) for a in Vs]))
        for rep in range(2):
            t0 = time.perf_counter()
            assert ours.SCIPlapackComputeEigenvectorsNegativeBatch(*args) == 1
            t["batch"] = time.perf_counter() - t0
        gbs = count * (16.0 * n * n + 8.0 * n) / t["batch"] / 1e9
        print(f"{n:6d} {count:9d} {count / t['ref']:22.0f} /s {count / t['single']:24.0f} /s {count / t['batch']:20.0f} /s {gbs:11.3f}")


if __name__ == "__main__":
    main()
