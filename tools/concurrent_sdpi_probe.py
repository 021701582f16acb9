"""SCIP's concurrent mode on one GPU, through the REFERENCE's own sdpi.c: T host threads, each with its own SCIP_SDPI object
(relax_sdp.c:5387 creates one per SCIP instance; settings/scip-[1-8].set run 1..8 instances side by side), every thread walking
its share of a committed frontier the way relax_sdp.c does: SCIPsdpiChgBounds -> SCIPsdpiSolve -> SCIPsdpiGetDualSol.  Below it
the binding (sdpi/sdpisolver_cuda.c) gives every object its own device handle and stream, so the one-launch kernels of different
threads share the GPU without any queue in between.  Prints nodes/s per thread count and the worst difference to the committed
oracle bound.  Test infrastructure (uses oracle/_ref/libsdpi_cuda.so = unmodified sdpi.c + the binding)."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import sdpi_ref                    # noqa: E402
from scip_sdp_b200 import nodesets              # noqa: E402

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["example_TT", "example_MkP", "example_CLS"]
NN = int(sys.argv[2]) if len(sys.argv) > 2 else 128
threads = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8, 16, 32]
lib = sdpi_ref.SdpiLib(sdpi_ref.LIB_ORACLE if os.environ.get("SDPI_LIB") == "oracle" else sdpi_ref.LIB_CUDA)
table = nodesets.golden()
for name in names:
    M = nodesets.WORKLOADS[name][0]()
    codes, want = nodesets.frontier_of_rank(name, 0, table=table)
    codes, want = codes[:NN], want[:NN]
    lbs, ubs = nodesets.node_bounds(M, codes)
    idx = np.arange(M.nvars, dtype=np.int32)
    for T in threads:
        objs = [sdpi_ref.Sdpi(lib, gaptol=1e-5, sdpsolverfeastol=1e-5, feastol=1e-5) for _ in range(T)]
        for s in objs:
            s.load_model(M)
            s.solve()                                   # warm-up: device buffers of the handle, kernel attributes
        got = np.full(len(codes), np.nan)
        okflag = np.zeros(len(codes), dtype=bool)

        def work(t, s):
            for k in range(t, len(codes), T):
                s.chg_bounds(idx, lbs[k], ubs[k])
                s.solve()
                okflag[k] = s.flag("IsOptimal")
                got[k] = s.dual_sol()[0]
        th = [threading.Thread(target=work, args=(t, s)) for t, s in enumerate(objs)]
        t0 = time.perf_counter()
        for x in th: x.start()
        for x in th: x.join()
        dt = time.perf_counter() - t0
        rel = np.abs(got - want) / np.maximum(1.0, np.abs(want))
        print(f"{name} threads {T:3d}: {len(codes) / dt:9.1f} nodes/s  ({1e3 * dt * T / len(codes):6.2f} ms per node per thread), "
              f"optimal {int(okflag.sum())}/{len(codes)}, max rel diff to the oracle bound {np.nanmax(rel):.2e}", flush=True)
        for s in objs: s.close()
