"""Oracle bounds of the node workloads of bench.py (scip_sdp_b200.nodesets): every node of every rank of an 8-GPU run is presolved by
Misdp.node_problem and solved by the CPU oracle (settings ladder like frontier.branch_and_bound: FAST, then STABLE, then the penalty
formulation), in worker processes.  Codes are scanned in increasing order until 8 x (nodes per GPU) of them have been solved to optimality (status 0; the others —
1 infeasibility certificate, 2 infeasible by presolve, 3 all variables fixed, 4 not solved — never enter the frontier).
Output: frontier_bounds.npz with, per workload, `<name>_codes` and `<name>_bound`.
   python tests/golden/make_frontier_bounds.py [workload ...]"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frontier_bounds.npz")
KW = dict(gaptol=1e-5, feastol=1e-5)


def solve_codes(args):
    name, codes, threads = args
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from scip_sdp_b200 import abi, frontier, nodesets
    lib = abi.Lib(abi.ORACLE_LIB)
    lib.lib.sdporacle_set_threads(threads)
    cpu = abi.Solver(lib)
    M = nodesets.WORKLOADS[name][0]()
    lbs, ubs = nodesets.node_bounds(M, codes)
    out = []
    for lb, ub in zip(lbs, ubs):
        st, fp, info = M.node_problem_fast(lb, ub, feastol=KW["feastol"])
        if st == "infeasible":
            out.append((2, 0.0)); continue
        if st == "allfixed":
            out.append((3, float(info["fixedobj"]))); continue
        r = cpu.solve(fp, fetch=False, **KW)
        if r["phase_name"] not in ("pdOPT", "pFEAS_dINF", "dINF"):
            r = cpu.solve(fp, fetch=False, setting=3, **KW)
        if r["phase_name"] not in ("pdOPT", "pFEAS_dINF", "dINF"):
            r = frontier._penalty_ladder(cpu, M, info["lb"], info["ub"], r, KW)
        if r["phase_name"] == "pdOPT":
            out.append((0, float(r["dobj"] + info["fixedobj"])))
        elif r["phase_name"] in ("pFEAS_dINF", "dINF"):
            out.append((1, 0.0))
        else:
            out.append((4, float(r["dobj"] + info["fixedobj"])))
    return out


def main():
    from scip_sdp_b200 import nodesets
    names = sys.argv[1:] or list(nodesets.WORKLOADS)
    data = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    cores = os.cpu_count() or 1
    for name in names:
        if name == "MkP-120":            # root relaxation only (see nodesets.WORKLOADS), replicated for the 8 ranks
            out = solve_codes((name, [-1], cores))
            assert out[0][0] == 0, out
            data[name + "_codes"], data[name + "_bound"] = np.full(nodesets.MAX_RANKS, -1, dtype=np.int64), np.full(nodesets.MAX_RANKS, out[0][1])
            print(f"{name}: root relaxation {out[0][1]!r}", flush=True)
            np.savez_compressed(OUT, **data)
            continue
        per = nodesets.WORKLOADS[name][1]
        want = per * nodesets.MAX_RANKS
        small = nodesets.WORKLOADS[name][2] == "nodes"
        workers = cores if small else 2
        threads = 1 if small else max(1, cores // 2)
        keep_codes, keep_bounds, counts, nextcode = [], [], np.zeros(5, dtype=int), 0
        t0 = time.time()
        with ProcessPoolExecutor(workers) as ex:
            while len(keep_codes) < want:
                n = max(workers, min(4 * (want - len(keep_codes)), 4096) if small else 2 * workers)
                codes = list(range(nextcode, nextcode + n)); nextcode += n
                chunks = [codes[i::workers] for i in range(workers)]
                parts = list(ex.map(solve_codes, [(name, c, threads) for c in chunks]))
                got = {}
                for c, part in zip(chunks, parts):
                    got.update(zip(c, part))
                for code in codes:
                    s, b = got[code]
                    counts[s] += 1
                    if s == 0 and len(keep_codes) < want:
                        keep_codes.append(code); keep_bounds.append(b)
        data[name + "_codes"], data[name + "_bound"] = np.array(keep_codes, dtype=np.int64), np.array(keep_bounds)
        print(f"{name}: {want} solved nodes out of the first {nextcode} codes in {time.time() - t0:.1f} s, status counts {counts.tolist()}", flush=True)
        np.savez_compressed(OUT, **data)


if __name__ == "__main__":
    main()
