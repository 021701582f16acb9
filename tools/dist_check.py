"""Sharded Schur complement (one SDP over several GPUs) under torchrun: every rank solves the same relaxations with the clique
joined and checks that all ranks end bit-identical and equal to an unsharded solve of the same handle.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi, frontier, generators  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ["SDPCUDA_DEVICE"] = str(local)
saved = os.dup(1); os.dup2(2, 1)                    # NCCL prints its banner on stdout
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier()
os.dup2(saved, 1)
S = abi.Solver(abi.Lib(abi.PRODUCT_LIB), local)
cases = {"maxcut-300": lambda: generators.maxcut(300, 0.05, seed=21), "mkp-30": lambda: generators.mkp(30, seed=22),
         "truss-80": lambda: generators.truss(5, 5, 80, seed=23), "cls-60": lambda: generators.cls(60, 30, 5, seed=24)}
ref = {}
for name, make in cases.items():
    fp, _ = make().flatten()
    os.environ["SDPCUDA_PATH"] = "m"
    ref[name] = S.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)
frontier.shard_one_sdp(S, dist, device=f"cuda:{local}")
ok = True
for name, make in cases.items():
    fp, _ = make().flatten()
    r = S.solve(fp, gaptol=1e-7, feastol=1e-7, fetch=False)
    y = S.get_y()
    t = torch.tensor(np.concatenate([[r["dobj"], r["pobj"], float(r["iterations"])], y]), dtype=torch.float64, device=f"cuda:{local}")
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    same_ranks = bool(torch.equal(tmax, tmin))
    # bit-identical to the unsharded solve without LP rows.  The LP block of M is accumulated with atomics whose order changes
    # from run to run in the last bits (on one GPU as well), so instances with LP rows agree to the solver tolerance only
    if fp.nlp == 0:
        same_single = (r["dobj"] == ref[name]["dobj"] and r["iterations"] == ref[name]["iterations"])
    else:
        same_single = abs(r["dobj"] - ref[name]["dobj"]) <= 1e-6 * max(1.0, abs(ref[name]["dobj"]))
    if rank == 0:
        print(f"{name}: m={fp.m} {r['phase_name']} it={r['iterations']} dobj={r['dobj']:.12g} ranks identical={same_ranks} equals unsharded={same_single}", flush=True)
    ok = ok and same_ranks and same_single and r["phase_name"] == "pdOPT"
S.dist_finalize()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)
