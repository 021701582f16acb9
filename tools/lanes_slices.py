"""mid-size node workloads on the slices of several ranks, one process: separates the effect of the slice from that of the launch"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from scip_sdp_b200 import abi, nodesets
lib = abi.Lib(abi.PRODUCT_LIB); g = abi.Solver(lib, device=0); t = nodesets.golden()
for name in sys.argv[1:]:
    for rank in (0, 1, 2, 3):
        for rep in range(3):
            r = bench.gpu_node_workload(g, lib, name, rank, t, 1)
            print(name, "slice of rank", rank, "rep", rep, "ms per frontier", round(1e3 * r["wall_s"], 2), "device one-at-a-time ms", round(r["device_ms"], 2), "launches", r["launches"], flush=True)
