import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi
S = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
print(S.time_kernel(8, 0, 1)); print(S.time_kernel(9, 0, 1))
for rep in range(3):
    print("diag64", S.time_kernel(2, 64, 50))
print("potrf 128", S.time_kernel(2, 128, 20), "potrf 256", S.time_kernel(2, 256, 20))
for leaf in ("64", "128"):
    os.environ["SDPCUDA_LEAF"] = leaf
    for n in (64, 128, 256, 512, 1024, 1501, 2000, 4096, 7140):
        ms, work = S.time_kernel(2, n, 10)
        ms3, work3 = S.time_kernel(3, n, 10)
        print(f"leaf={leaf} n={n}: potrf+inverse {ms * 1e3:.1f} us ({work / ms * 1e-9:.2f} TF)   potrf {ms3 * 1e3:.1f} us ({work3 / ms3 * 1e-9:.2f} TF)", flush=True)
