import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scip_sdp_b200 import abi
S = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
print(S.time_kernel(8, 0, 1)); print(S.time_kernel(9, 0, 1))
for rep in range(3):
    print("diag64", S.time_kernel(2, 64, 50))
print("potrf 128", S.time_kernel(2, 128, 20), "potrf 256", S.time_kernel(2, 256, 20))
