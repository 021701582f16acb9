import os,sys
sys.path.insert(0,os.getcwd())
from scip_sdp_b200 import abi
S = abi.Solver(abi.Lib(abi.PRODUCT_LIB), 0)
print(S.time_kernel(3, 64, 3))
