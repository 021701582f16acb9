"""Regenerates the instance fixtures in tests/golden from the reference tree (run in the build container only):
copies of instances/example_{small,inf}.dat-s and example_{TT,CLS,MkP}.dat-s.gz (the BASELINE.json configs) and of the
further check/testset/short.test instances the B&B harness can decide (CBF files without rank-1 constraints, tightenmatrices).
They are data files of the reference, kept verbatim so that the GPU box (no /root/reference) can run the parity tests."""
import os
import shutil

SRC = "/root/reference/instances"
DST = os.path.dirname(os.path.abspath(__file__))
for name in ["example_small.dat-s", "example_inf.dat-s", "example_TT.dat-s.gz", "example_CLS.dat-s.gz", "example_MkP.dat-s.gz",
             "example_small_cbf.cbf", "example_cbf_primal.cbf", "example_cbf_mix.cbf", "example_cbf_dual.cbf", "example_multaggr.cbf",
             "example_diagzeroimpl.cbf", "example_tightenmatrices.dat-s", "example_small_ind.dat-s"]:
    shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    print("copied", name)
