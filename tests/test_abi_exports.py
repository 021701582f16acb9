"""CPU suite, part 2: the C-ABI libraries load and export every symbol the headers declare (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scip-sdp_b200")


def _exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def _declared(header):
    txt = open(header).read()
    return sorted(set(re.findall(r"\b(sdpcuda_[a-z0-9_]+)\s*\(", txt)))


def test_libsdpcuda_exports_every_declared_symbol():
    lib = os.path.join(PKG, "lib", "libsdpcuda.so")
    assert os.path.exists(lib), "run __graft_entry__.build() first"
    names = _declared(os.path.join(ROOT, "include", "sdpcuda.h"))
    assert len(names) >= 15
    missing = [n for n in names if n not in _exported(lib)]
    assert not missing, missing
    ctypes.CDLL(lib)      # loads without a GPU (no CUDA call happens at load time)


def test_oracle_exports_the_same_abi():
    lib = os.path.join(ROOT, "oracle", "liboracle_sdp.so")
    names = _declared(os.path.join(ROOT, "include", "sdpcuda.h"))
    missing = [n for n in names if n not in _exported(lib)]
    assert not missing, missing


# the 53 functions of the reference's src/sdpi/sdpisolver.h:79-724 (SURVEY.md section 8b checklist)
SDPISOLVER = """GetSolverName GetSolverDesc GetSolverPointer GetDefaultSdpiSolverNpenaltyIncreases DoesWarmstartNeedPrimal Create Free
IncreaseCounter ResetCounter LoadAndSolve LoadAndSolveWithPenalty WasSolved FeasibilityKnown GetSolFeasibility IsPrimalUnbounded
IsPrimalInfeasible IsPrimalFeasible IsDualUnbounded IsDualInfeasible IsDualFeasible IsConverged IsObjlimExc IsIterlimExc IsTimelimExc
GetInternalStatus IsOptimal IsAcceptable IgnoreInstability GetObjval GetDualSol GetPreoptimalPrimalNonzeros GetPreoptimalSol
GetPrimalBoundVars GetPrimalLPSides GetPrimalNonzeros GetPrimalMatrix GetPrimalSolutionMatrix GetMaxPrimalEntry GetTime GetIterations
GetSdpCalls SettingsUsed Infinity IsInfinity GetRealpar SetRealpar GetIntpar SetIntpar ComputeLambdastar ComputePenaltyparam
ComputeMaxPenaltyparam ReadSDP WriteSDP""".split()


def test_binding_exports_all_53_sdpisolver_functions():
    lib = os.path.join(PKG, "lib", "libsdpisolver_cuda.so")
    assert os.path.exists(lib), "run __graft_entry__.build() first"
    assert len(SDPISOLVER) == 53
    exp = _exported(lib)
    missing = ["SCIPsdpiSolver" + n for n in SDPISOLVER if "SCIPsdpiSolver" + n not in exp]
    assert not missing, missing
    for n in ["SCIPlapackComputeIthEigenvalue", "SCIPlapackComputeEigenvectorsNegative", "SCIPlapackComputeEigenvectorDecomposition"]:
        assert n in exp, n


def test_product_library_does_not_link_the_oracle():
    for name in ["libsdpcuda.so", "libsdpisolver_cuda.so"]:
        out = subprocess.run(["ldd", os.path.join(PKG, "lib", name)], capture_output=True, text=True).stdout
        assert "oracle" not in out and "openblas" not in out, out


def test_checker_libraries_bind_their_own_backend():
    """libsdpi_oracle.so and libsdpi_cuda.so both reference sdpcuda_* symbols: each must resolve them inside its own
    dependency tree (RTLD_LOCAL loads), never through whichever implementation happened to be loaded first"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from harness import sdpi_ref
    from scip_sdp_b200 import abi
    if not os.path.exists(sdpi_ref.LIB_ORACLE):
        pytest.skip("oracle/_ref not built")
    ctypes.CDLL(os.path.join(PKG, "lib", "libsdpcuda.so"), mode=ctypes.RTLD_LOCAL)      # product library loaded first on purpose
    lib = sdpi_ref.SdpiLib(sdpi_ref.LIB_ORACLE)
    name = ctypes.CDLL(sdpi_ref.LIB_ORACLE, mode=ctypes.RTLD_LOCAL).sdpcuda_backend_name
    name.restype = ctypes.c_char_p
    assert name() == b"cpu-oracle"
    # and a solve through the reference's sdpi.c really runs (no CUDA device exists in the CPU test environment)
    from golden.checksdpi_cases import CASES
    from harness import checksdpi_port
    os.environ.setdefault("SHIM_QUIET", "1")
    checksdpi_port.run_case(lib, CASES["test1"], "test1")
