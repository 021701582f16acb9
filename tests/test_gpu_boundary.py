"""GPU suite, part 5: the same boundary checks on the product library lib/libsdpisolver_cuda.so (binding + GPU checker + CUDA)."""
import pytest

from harness import boundary_cases
from scip_sdp_b200 import sdpisolver_host

pytestmark = pytest.mark.gpu


def test_penalty_formulation_call_patterns_on_gpu():
    boundary_cases.run_penalty_patterns(sdpisolver_host.BINDING_LIB)


def test_primal_matrix_getters_are_consistent_on_gpu():
    boundary_cases.run_primal_getters(sdpisolver_host.BINDING_LIB)


def test_warmstart_and_preoptimal_solution_on_gpu():
    boundary_cases.run_warmstart_and_preoptimal(sdpisolver_host.BINDING_LIB)


def test_resolves_of_a_node_reuse_the_resident_problem_on_gpu():
    """row f4: host->device traffic of the re-solves is the two patched vectors only"""
    boundary_cases.run_resident_resolves(sdpisolver_host.BINDING_LIB, device=True)


def test_conflict_cut_reductions_on_the_resident_primal_solution_on_gpu():
    boundary_cases.run_primal_inner_products(sdpisolver_host.BINDING_LIB)


def test_conflict_cut_reductions_after_a_packed_single_solve_on_gpu(monkeypatch):
    """the same reductions when the relaxation went through the packed one-launch path (the default with several solver objects
    alive): X then lives in the work space of the batch kernel"""
    monkeypatch.setenv("SDPCUDA_PACKED_SOLVE", "1")
    boundary_cases.run_primal_inner_products(sdpisolver_host.BINDING_LIB)


def test_concurrent_solver_objects_through_the_reference_sdpi_on_gpu():
    """SCIP's concurrent mode (settings/scip-[1-8].set): several threads, each with its own SCIP_SDPI object (relax_sdp.c:5387),
    walk a committed frontier through the reference's unmodified sdpi.c — SCIPsdpiChgBounds, SCIPsdpiSolve, SCIPsdpiGetDualSol.
    Every object owns a device handle and stream; with >= 3 objects alive the binding's solves take the packed one-launch path.
    Every node must end optimal and within 1e-5 of the committed oracle bound, and the threaded walk must reproduce the bounds of
    a walk by a single thread with the same number of objects alive (same kernels => same numbers)."""
    import os
    import threading
    import numpy as np
    from harness import sdpi_ref
    from scip_sdp_b200 import nodesets
    if not os.path.exists(sdpi_ref.LIB_CUDA):
        pytest.skip("oracle/_ref/libsdpi_cuda.so (reference sdpi.c + the binding) not built")
    lib = sdpi_ref.SdpiLib(sdpi_ref.LIB_CUDA)
    table = nodesets.golden()
    T = 6
    for name, nn in (("example_TT", 48), ("example_MkP", 24), ("example_CLS", 12)):
        M = nodesets.WORKLOADS[name][0]()
        codes, want = nodesets.frontier_of_rank(name, 0, table=table)
        codes, want = codes[:nn], want[:nn]
        lbs, ubs = nodesets.node_bounds(M, codes)
        idx = np.arange(M.nvars, dtype=np.int32)
        objs = [sdpi_ref.Sdpi(lib, gaptol=1e-5, sdpsolverfeastol=1e-5, feastol=1e-5) for _ in range(T)]
        try:
            for s in objs:
                s.load_model(M)
            got = np.full((2, nn), np.nan)
            opt = np.zeros((2, nn), dtype=bool)

            def walk(run, s, ks):
                for k in ks:
                    s.chg_bounds(idx, lbs[k], ubs[k])
                    s.solve()
                    opt[run, k] = s.flag("IsOptimal")
                    got[run, k] = s.dual_sol()[0]
            th = [threading.Thread(target=walk, args=(0, s, range(t, nn, T))) for t, s in enumerate(objs)]
            for x in th:
                x.start()
            for x in th:
                x.join()
            walk(1, objs[0], range(nn))
        finally:
            for s in objs:
                s.close()
        assert opt.all(), (name, np.flatnonzero(~opt.all(axis=0)))
        rel = np.abs(got - want) / np.maximum(1.0, np.abs(want))
        assert rel.max() <= 1e-5, (name, rel.max())
        assert np.array_equal(got[0], got[1]), (name, np.abs(got[0] - got[1]).max())
