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
