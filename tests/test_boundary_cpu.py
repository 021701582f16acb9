"""CPU suite, part 4: SCIPsdpiSolver* boundary checks of our binding linked to the CPU oracle (oracle/_ref/libsdpi_oracle.so)."""
import os

import pytest

from harness import boundary_cases, sdpi_ref

pytestmark = pytest.mark.skipif(not os.path.exists(sdpi_ref.LIB_ORACLE), reason="oracle/_ref/libsdpi_oracle.so not built")


def test_penalty_formulation_call_patterns():
    boundary_cases.run_penalty_patterns(sdpi_ref.LIB_ORACLE)


def test_primal_matrix_getters_are_consistent():
    boundary_cases.run_primal_getters(sdpi_ref.LIB_ORACLE)


def test_warmstart_and_preoptimal_solution():
    boundary_cases.run_warmstart_and_preoptimal(sdpi_ref.LIB_ORACLE)
