"""B200-native Newton-step hot path of the predictor-corrector interior-point LP solver
payakorn/InteriorPointMethod: hand-written sm_100a CUDA behind a C ABI (include/ipm_b200.h), with a thin
Python mirror of the reference's call surface."""
from .solver import (NewtonStep, Result, check_optimality, corrected, direction_corrected_sparse,  # noqa: F401
                     direction_predicted_sparse, duality_gap, full_stepsize, interior, interior_kkt,
                     interior_sparse,
                     newton_iteration, predicted_stepsize, predicted_stepsize_lb_ub, full_stepsize_lb_ub,
                     release_cached_step, solve, solve_linear, step_size)
from . import general_form  # noqa: F401
from .general_form import (add_bound_into_matrix, create_problem_from_mps_matlab, get_Abc,  # noqa: F401
                           new_interior_sparse, standard_form)
from .problems import (create_problem_from_mps, load_golden_problem, synthetic_dense_batch,  # noqa: F401
                       synthetic_dense_lp)

__all__ = ["NewtonStep", "Result", "solve", "interior_sparse", "interior", "interior_kkt", "direction_predicted_sparse",
           "direction_corrected_sparse", "check_optimality", "predicted_stepsize", "full_stepsize", "duality_gap",
           "corrected", "solve_linear", "newton_iteration", "release_cached_step", "create_problem_from_mps",
           "load_golden_problem", "synthetic_dense_lp", "synthetic_dense_batch", "get_Abc", "add_bound_into_matrix",
           "standard_form", "new_interior_sparse", "create_problem_from_mps_matlab", "step_size",
           "predicted_stepsize_lb_ub", "full_stepsize_lb_ub"]
