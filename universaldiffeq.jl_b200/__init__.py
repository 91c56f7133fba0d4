"""universaldiffeq.jl_b200 -- B200-native loss-and-gradient engine for UniversalDiffEq-style UDE training.

Host-side mirror of the reference's API for the hot path only: model constructors,
train!, cross-validation drivers, all calling libude_cuda.so through its C ABI (include/ude_cuda.h).
The directory name carries a dot, so import it through the repo-root shim:  ``import ude_b200 as ude``.
"""
from . import problem, smoothing, synthetic, sharding, simulate  # noqa: F401
from .problem import Problem  # noqa: F401
from . import _capi  # noqa: F401
from ._capi import (Engine, UdeError, device_count, measure_fp64_peak, get_unique_id, load_library, LIB_PATH,  # noqa: F401
                    OPT_SPARSE_UHAT_IO, OPT_UKF_CAUSAL_INTERVAL)
from . import user_rhs  # noqa: F401
from .user_rhs import expression_rhs, UserRhs  # noqa: F401
from .models import (one_hot_series_covariates, CustomDerivatives, NODE, MultiNODE, MultiCustomDerivatives, SimpleNeuralNetwork, CudaRK4, CudaTsit5,  # noqa: F401
                     LV_UDE, LV_UDE_X, LV_UDE_ABS, FISHERY, node_rhs, node_time_rhs, nn_decay_rhs, series_growth_rhs, UDE)
from .training import (train_, adam_multi_gpu, gradient_descent_, marginal_likelihood_, marginal_likelihood_many_, spline_gradient_matching_, SplineGradientMatching, neural_gradient_matching_, NeuralGradientMatching, kalman_filter_, save_model_parameters, load_model_parameters_, conditional_likelihood, default_loss, multiple_shooting_loss, shooting_loss,  # noqa: F401
                       derivative_matching_loss, errors_to_matrix, build_problem, predictions, forecast, adam_host)
from .cross_validation import (leave_future_out, leave_future_out_batched, leave_future_out_predict, summarize_leave_future_out,  # noqa: F401
                               interval_holdout_cv, stack_models, leave_site_out, energy_score, predict)
from .simulate import LotkaVolterra, LorenzLotkaVolterra, simulate_coral_data  # noqa: F401
