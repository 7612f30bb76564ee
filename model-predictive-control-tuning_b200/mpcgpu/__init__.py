"""mpcgpu: B200-native batched closed-loop MPC evaluation behind the reference's evaluator API."""
from .plant import Channels, c2d_fopdt, simulate, cond_min  # noqa: F401
from .problems import LinearProblem, shell3x3, woodberry, shell7x5, synthetic_population, CASES  # noqa: F401
from .api import Evaluator, MultiEvaluator, MpcGpuError, closedloop_toolbox, gam_fun, vns_cost, measure_fp64_peak, row2col, col2row  # noqa: F401
from .dtcgpc import DtcProblem, DtcEvaluator, woodberry_dtc, synthetic_dtc_population, robustness_filter, mimo_filter, dtc_gpc_ww  # noqa: F401
from .nmpc import NmpcProblem, NmpcEvaluator, vandevusse, synthetic_nmpc_population, closedloop_toolbox_nmpc  # noqa: F401
from . import tuner  # noqa: F401
from .ssnmpc import SsnmpcProblem, SsnmpcEvaluator, explicit_nmpc, synthetic_ssnmpc_population, ClosedLoopNMPC  # noqa: F401
