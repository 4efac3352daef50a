"""B200-native batched iLQG/DDP engine: a drop-in for the hot path of
DifferentialDynamicProgramming.jl (src/backward_pass.jl, src/forward_pass.jl, src/boxQP.jl).

The directory name carries the reference's ``.jl``; import it as ``ddp_b200`` (the loader module
at the repository root) -- ``import ddp_b200 as ddp``.
"""
from ._lib import DDPError, DDPLibraryMissing, LIB_PATH, load  # noqa: F401
from .api import (DEFAULT_ALPHA, STATUS, DevArray, Engine, GaussianPolicy, HostIteration, LinearModel,  # noqa: F401
                  PendcartModel, PosDefException, SimpleLTVModel, TRACE_DTYPE, back_pass, back_pass_gps, boxQP, forward_pass, forward_costs, iLQG, iLQGkl, iLQGkl_device, KL_STATUS,
                  kl_div_wiki, iterate_chunked, demoQP)
