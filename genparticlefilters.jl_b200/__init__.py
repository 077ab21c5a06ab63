"""genparticlefilters.jl_b200 -- B200-native engine behind GenParticleFilters.jl's API.

Host mirror of the reference's exported functions (api.py) over the C-ABI shared library
libgenpf_cuda.so (csrc/, include/genpf.h).  Import as ``import genpf_b200`` (the directory
name carries a dot, so the repo-root shim genpf_b200.py loads it).
"""
from . import _lib
from ._lib import GenPFError, load
from .api import *  # noqa: F401,F403
from .api import (choiceproduct, get_log_weights, get_traces, pf_introduce, sample_unweighted_traces)  # noqa: F401
from .api import (DeviceModel, DevicePFState, GenPFErrorException, ParticleFilterState, ParticleFilterSubState,
                  effective_sample_size, get_ess, get_lml_est, get_log_norm_weights, get_norm_weights,
                  log_ml_estimate, logsumexp_host, mean, mh, move_reweight, pf_coalesce, pf_dereplicate, pf_initialize,
                  pf_move_accept, pf_move_reweight, pf_multinomial_resample, pf_multinomial_resize, pf_optimal_resize,
                  pf_rejuvenate,
                  pf_replicate, pf_resample, pf_residual_resample, pf_residual_resize, pf_resize, pf_step,
                  pf_stratified_resample, pf_update, proportionmap, var)

# SURVEY.md 8(d) algorithmic bytes, defined ONCE here (bench.py and DESIGN.md cite this)
ALGO_BYTES = {
    "ess_lse": 8,               # read lw once
    "resample_host_abi": 24,    # R lw 8 + W parents(Int64) 8 + W lw 8
    "object_motion_step": 117,  # update 34 + resample 56 + MH 27
    "lingauss_step": 68,        # update 32 + resample 36
}
