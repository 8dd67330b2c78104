"""coulomb_oscillators_b200 -- B200-native (sm_100a) force-evaluation / time-stepping path of
locuoco/coulomb_oscillators behind the C ABI of include/nbco.h.

The product is libnbco.so (hand-written CUDA + C++ host code).  This Python package is only the
thin ctypes binding used by tests/ and bench.py; it contains no numerical code and no fallback:
importing it without the built library, or creating a context without a GPU, raises.
"""
from ._lib import (  # noqa: F401
    Config, Context, NbcoError, lib, lib_path,
    EVAL_DIRECT3, EVAL_FMM3_KD, EVAL_COULOMB_DIRECT3, EVAL_COULOMB_FMM3_KD,
    EVAL_DIRECT2, EVAL_FMM2, EVAL_COULOMB_DIRECT2, EVAL_COULOMB_FMM2,
    EULER, LEAPFROG, FORESTRUTH, PEFRL,
    init_ga, init_test_cube, shard_range, default_param,
    init_ga2, init_kv2, beam_params2, default_param2, fmm2_levels, OMEGA0_2D, EMIT_2D,
)
