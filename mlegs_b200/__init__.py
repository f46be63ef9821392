"""mlegs_b200 -- B200-native (sm_100a) implementation of the MLegS spectral-transform /
nonlinear-term hot path behind the reference's scalar / tfm_kit interfaces.

The compute path is libmlegs_b200.so (hand-written CUDA, C ABI in include/mlegs_b200.h).
This package is only the ctypes binding and the host-side mirror of the reference's types.
"""
from ._lib import MlegsError, Params, Field, lib, LIB_PATH   # noqa: F401
from .kit import TfmKit, make_params, finalize               # noqa: F401
from .scalar import *                                         # noqa: F401,F403
from .scalar import Scalar                                    # noqa: F401
from . import dist                                            # noqa: F401,E402
from . import io                                              # noqa: F401,E402
