"""dxt-lossless-transform on B200: the BCn block transform / untransform / best-settings path of
Sewer56/dxt-lossless-transform as hand-written sm_100a CUDA kernels behind the reference's C ABI.

The product is ``libdxt_lossless_transform_cuda.so`` (csrc/, declared in include/*.h); this package
is the host-side mirror of the reference's Rust interface over that ABI.
"""
from .api import *  # noqa: F401,F403
from .api import (  # noqa: F401
    Bc1AutoTransformBuilder,
    Bc1EstimateSettings,
    Bc1ManualTransformBuilder,
    Bc1TransformSettings,
    Bc2AutoTransformBuilder,
    Bc2ManualTransformBuilder,
    Bc2TransformSettings,
    Bc3TransformSettings,
    LosslessTransformUtilsSizeEstimation,
    YCoCgVariant,
)

from . import experimental, file_formats  # noqa: E402,F401
