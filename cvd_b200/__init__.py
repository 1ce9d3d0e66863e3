"""cvd_b200 -- B200-native (sm_100a) implementation of CVD's cross-video synchronisation hot
path: the epipolar cross-view attention module and the temporal-attention glue around it.

Public surface = the reference's own plugin interface for this path
(``animatediff/models/epi_module.py``, ``motion_module.py``, ``attention_processor.py``):

    get_epi_module, EpiModule, EpiTransformer3DModel, EpiTransformerBlock, EpiSelfAttention,
    EpiEncoding, EpiAttnProcessor, get_motion_module, VanillaTemporalModule, ...

All compute runs in ``cvd_b200/lib/libcvd_b200.so`` (C ABI: ``include/cvd_b200.h``); importing
the package does not load it, the first op does -- and raises if it has not been built.
"""
from . import _lib, geometry, ops, training  # noqa: F401
from ._lib import CvdError  # noqa: F401
from .graphs import GraphedModule  # noqa: F401
from .epi_module import (EpiAttnProcessor, EpiEncoding, EpiModule, EpiSelfAttention, EpiTransformer3DModel,  # noqa: F401
                         EpiTransformerBlock, FeedForward, get_epi_module)
from .motion_module import (AttnProcessor, PoseAdaptorAttnProcessor, PositionalEncoding, TemporalSelfAttention,  # noqa: F401
                            TemporalTransformer3DModel, TemporalTransformerBlock, VanillaTemporalModule,
                            get_motion_module)

from .pose_encoder import CameraPoseEncoder  # noqa: F401
from .unet import UNet3DConditionModelPoseCond  # noqa: F401
from .pipeline import MultiviewPipeline, PairPipeline  # noqa: F401

__version__ = "0.2.0"
