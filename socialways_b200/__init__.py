"""socialways_b200 -- B200-native (sm_100a) implementation of the Social Ways hot path.

Public surface = the reference's own operator interface for the path (reference_api.py, trainer.py) and for the rows
around it (statistics.py = calc_statistics.py, dataset.py = create_dataset, fused_optim.py = the Adam optimisers) on top
of a C-ABI CUDA library (include/socialways_b200.h, csrc/).  No CPU fallback exists.
"""
from ._lib import SocialWaysCudaError, LIB_PATH, exported_symbols  # noqa: F401
from .reference_api import (AttentionPooling, DecoderFC, DecoderLstm, Discriminator, EmbedSocialFeatures,  # noqa: F401
                            EncoderLstm, Generator, SocialFeatures, get_traj_4d, predict_cv)
from . import dataset, fused_optim, ops, packing, statistics  # noqa: F401

__version__ = "0.1.0"
