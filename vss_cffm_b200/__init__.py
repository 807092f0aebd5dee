"""vss_cffm_b200 -- B200-native (sm_100a) implementation of the VSS-CFFM inference hot path behind
the reference's mmseg plugin surface.  Importing the package registers ``mit_b0..b5``,
``CFFMHead_clips_resize1_8``, ``CFFMHead_clips_resize1_8_finetune_w_prototype3``,
``EncoderDecoder_clips`` and ``CrossEntropyLoss`` in the registries of ``vss_cffm_b200.registry``.

All arithmetic runs in libcffm_b200.so (hand-written CUDA, C ABI in include/cffm_b200.h); there is
no CPU or PyTorch fallback.
"""
from . import _abi, registry  # noqa: F401
from .config import Config  # noqa: F401
from .registry import (BACKBONES, HEADS, LOSSES, SEGMENTORS, build_backbone, build_head, build_loss,  # noqa: F401
                       build_segmentor)
from . import losses, mit, cffm_head, segmentor  # noqa: F401,E402  (registration side effects)
from .configs import model_cfg  # noqa: F401

__version__ = "0.1.0"
