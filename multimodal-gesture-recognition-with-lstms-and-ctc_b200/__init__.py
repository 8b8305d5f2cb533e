"""B200-native BLSTM -> softmax -> CTC -> decode hot path of
AlexGidiotis/Multimodal-Gesture-Recognition-with-LSTMs-and-CTC (importable as ``mgr_b200``).

Python host mirroring the reference's operator surface over hand-written sm_100a CUDA
(libgr_b200.so, C ABI in include/gr_b200.h).  There is no CPU fallback.
"""
from . import _lib, ops  # noqa: F401
from ._lib import GrError, InvalidArgumentError  # noqa: F401
from .losses import ctc_lambda_func, ctc_batch_cost, softmax_ctc  # noqa: F401
from .layers import BidirectionalLSTM, DenseSoftmax, blstm  # noqa: F401
from .models import (SpeechNet, SkeletalNet, FusionNet, UnimodalNet, KerasAdam, fusion_optimizer,  # noqa: F401
                     FusionTrainer)
from .sequence_decoding import decode_batch, decode_batch_speech, decode_ids, ctc_decode  # noqa: F401
from . import custom_ops  # noqa: F401  (registers torch.ops.mgr_b200.*)

__version__ = "0.1.0"
