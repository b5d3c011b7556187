"""givepose_b200 -- B200-native (sm_100a) DCNv3 + per-RoI PoseNet hot path of ziqin-h/GIVEPose.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every hot op is a hand-written
CUDA kernel behind the C ABI in ``include/givepose_b200.h`` (``givepose_b200/lib/libgivepose_b200.so``).
There is no CPU fallback: importing the ops without the built library raises.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .functions import DCNv3Function, dcnv3_backward, dcnv3_forward  # noqa: F401

__version__ = "0.1.0"
