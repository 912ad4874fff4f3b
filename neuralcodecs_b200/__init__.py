"""neuralcodecs_b200 -- Blackwell (sm_100a) backend for the NeuralCodecs codec hot path.

Host-side mirror (Python, over ctypes) of the reference's model classes; all arithmetic is
in ``libneuralcodecs_cuda.so`` (hand-written CUDA, C ABI in include/neuralcodecs_cuda.h).
There is no CPU fallback and nothing here imports ``oracle/``.
"""
from .config import DACConfig, DeviceConfiguration, EncodecConfig, SNACConfig  # noqa: F401
from .dac import DAC  # noqa: F401
from .snac import SNAC  # noqa: F401
from .encodec import Encodec, EncodecCompressor  # noqa: F401
from ._lib import CodecException  # noqa: F401

__all__ = ["DAC", "DACConfig", "SNAC", "SNACConfig", "Encodec", "EncodecCompressor", "EncodecConfig", "DeviceConfiguration", "CodecException"]
