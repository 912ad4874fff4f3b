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
from .dac_file import DACFile  # noqa: F401

__all__ = ["DAC", "DACConfig", "SNAC", "SNACConfig", "Encodec", "EncodecCompressor", "EncodecConfig", "DeviceConfiguration", "CodecException", "DACFile"]


def inspect_weights(path: str) -> dict:
    """Host-only look into a weight file (``nc_inspect_weights``): {"format": "torch_zip" | "safetensors", "metadata": {...},
    "tensors": {name: {"dtype", "shape"}}}.  Mirrors what DACUnpickler.LoadWithConfig reads before the model exists
    (Config/DAC/DACUnpickler.cs:383-424)."""
    import ctypes as C
    import json

    from . import _lib
    size = 1 << 20
    while True:
        buf = C.create_string_buffer(size)
        st = _lib.lib().nc_inspect_weights(str(path).encode(), buf, len(buf))
        if st == _lib.NC_INVALID_ARGUMENT and "too small" in _lib.last_error() and size < (1 << 28):
            size *= 4
            continue
        _lib.check(st, "weights", "Inspect")
        return json.loads(buf.value.decode())


__all__.append("inspect_weights")
