"""Device versions of the reference's input-conditioning helpers (NeuralCodecs.Core/Utils/AudioUtils.cs): the step before
the codec path.  Each takes a loaded model (any codec) whose device and stream run the kernel; results are bit-identical
to the reference's double / float host loops."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def ResampleLinear(model, input, sourceSampleRate: int, targetSampleRate: int) -> np.ndarray:
    """AudioUtils.ResampleLinear (AudioUtils.cs:329-352) / SNAC.ResampleAudio (Models/SNAC.cs:284-308).
    input: [L] or [B, L] float32 -> [L'] or [B, L'], L' = int(L * target / source)."""
    a = np.ascontiguousarray(input, dtype=np.float32)
    flat = a.ndim == 1
    a = a.reshape(1, -1) if flat else a
    B, L = a.shape
    n = C.c_int64()
    lib = _lib.lib()
    _lib.check(lib.nc_resample_linear(model._handle(), None, B, L, int(sourceSampleRate), int(targetSampleRate), None, 0,
                                      C.byref(n)), "AudioUtils", "ResampleLinear")
    out = np.empty((B, n.value), np.float32)
    if n.value:
        _lib.check(lib.nc_resample_linear(model._handle(), a.ctypes.data_as(C.c_void_p), B, L, int(sourceSampleRate),
                                          int(targetSampleRate), out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)),
                   "AudioUtils", "ResampleLinear")
    return out.reshape(-1) if flat else out


def ConvertToMono(model, input, channels: int) -> np.ndarray:
    """AudioUtils.ConvertToMono(float[], channels) (AudioUtils.cs:45-62): interleaved samples -> channel average."""
    a = np.ascontiguousarray(input, dtype=np.float32).reshape(-1)
    frames = a.size // int(channels)
    out = np.empty(frames, np.float32)
    if frames:
        _lib.check(_lib.lib().nc_convert_to_mono(model._handle(), a.ctypes.data_as(C.c_void_p), frames, int(channels),
                                                 out.ctypes.data_as(C.c_void_p)), "AudioUtils", "ConvertToMono")
    return out
