"""ctypes binding of libneuralcodecs_cuda.so (include/neuralcodecs_cuda.h).

This is the Python stand-in for the C# ``[LibraryImport]`` stubs shown in INTEGRATION.md:
the same entry points, the same status-code -> exception mapping.  There is no fallback:
if the shared library has not been built, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libneuralcodecs_cuda.so")

NC_OK = 0
NC_INVALID_ARGUMENT = 1
NC_FILE_NOT_FOUND = 2
NC_BAD_WEIGHTS = 3
NC_SHAPE_MISMATCH = 4
NC_CUDA_UNAVAILABLE = 5
NC_CUDA_ERROR = 6
NC_OUT_OF_MEMORY = 7
NC_INTERNAL = 8
NC_UNSUPPORTED = 9

NC_CODEC_DAC, NC_CODEC_SNAC, NC_CODEC_ENCODEC = 1, 2, 3
NC_MAX_RATES = 8


class CodecException(RuntimeError):
    """Mirror of NeuralCodecs.Core/Exceptions/CodecException.cs:8-40 for runtime failures."""

    def __init__(self, codec: str, operation: str, message: str):
        super().__init__(f"{codec} {operation} failed: {message}")
        self.codec, self.operation = codec, operation


class nc_dac_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("sample_rate", C.c_int32), ("encoder_dim", C.c_int32),
        ("n_encoder_rates", C.c_int32), ("encoder_rates", C.c_int32 * NC_MAX_RATES),
        ("decoder_dim", C.c_int32), ("n_decoder_rates", C.c_int32),
        ("decoder_rates", C.c_int32 * NC_MAX_RATES), ("n_codebooks", C.c_int32),
        ("codebook_size", C.c_int32), ("codebook_dim", C.c_int32), ("latent_dim", C.c_int32),
    ]


class nc_snac_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("sample_rate", C.c_int32), ("encoder_dim", C.c_int32),
        ("n_encoder_rates", C.c_int32), ("encoder_rates", C.c_int32 * NC_MAX_RATES),
        ("decoder_dim", C.c_int32), ("n_decoder_rates", C.c_int32),
        ("decoder_rates", C.c_int32 * NC_MAX_RATES), ("latent_dim", C.c_int32),
        ("attn_window_size", C.c_int32), ("codebook_size", C.c_int32), ("codebook_dim", C.c_int32),
        ("n_vq_strides", C.c_int32), ("vq_strides", C.c_int32 * NC_MAX_RATES),
        ("noise", C.c_int32), ("depthwise", C.c_int32),
    ]


class nc_encodec_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("sample_rate", C.c_int32), ("channels", C.c_int32),
        ("n_filters", C.c_int32), ("dimension", C.c_int32), ("n_ratios", C.c_int32),
        ("ratios", C.c_int32 * NC_MAX_RATES), ("n_residual_layers", C.c_int32),
        ("lstm_layers", C.c_int32), ("codebook_size", C.c_int32), ("n_quantizers", C.c_int32),
        ("causal", C.c_int32), ("norm_type", C.c_int32), ("normalize", C.c_int32), ("segment_s", C.c_float),
        ("overlap", C.c_float),
    ]


# every symbol include/neuralcodecs_cuda.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F = C.POINTER(C.c_float)
_I64 = C.POINTER(C.c_int64)
SIGNATURES = {
    "nc_version": (C.c_char_p, []),
    "nc_last_error": (C.c_char_p, []),
    "nc_device_count": (C.c_int, []),
    "nc_create": (C.c_int, [C.c_int, _P, C.c_size_t, C.c_int, C.POINTER(_P)]),
    "nc_destroy": (C.c_int, [_P]),
    "nc_load_weights": (C.c_int, [_P, C.c_char_p]),
    "nc_set_tensor": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int, _I64, _P]),
    "nc_finalize_weights": (C.c_int, [_P]),
    "nc_set_option": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "nc_dac_query_shapes": (C.c_int, [_P, C.c_int64, _I64, _I64, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "nc_dac_encode": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _I64]),
    "nc_dac_decode": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P]),
    "nc_dac_from_codes": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P]),
    "nc_dac_decode_codes": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P]),
    "nc_dac_decode_dia": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _I64, _P, C.c_int64]),
    "nc_dac_forward": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _I64]),
    "nc_dac_forward_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _I64]),
    "nc_dac_decode_codes_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P]),
    "nc_snac_query_shapes": (C.c_int, [_P, C.c_int64, _I64, _I64, C.POINTER(C.c_int32), _I64, C.POINTER(C.c_int32), _I64]),
    "nc_snac_encode": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P]),
    "nc_snac_decode": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P, C.c_uint64, _P]),
    "nc_snac_forward": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P, C.c_uint64, _P, _P]),
    "nc_snac_forward_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P, C.c_uint64, _P, _P]),
    "nc_encodec_query_shapes": (C.c_int, [_P, C.c_int64, C.c_float, _I64, C.POINTER(C.c_int32), _I64]),
    "nc_encodec_encode": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_float, _P]),
    "nc_encodec_decode": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P]),
    "nc_encodec_forward": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_float, _P, _P]),
    "nc_encodec_forward_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_float, _P, _P]),
    "nc_encodec_query_frames": (C.c_int, [_P, C.c_int64, C.c_float, C.POINTER(C.c_int32), _I64, C.c_int32, _I64,
                                          C.POINTER(C.c_int32), _I64]),
    "nc_encodec_encode_frames": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_float, _P, _P]),
    "nc_encodec_decode_frames": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _I64, C.c_int32, _P]),
    "nc_encodec_query_decoded": (C.c_int, [_P, _I64, C.c_int32, _I64]),
    "nc_encodec_forward_frames_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_float, _P, _P, _P]),
    "nc_snac_process_audio": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, _P, C.c_uint64, _P, C.c_int64, _I64]),
    "nc_inspect_weights": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "nc_resample_linear": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int64, _I64]),
    "nc_convert_to_mono": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P]),
    "nc_encodec_ecdc_size": (C.c_int, [_P, C.c_int64, C.c_float, _I64, _I64]),
    "nc_encodec_compress": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_float, _P, C.c_int64, _I64]),
    "nc_encodec_ecdc_info": (C.c_int, [_P, C.c_int64, _I64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_float), C.POINTER(C.c_int32), _I64]),
    "nc_encodec_decompress": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, _P, C.c_int64, _I64, C.POINTER(C.c_int32)]),
    "nc_get_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "nc_describe": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "nc_launch_count": (C.c_uint64, [_P]),
    "nc_profile_report": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C neuralcodecs_b200/csrc`.  neuralcodecs_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def last_error() -> str:
    return lib().nc_last_error().decode("utf-8", "replace")


def check(status: int, codec: str = "codec", operation: str = "Operation") -> None:
    """Status -> exception mapping of INTEGRATION.md (the C# layer does the same)."""
    if status == NC_OK:
        return
    msg = last_error()
    if status == NC_INVALID_ARGUMENT:
        raise ValueError(msg)                      # ArgumentException
    if status == NC_FILE_NOT_FOUND:
        raise FileNotFoundError(msg)               # FileNotFoundException
    if status in (NC_BAD_WEIGHTS, NC_SHAPE_MISMATCH, NC_CUDA_UNAVAILABLE, NC_UNSUPPORTED):
        raise RuntimeError(msg)                    # InvalidOperationException
    if status == NC_OUT_OF_MEMORY:
        raise MemoryError(msg)
    raise CodecException(codec, operation, msg)    # CodecException(name, CodecOperation, msg)
