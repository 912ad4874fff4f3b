"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, and refuses to work without a CUDA device (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "neuralcodecs_cuda.h")


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"NC_API\s+[\w\s\*]+?\b(nc_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("nc_create", "nc_destroy", "nc_load_weights", "nc_dac_encode", "nc_dac_decode",
                 "nc_dac_from_codes", "nc_dac_decode_codes", "nc_dac_forward", "nc_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from neuralcodecs_b200 import _lib
    lib = _lib.lib()
    declared = _declared()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES and the header disagree"
    assert b"sm_100a" in lib.nc_version()


def test_config_structs_match_header_layout():
    from neuralcodecs_b200 import _lib
    # 4-byte fields only: sizes follow directly from the header's field lists
    assert C.sizeof(_lib.nc_dac_config) == 4 * (4 + 8 + 2 + 8 + 4)
    assert C.sizeof(_lib.nc_snac_config) == 4 * (4 + 8 + 2 + 8 + 4 + 1 + 8 + 2)
    assert C.sizeof(_lib.nc_encodec_config) == 4 * (6 + 8 + 5 + 4)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import neuralcodecs_b200 as nc
    from neuralcodecs_b200 import _lib
    assert _lib.lib().nc_device_count() == 0
    with pytest.raises(RuntimeError, match="CUDA requested but not available"):
        nc.DAC(nc.DACConfig.DAC44kHz())


def test_null_and_bad_arguments_are_status_codes_not_crashes():
    from neuralcodecs_b200 import _lib
    lib = _lib.lib()
    assert lib.nc_destroy(None) == _lib.NC_OK
    assert lib.nc_load_weights(None, b"x") == _lib.NC_INVALID_ARGUMENT
    assert "null handle" in _lib.last_error()
    h = C.c_void_p()
    assert lib.nc_create(_lib.NC_CODEC_DAC, None, 0, 0, C.byref(h)) != _lib.NC_OK
    assert lib.nc_launch_count(None) == 0
    # the segmented Encodec entry points (48 kHz preset) reject a null handle the same way
    n, tot, dl = C.c_int32(), C.c_int64(), C.c_int64()
    assert lib.nc_encodec_query_frames(None, 48000, 6.0, C.byref(n), None, 0, C.byref(tot), None, C.byref(dl)) == _lib.NC_INVALID_ARGUMENT
    assert lib.nc_encodec_encode_frames(None, None, 1, 48000, 6.0, None, None) == _lib.NC_INVALID_ARGUMENT
    assert lib.nc_encodec_decode_frames(None, None, None, 1, 4, None, 1, None) == _lib.NC_INVALID_ARGUMENT
    assert lib.nc_encodec_query_decoded(None, None, 1, C.byref(dl)) == _lib.NC_INVALID_ARGUMENT


def test_encodec_config_presets_and_segment_algebra():
    import neuralcodecs_b200 as nc
    c24, c48 = nc.EncodecConfig.Encodec24Khz(), nc.EncodecConfig.Encodec48Khz()       # EncodecConfig.cs:9-66
    assert (c24.channels, c24.use_causal_conv, c24.norm_type, c24.normalize, c24.segment_length, c24.segment_stride) == \
        (1, True, "weight_norm", False, None, None)
    assert (c48.sample_rate, c48.channels, c48.use_causal_conv, c48.norm_type, c48.normalize) == (48000, 2, False, "time_group_norm", True)
    assert (c48.segment_length, c48.segment_stride, c48.num_quantizers, c48.hop_length) == (48000, 47520, 16, 320)   # Encodec.cs:70-71,190-196
    j = nc.EncodecConfig.from_json('{"sampling_rate": 48000, "audio_channels": 2, "chunk_length_s": 1.0, "overlap": 0.01,'
                                   ' "norm_type": "time_group_norm", "normalize": true, "use_causal_conv": false}')
    assert (j.segment_length, j.segment_stride, j.norm_type) == (48000, 47520, "time_group_norm")
    j.overlap = None                                                                    # Encodec.cs:84: `config.Overlap ?? 0`
    assert j.segment_stride == 48000


def test_dac_config_presets_and_json():
    import neuralcodecs_b200 as nc
    c = nc.DACConfig.DAC44kHz()
    assert (c.hop_length, c.resolved_latent_dim, c.num_codebooks) == (512, 1024, 9)
    c24 = nc.DACConfig.DAC24kHz()
    assert c24.encoder_rates == [2, 4, 5, 8] and c24.hop_length == 320 and c24.num_codebooks == 32
    j = nc.DACConfig.from_json('{"sampling_rate": 16000, "n_codebooks": 12, "downsampling_ratios": [2,4,5,8],'
                               ' "upsampling_ratios": [8,5,4,2], "encoder_hidden_size": 64}')
    assert j.sample_rate == 16000 and j.num_codebooks == 12 and j.hop_length == 320


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neuralcodecs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports oracle/"
