import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CACHE = os.path.join(ROOT, "tests", ".cache")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def cache_dir():
    os.makedirs(CACHE, exist_ok=True)
    return CACHE


def _write_weights(cfg, path, codebooks, seconds):
    from oracle import synth
    if not os.path.exists(path):
        sd = synth.make_dac_weights_hf(cfg, codebooks=codebooks, codebook_seconds=seconds)
        synth.save_safetensors(sd, path + ".tmp")
        os.replace(path + ".tmp", path)
    return path


@pytest.fixture(scope="session")
def dac_tiny(cache_dir):
    """(oracle cfg, engine cfg, weights path): 16 kHz, channels not multiples of 32 -> fp32 CUDA-core path."""
    from oracle import dac as odac
    import neuralcodecs_b200 as nc
    co = odac.DACConfig(sample_rate=16000, encoder_dim=16, decoder_dim=128, n_codebooks=4, codebook_size=64)
    ce = nc.DACConfig(sample_rate=16000, encoder_dim=16, decoder_dim=128, num_codebooks=4, codebook_size=64)
    return co, ce, _write_weights(co, os.path.join(cache_dir, "dac_tiny.safetensors"), "data", 2.0)


@pytest.fixture(scope="session")
def dac_mid(cache_dir):
    """Small config whose channel counts are multiples of 32 -> tcgen05 path."""
    from oracle import dac as odac
    import neuralcodecs_b200 as nc
    co = odac.DACConfig(sample_rate=16000, encoder_dim=32, decoder_dim=512, n_codebooks=4, codebook_size=256)
    ce = nc.DACConfig(sample_rate=16000, encoder_dim=32, decoder_dim=512, num_codebooks=4, codebook_size=256)
    return co, ce, _write_weights(co, os.path.join(cache_dir, "dac_mid.safetensors"), "data", 2.0)


@pytest.fixture(scope="session")
def dac_24k_geometry(cache_dir):
    """The 24 kHz preset's geometry (odd stride 5: rates 2,4,5,8 / 8,5,4,2; DACConfig.cs:113-124) at reduced width."""
    from oracle import dac as odac
    import neuralcodecs_b200 as nc
    co = odac.DACConfig(sample_rate=24000, encoder_dim=32, encoder_rates=[2, 4, 5, 8], decoder_dim=512, decoder_rates=[8, 5, 4, 2],
                        n_codebooks=6, codebook_size=128)
    ce = nc.DACConfig(sample_rate=24000, encoder_dim=32, encoder_rates=[2, 4, 5, 8], decoder_dim=512, decoder_rates=[8, 5, 4, 2],
                      num_codebooks=6, codebook_size=128)
    return co, ce, _write_weights(co, os.path.join(cache_dir, "dac_24k_geometry.safetensors"), "data", 1.0)


@pytest.fixture(scope="session")
def dac_full(cache_dir):
    """DAC 44.1 kHz preset (BASELINE configs #1/#4/#5) with seeded weights + data-fitted codebooks."""
    from oracle import dac as odac
    import neuralcodecs_b200 as nc
    co, ce = odac.DACConfig.dac_44khz(), nc.DACConfig.DAC44kHz()
    return co, ce, _write_weights(co, os.path.join(cache_dir, "dac44_seed4321.safetensors"), "data", 10.0)


def _write_snac(cfg, path, clips, seconds):
    from oracle import synth
    if not os.path.exists(path):
        sd = synth.make_snac_weights(cfg, codebook_clips=clips, codebook_seconds=seconds)
        synth.save_safetensors(sd, path + ".tmp")
        os.replace(path + ".tmp", path)
    return path


@pytest.fixture(scope="session")
def snac_tiny(cache_dir):
    """Small depthwise SNAC (no attention), odd channel counts -> exercises channel padding."""
    from oracle import snac as osnac
    import neuralcodecs_b200 as nc
    kw = dict(sample_rate=16000, encoder_dim=12, encoder_rates=[2, 4, 4], decoder_dim=96, decoder_rates=[4, 4, 2],
              attn_window_size=None, codebook_size=128, vq_strides=[4, 2, 1])
    co, ce = osnac.SNACConfig(**kw), nc.SNACConfig(**kw)
    return co, ce, _write_snac(co, os.path.join(cache_dir, "snac_tiny.safetensors"), 2, 2.0)


@pytest.fixture(scope="session")
def snac_24k(cache_dir):
    """SNAC 24 kHz preset (BASELINE config #2)."""
    from oracle import snac as osnac
    import neuralcodecs_b200 as nc
    co, ce = osnac.SNACConfig.snac_24khz(), nc.SNACConfig.SNAC24kHz()
    return co, ce, _write_snac(co, os.path.join(cache_dir, "snac24_seed4321.safetensors"), 4, 4.0)


def _write_encodec(cfg, path, clips, seconds):
    from oracle import synth
    if not os.path.exists(path):
        sd = synth.make_encodec_weights(cfg, codebook_clips=clips, codebook_seconds=seconds)
        synth.save_safetensors(sd, path + ".tmp")
        os.replace(path + ".tmp", path)
    return path


@pytest.fixture(scope="session")
def encodec_nolstm(cache_dir):
    """24 kHz architecture without the LSTM (isolates the conv stacks + VQ)."""
    from oracle import encodec as oenc
    import neuralcodecs_b200 as nc
    co, ce = oenc.EncodecConfig(num_lstm_layers=0), nc.EncodecConfig(num_lstm_layers=0)
    return co, ce, _write_encodec(co, os.path.join(cache_dir, "encodec_nolstm.safetensors"), 2, 4.0)


@pytest.fixture(scope="session")
def encodec_24k(cache_dir):
    """Encodec 24 kHz preset at 6 kbps (BASELINE config #3)."""
    from oracle import encodec as oenc
    import neuralcodecs_b200 as nc
    co, ce = oenc.EncodecConfig(), nc.EncodecConfig.Encodec24Khz()
    return co, ce, _write_encodec(co, os.path.join(cache_dir, "encodec24_seed4321.safetensors"), 2, 6.0)


@pytest.fixture(scope="session")
def snac_attn(cache_dir):
    """Small SNAC with LocalMHA (window 32), an odd stride (3: conv_transpose output_padding = 1) and 4 VQ strides."""
    from oracle import snac as osnac
    import neuralcodecs_b200 as nc
    kw = dict(sample_rate=32000, encoder_dim=16, encoder_rates=[2, 3, 2], decoder_dim=256, decoder_rates=[2, 3, 2],
              attn_window_size=32, codebook_size=128, vq_strides=[4, 2, 1])
    co, ce = osnac.SNACConfig(**kw), nc.SNACConfig(**kw)
    return co, ce, _write_snac(co, os.path.join(cache_dir, "snac_attn.safetensors"), 4, 1.0)


@pytest.fixture(scope="session", params=[16, 48])
def snac_attn_window(cache_dir, request):
    """The snac_attn geometry with a LocalMHA window other than the presets' 32 (Modules/SNAC/LocalMHA.cs:46-70 takes any)."""
    from oracle import snac as osnac
    import neuralcodecs_b200 as nc
    kw = dict(sample_rate=32000, encoder_dim=16, encoder_rates=[2, 3, 2], decoder_dim=256, decoder_rates=[2, 3, 2],
              attn_window_size=request.param, codebook_size=128, vq_strides=[4, 2, 1])
    co, ce = osnac.SNACConfig(**kw), nc.SNACConfig(**kw)
    return co, ce, _write_snac(co, os.path.join(cache_dir, f"snac_attn_w{request.param}.safetensors"), 4, 1.0)


@pytest.fixture(scope="session")
def snac_44k(cache_dir):
    """SNAC 44 kHz preset (LocalMHA, vq strides 8/4/2/1, stride-3 blocks): SURVEY 8(d) extra coverage run."""
    from oracle import snac as osnac
    import neuralcodecs_b200 as nc
    co, ce = osnac.SNACConfig.snac_44khz(), nc.SNACConfig.SNAC44kHz()
    return co, ce, _write_snac(co, os.path.join(cache_dir, "snac44_seed4321.safetensors"), 2, 2.0)


@pytest.fixture(scope="session")
def encodec_48k(cache_dir):
    """Encodec 48 kHz preset (stereo, non-causal, time_group_norm, normalize, 1 s segments, 1 % overlap) at 6 kbps."""
    from oracle import encodec as oenc
    import neuralcodecs_b200 as nc
    co, ce = oenc.EncodecConfig.encodec_48khz(), nc.EncodecConfig.Encodec48Khz()
    return co, ce, _write_encodec(co, os.path.join(cache_dir, "encodec48_seed4321.safetensors"), 2, 3.0)


@pytest.fixture(scope="session")
def encodec_48k_one_frame(cache_dir):
    """The 48 kHz architecture without segments and without the loudness scale (one un-normalised frame per clip)."""
    from oracle import encodec as oenc
    import neuralcodecs_b200 as nc
    co, ce = oenc.EncodecConfig.encodec_48khz(), nc.EncodecConfig.Encodec48Khz()
    co.chunk_length_s = ce.chunk_length_s = None
    co.normalize = ce.normalize = False
    return co, ce, _write_encodec(co, os.path.join(cache_dir, "encodec48_oneframe_seed4321.safetensors"), 2, 3.0)
