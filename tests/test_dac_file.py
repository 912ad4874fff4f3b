"""`.dac` code container (AudioTools/DACFile.cs:27-105): the backend's reader / writer against the oracle
restatement and hand-derived known answers.  Host-only (no GPU)."""
import json
import struct

import numpy as np
import pytest

import neuralcodecs_b200 as nc
from neuralcodecs_b200 import dac_file
from oracle import dac_file as oracle_file


def test_known_answer_bytes(tmp_path):
    """Bytes derived by hand from BinaryWriter's encodings for a 2x3 code tensor."""
    cfg = nc.DACConfig.DAC44kHz()
    codes = np.array([[1, 2, 1023], [0, 515, 7]], dtype=np.int64)
    path = tmp_path / "a.dac"
    nc.DACFile([codes], cfg).Save(str(path))
    blob = path.read_bytes()
    js = dac_file.config_to_json(cfg)
    n = len(js)
    assert n >= 128                                            # the 7-bit length takes two bytes for this JSON
    head = struct.pack("<i", n) + bytes([(n & 0x7F) | 0x80, n >> 7]) + js.encode()
    tail = (struct.pack("<i", 1) + struct.pack("<i", 2) + struct.pack("<qq", 2, 3) + struct.pack("<i", 6) +
            struct.pack("<6i", 1, 2, 1023, 0, 515, 7))
    assert blob == head + tail
    assert blob == oracle_file.save(js, [codes])
    # System.Text.Json property names / order of Config/DAC/DACConfig.cs
    d = json.loads(js)
    assert list(d)[:5] == ["Metadata", "model_type", "codebook_dim", "codebook_loss_weight", "codebook_size"]
    assert d["sampling_rate"] == 44100 and d["n_codebooks"] == 9 and d["downsampling_ratios"] == [2, 4, 8, 8]
    assert d["upsampling_ratios"] == [8, 8, 4, 2] and d["hop_length"] == 512 and d["latent_dim"] is None


@pytest.mark.parametrize("preset", ["DAC44kHz", "DAC24kHz", "DAC16kHz", "DAC44kHz_16kbps"])
def test_round_trip_and_cross_read(tmp_path, preset):
    cfg = getattr(nc.DACConfig, preset)()
    rng = np.random.default_rng(5)
    codes = [rng.integers(0, cfg.codebook_size, size=(2, cfg.num_codebooks, 431), dtype=np.int64),
             rng.integers(0, cfg.codebook_size, size=(1, cfg.num_codebooks, 17), dtype=np.int64),
             np.zeros((1, cfg.num_codebooks, 0), np.int64)]   # empty chunk: count 0
    path = tmp_path / "b.dac"
    nc.DACFile(codes, cfg).Save(str(path))
    back = nc.DACFile.Load(str(path))
    assert len(back.Codes) == 3
    for a, b in zip(codes, back.Codes):
        assert b.dtype == np.int64 and a.shape == b.shape
        np.testing.assert_array_equal(a, b)
    for f in ("sample_rate", "encoder_dim", "encoder_rates", "decoder_dim", "decoder_rates", "num_codebooks", "codebook_size",
              "codebook_dim", "latent_dim"):
        assert getattr(back.Config, f) == getattr(cfg, f), f
    # the oracle reads what the backend wrote, and the backend reads what the oracle writes
    ocfg, ocodes = oracle_file.load(path.read_bytes())
    assert ocfg["sampling_rate"] == cfg.sample_rate
    for a, b in zip(codes, ocodes):
        np.testing.assert_array_equal(a, b)
    p2 = tmp_path / "c.dac"
    p2.write_bytes(oracle_file.save(dac_file.config_to_json(cfg), codes))
    again = nc.DACFile.Load(str(p2))
    for a, b in zip(codes, again.Codes):
        np.testing.assert_array_equal(a, b)


def test_truncated_and_inconsistent_files(tmp_path):
    cfg = nc.DACConfig.DAC44kHz()
    codes = np.arange(24, dtype=np.int64).reshape(1, 3, 8)
    path = tmp_path / "d.dac"
    nc.DACFile([codes], cfg).Save(str(path))
    blob = path.read_bytes()
    for cut in (2, 40, len(blob) - 5):
        bad = tmp_path / f"cut{cut}.dac"
        bad.write_bytes(blob[:cut])
        with pytest.raises((EOFError, ValueError)):            # BinaryReader: EndOfStreamException
            nc.DACFile.Load(str(bad))
    # element count that disagrees with the shape: tensor(data).reshape(shape) throws in the reference
    mod = bytearray(blob)
    off = len(blob) - 4 * 24 - 4 - 8 * 3
    mod[off:off + 8] = struct.pack("<q", 5)
    bad = tmp_path / "shape.dac"
    bad.write_bytes(bytes(mod))
    with pytest.raises(ValueError):
        nc.DACFile.Load(str(bad))
