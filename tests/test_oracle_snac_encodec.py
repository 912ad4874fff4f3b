"""CPU checks of the SNAC and Encodec oracles: length algebra, naive restatements of the ops that differ from DAC,
and (Encodec) committed transformers.EncodecModel fixtures."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import encodec as oenc
from oracle import snac as osnac
from oracle import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "encodec_hf_small.npz")


# ------------------------------------------------------------------------------------------------ SNAC
@pytest.fixture(scope="module")
def snac_small():
    cfg = osnac.SNACConfig(sample_rate=16000, encoder_dim=12, encoder_rates=[2, 4, 4], decoder_dim=96, decoder_rates=[4, 4, 2],
                           attn_window_size=None, codebook_size=64, vq_strides=[4, 2, 1])
    sd = synth.make_snac_weights(cfg, codebooks="normal")
    return cfg, osnac.SNACOracle(cfg, {k: torch.from_numpy(v) for k, v in sd.items()})


def test_snac_24k_length_algebra():
    cfg = osnac.SNACConfig.snac_24khz()
    assert cfg.hop_length == 512 and cfg.pad_multiple == 2048 and cfg.latent_dim == 768   # 240000 -> 241664
    m = osnac.SNACOracle(cfg, {})
    assert m.preprocess(torch.zeros(1, 1, 240000)).shape[-1] == 241664
    assert m.noise_lengths(472) == [3776, 30208, 120832, 241664]
    c44 = osnac.SNACConfig.snac_44khz()
    assert c44.pad_multiple == 384 * math.lcm(8, 32)                                      # SNAC.cs:74-77


def test_snac_forward_shapes_trim_and_noise(snac_small):
    cfg, m = snac_small
    x = torch.from_numpy(synth.synth_audio(2, 3001, cfg.sample_rate)).unsqueeze(1)
    Lp = m.preprocess(x).shape[-1]
    assert Lp % (cfg.hop_length * 4) == 0
    T = Lp // cfg.hop_length
    noise = [torch.from_numpy(n) for n in synth.snac_noise(2, m.noise_lengths(T))]
    out = m.forward(x, noise)
    assert out["audio"].shape == x.shape                                   # trimmed (SNAC.cs:103)
    assert [c.shape[1] for c in out["codes"]] == [T // 4, T // 2, T]
    dec = m.decode(out["codes"], noise)
    assert dec.shape[-1] == Lp                                             # Decode does not trim (SNAC.cs:157-165)
    # FromCodes reproduces the quantised latent up to the straight-through roundings
    np.testing.assert_allclose(m.rvq_from_codes(out["codes"]).numpy(), out["zq"].numpy(), atol=1e-5)
    # noise enters as x + n * linear(x): zero noise and non-zero noise differ
    assert float((m.decode(out["codes"], None) - dec).abs().max()) > 0


def test_snac_depthwise_unit_against_naive(snac_small):
    cfg, m = snac_small
    p = "encoder.block.1.block.1"                                          # ResidualUnit, dilation 3, groups = 12
    x = torch.randn(1, 12, 50)
    y = m.residual_unit(p, x, 3, 12).numpy()
    xs = m.snake(p + ".block.0", x).numpy()[0]
    w = m._weight(p + ".block.1").numpy()[:, 0, :]
    b = m.sd[p + ".block.1.bias"].numpy()
    h = np.zeros((12, 50))
    for t in range(50):
        for j in range(7):
            ti = t + (j - 3) * 3
            if 0 <= ti < 50:
                h[:, t] += w[:, j] * xs[:, ti]
    h += b[:, None]
    hs = m.snake(p + ".block.2", torch.from_numpy(h[None].astype(np.float32)))
    y2 = m.wnconv1d(p + ".block.3", hs).numpy() + x.numpy()
    np.testing.assert_allclose(y, y2, atol=2e-5)


def test_snac_vq_stride_pools_then_repeats(snac_small):
    cfg, m = snac_small
    z = torch.randn(1, cfg.latent_dim, 8)
    zq, idx, ze = m.vq_forward(0, z)                                       # stride 4
    assert idx.shape == (1, 2) and zq.shape == z.shape
    np.testing.assert_array_equal(zq[..., 0].numpy(), zq[..., 3].numpy())  # repeat_interleave(4)
    pooled = z.reshape(1, cfg.latent_dim, 2, 4).mean(-1)
    np.testing.assert_allclose(m.wnconv1d("quantizer.quantizers.0.in_proj", pooled).numpy(), ze.numpy(), atol=1e-6)


# ------------------------------------------------------------------------------------------------ Encodec
def test_encodec_config_algebra():
    cfg = oenc.EncodecConfig()
    assert (cfg.hop_length, cfg.frame_rate, cfg.num_quantizers, cfg.n_q_for_bandwidth()) == (320, 75, 32, 8)
    assert [cfg.n_q_for_bandwidth(b) for b in (1.5, 3.0, 12.0, 24.0)] == [2, 4, 16, 32]


@pytest.fixture(scope="module")
def encodec_gold():
    g = np.load(GOLD)
    cfg = oenc.EncodecConfig(num_filters=8, hidden_size=32, codebook_size=64, upsampling_ratios=[4, 3, 2],
                             target_bandwidths=[1.5, 3.0, 6.0], bandwidth=3.0)
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
    return g, cfg, oenc.EncodecOracle(cfg, sd)


def test_encodec_matches_hf_fixture(encodec_gold):
    g, cfg, m = encodec_gold
    x = torch.from_numpy(g["audio_in"])
    np.testing.assert_array_equal(synth.synth_audio(2, 2503, 24000), g["audio_in"][:, 0])
    with torch.inference_mode():
        emb = m.encoder(x)
    assert emb.shape == g["encoder_out"].shape
    np.testing.assert_allclose(emb.numpy(), g["encoder_out"], atol=2e-6)
    codes = m.rvq_encode(torch.from_numpy(g["encoder_out"]), 3.0)
    np.testing.assert_array_equal(codes.numpy(), g["codes"])               # un-normalised Euclidean argmin: pinned by HF too
    audio = m.decode(torch.from_numpy(g["codes"]))
    np.testing.assert_allclose(audio.numpy(), g["decoder_out"], atol=2e-6)
    assert m.forward(x, 3.0)["audio"].shape == x.shape                     # sliced to the input length (Encodec.cs:292-296)


def test_encodec_causal_reflect_padding_rules():
    # SConv1d.cs:144-173: causal -> all of padding_total on the left (reflect), "extra" on the right so that frames = ceil(L/stride)
    for length, k, s in ((2503, 8, 4), (17, 4, 2), (24000, 16, 8), (7, 7, 1)):
        pt = k - s
        extra = oenc.EncodecOracle._extra_padding(length, k, s, pt)
        assert (length + pt + extra - k) % s == 0 and 0 <= extra < s
        assert (length + pt + extra - k) // s + 1 == math.ceil(length / s)
    x = torch.arange(5.0).reshape(1, 1, 5)
    np.testing.assert_array_equal(oenc.EncodecOracle._pad1d(x, 2, 1).numpy().ravel(), [2, 1, 0, 1, 2, 3, 4, 3])
    small = oenc.EncodecOracle._pad1d(torch.ones(1, 1, 2), 3, 0)           # small-input branch: zero-extend, then reflect
    assert small.shape[-1] == 2 + 2 + 3


def test_encodec_lstm_matches_manual_recurrence():
    cfg = oenc.EncodecConfig(num_filters=8, hidden_size=32, upsampling_ratios=[2], codebook_size=16, num_lstm_layers=2)
    H = 16
    torch.manual_seed(0)
    sd = {}
    for l in range(2):
        for n, shape in (("weight_ih", (4 * H, H)), ("weight_hh", (4 * H, H)), ("bias_ih", (4 * H,)), ("bias_hh", (4 * H,))):
            sd[f"p.lstm.{n}_l{l}"] = torch.randn(*shape) * 0.3
    m = oenc.EncodecOracle(cfg, sd)
    x = torch.randn(2, H, 9)
    y = m.slstm("p", x)
    seq = x.permute(2, 0, 1)
    inp = seq
    for l in range(2):
        h = torch.zeros(2, H); c = torch.zeros(2, H); outs = []
        for t in range(9):
            gates = inp[t] @ sd[f"p.lstm.weight_ih_l{l}"].t() + sd[f"p.lstm.bias_ih_l{l}"] + h @ sd[f"p.lstm.weight_hh_l{l}"].t() + sd[f"p.lstm.bias_hh_l{l}"]
            i, f, g, o = gates.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs)
    np.testing.assert_allclose(y.numpy(), (inp + seq).permute(1, 2, 0).numpy(), atol=2e-6)   # + skip (SLSTM.cs:52-55)


# ------------------------------------------------------------------ .ecdc container (no language model)
def test_bitpacker_known_answers():
    """Hand-derived from BitPacker.cs:60-110: LSB-first accumulator, low byte out first, Flush pads the last byte with zeros."""
    p = oenc.BitPacker(10)
    for v in (1, 2, 3):
        p.push(v)
    # 1 | 2<<10 | 3<<20 = 0x00300801 over 30 bits -> bytes 01 08 30 and the 6 remaining bits (0) flushed as 00
    assert p.flush() == bytes([0x01, 0x08, 0x30, 0x00])
    p = oenc.BitPacker(10)
    p.push(1023)
    assert p.flush() == bytes([0xFF, 0x03])
    p = oenc.BitPacker(10)
    for v in (1023, 0, 1023, 512):
        p.push(v)                                       # exactly 40 bits -> 5 bytes, nothing left to flush
    # bits 0-9 and 20-29 set, bit 39 set
    assert p.flush() == bytes([0xFF, 0x03, 0xF0, 0x3F, 0x80])
    p = oenc.BitPacker(3)
    for v in (5, 7, 1):
        p.push(v)                                       # 101 | 111<<3 | 001<<6 = 0b0_0111_1101 -> 7D, then bit 8 = 0
    assert p.flush() == bytes([0x7D, 0x00])
    with pytest.raises(ValueError):
        oenc.BitPacker(0)


def test_bitunpacker_inverts_packer_and_ends_cleanly():
    rng = np.random.default_rng(5)
    for bits in (1, 3, 8, 10, 11, 16):
        vals = rng.integers(0, 1 << bits, size=257).tolist()
        p = oenc.BitPacker(bits)
        for v in vals:
            p.push(v)
        data = p.flush()
        assert len(data) == (len(vals) * bits + 7) // 8
        u = oenc.BitUnpacker(bits, data)
        assert [u.pull() for _ in vals] == vals
        tail = u.pull()                                 # only flush padding can remain: all-zero bits or end of stream
        assert tail in (None, 0)
    assert oenc.BitUnpacker(10, b"\xff").pull() is None  # BitUnpacker.cs:66-69 -> "Stream ended too soon" upstream


def test_ecdc_header_known_answer_and_validation():
    h = oenc.ecdc_header("encodec_24khz", 240000, 8, False, 1, 24000, 6.0)
    js = b'{"m":"encodec_24khz","al":240000,"nc":8,"lm":false,"ch":1,"sr":24000,"bw":6}'
    assert h == b"ECDC\x00" + len(js).to_bytes(4, "big") + js
    assert oenc.ecdc_header("encodec_24khz", 1, 2, False, 1, 24000, 1.5).endswith(b'"bw":1.5}')
    meta, off = oenc.ecdc_read_header(h + b"\x01\x02")
    assert off == len(h) and meta == {"m": "encodec_24khz", "al": 240000, "nc": 8, "lm": False, "ch": 1, "sr": 24000, "bw": 6}
    with pytest.raises(ValueError, match="not in ECDC format"):
        oenc.ecdc_read_header(b"RIFF" + h[4:])
    with pytest.raises(ValueError, match="Version not supported"):
        oenc.ecdc_read_header(b"ECDC\x01" + h[5:])
    with pytest.raises(EOFError):
        oenc.ecdc_read_header(h[:20])
    bad = b'{"m":"x","al":1,"lm":false}'
    with pytest.raises(ValueError, match="Missing required metadata key: nc"):
        oenc.ecdc_read_header(b"ECDC\x00" + len(bad).to_bytes(4, "big") + bad)


def test_ecdc_codes_round_trip_and_size():
    cfg = oenc.EncodecConfig()
    rng = np.random.default_rng(11)
    for nq, length in ((8, 240000), (2, 12345), (32, 321)):
        T = math.ceil(length / cfg.hop_length)
        codes = rng.integers(0, cfg.codebook_size, size=(nq, T))
        data = oenc.ecdc_compress_codes(cfg, codes, length, 6.0)
        meta, off = oenc.ecdc_read_header(data)
        assert len(data) - off == (nq * T * 10 + 7) // 8 and meta["al"] == length and meta["nc"] == nq
        back, _ = oenc.ecdc_decompress_codes(cfg, data)
        assert np.array_equal(back, codes)
        with pytest.raises(EOFError, match="Stream ended too soon"):
            oenc.ecdc_decompress_codes(cfg, data[:-2])


def test_c_abi_header_parser_matches_oracle():
    """nc_encodec_ecdc_info is host-only: runs without a GPU."""
    import neuralcodecs_b200 as nc
    from neuralcodecs_b200 import CodecException
    h = oenc.ecdc_header("encodec_24khz", 72000, 4, False, 1, 24000, 3.0)
    info = nc.EncodecCompressor.ReadHeader(h + b"\x00" * 8)
    assert info == {"al": 72000, "nc": 4, "ch": 1, "sr": 24000, "bw": 3.0, "lm": False, "payload_offset": len(h)}
    # Python json.dumps spacing (facebookresearch/encodec streams), missing optional keys -> reference defaults
    js = b'{"m": "encodec_48khz", "al": 10, "nc": 2, "lm": true}'
    info = nc.EncodecCompressor.ReadHeader(b"ECDC\x00" + len(js).to_bytes(4, "big") + js)
    assert info["sr"] == 48000 and info["ch"] == 1 and info["lm"] is True and info["bw"] == 0.0
    for bad in (b"RIFF\x00\x00\x00\x00\x02{}", b"ECDC\x07\x00\x00\x00\x02{}", b"ECDC\x00\x00\x00\x00\x02{}", b"ECDC\x00\x00\x00\x01\x00{"):
        with pytest.raises((ValueError, CodecException)):
            nc.EncodecCompressor.ReadHeader(bad)


def test_resampler_restatement_matches_scalar_loop_and_known_answers():
    """oracle.snac.resample_linear (vectorised) == the reference's scalar loop (SNAC.cs:284-308); hand-derived values."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal(211).astype(np.float32)
    for src, dst in ((44100, 24000), (8000, 24000), (3, 7), (7, 3), (5, 5)):
        np.testing.assert_array_equal(osnac.resample_linear(x, src, dst), osnac.resample_linear_loop(x, src, dst))
    np.testing.assert_array_equal(osnac.resample_linear(np.array([0.0, 1.0, 2.0], np.float32), 1, 2),
                                  np.array([0.0, 0.5, 1.0, 1.5, 2.0, 2.0], np.float32))      # last sample held
    assert osnac.resample_linear(np.array([1.0, 2.0, 3.0, 4.0], np.float32), 2, 1).tolist() == [1.0, 3.0]
    assert osnac.convert_to_mono(np.array([1, 3, 5, 7, 9], np.float32), 2).tolist() == [2.0, 6.0]   # ragged tail dropped


# ------------------------------------------------------------------------------------------------ Encodec 48 kHz preset
GOLD48 = os.path.join(os.path.dirname(__file__), "golden", "encodec48_hf_small.npz")


@pytest.fixture(scope="module")
def encodec48_gold():
    g = np.load(GOLD48)
    cfg = oenc.EncodecConfig(sample_rate=48000, channels=2, num_filters=8, hidden_size=32, codebook_size=64,
                             upsampling_ratios=[4, 3, 2], target_bandwidths=[24.0, 48.0, 96.0], bandwidth=48.0, causal=False,
                             normalize=True, norm_type="time_group_norm", chunk_length_s=0.05, overlap=0.01)
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
    return g, cfg, oenc.EncodecOracle(cfg, sd)


def test_encodec48_preset_algebra():
    c = oenc.EncodecConfig.encodec_48khz()          # EncodecConfig.cs:37-66, Encodec.cs:70-71,86,190-196
    assert (c.segment_length, c.segment_stride, c.frame_rate, c.num_quantizers, c.n_q_for_bandwidth()) == (48000, 47520, 150, 16, 4)
    assert [c.n_q_for_bandwidth(b) for b in (3.0, 12.0, 24.0)] == [2, 8, 16]
    assert oenc.EncodecConfig().segment_length is None and oenc.EncodecConfig().segment_stride is None


def test_encodec48_frames_match_hf_fixture(encodec48_gold):
    """Per-frame encode (scale, latents, codes) and decode of a full 100-frame segment and of a ragged 1013-sample tail
    against transformers.EncodecModel._encode_frame / _decode_frame: pins the non-causal padding, GroupNorm after every conv
    (before the transposed convs' trim), stereo in/out and the loudness scale."""
    g, cfg, m = encodec48_gold
    x = torch.from_numpy(g["audio_in"])
    seg = cfg.segment_length
    assert (seg, cfg.segment_stride) == (2400, int(g["stride"]))
    for name, fr in (("full", x[..., :seg]), ("tail", x[..., seg:])):
        with torch.inference_mode():
            codes, scale = m.encode_frame(fr, 48.0)
            emb = m.encoder(fr / scale.view(-1, 1, 1))
            audio = m.decode_frame(torch.from_numpy(g[f"{name}_codes"]), torch.from_numpy(g[f"{name}_scale"]))
        np.testing.assert_allclose(scale.numpy(), g[f"{name}_scale"], rtol=1e-6)
        np.testing.assert_allclose(emb.numpy(), g[f"{name}_emb"], atol=5e-6)
        np.testing.assert_array_equal(codes.numpy(), g[f"{name}_codes"])
        np.testing.assert_allclose(audio.numpy(), g[f"{name}_audio"], atol=5e-6)


def test_encodec48_overlap_add_matches_hf_and_hand(encodec48_gold):
    g, cfg, _ = encodec48_gold
    full, tail = torch.from_numpy(g["full_audio"]), torch.from_numpy(g["tail_audio"])
    ola = oenc.linear_overlap_add([full, full * 0.5, tail], cfg.segment_stride)
    np.testing.assert_allclose(ola.numpy(), g["ola"], atol=1e-7)
    # hand case: two length-4 frames at stride 3: weights 0.5 - |{.2,.4,.6,.8} - .5| = {.2,.4,.4,.2}
    a, b = torch.tensor([[1.0, 1.0, 1.0, 1.0]]), torch.tensor([[3.0, 3.0, 3.0, 3.0]])
    out = oenc.linear_overlap_add([a, b], 3)
    np.testing.assert_allclose(out.numpy(), [[1, 1, 1, (0.2 * 1 + 0.2 * 3) / 0.4, 3, 3, 3]], rtol=1e-6)
    # a last frame too short to cover the previous one's end: the reference's narrow() throws (AudioTensorDSP.cs:214)
    with pytest.raises(RuntimeError):
        oenc.linear_overlap_add([torch.ones(1, 8), torch.ones(1, 8), torch.ones(1, 2)], 3)


def test_encodec48_segment_loop_is_the_references(encodec48_gold):
    """Encodec.Encode (Encodec.cs:273-282): offsets 0, stride, ... while offset < length; frames end at min(offset+segment, length)."""
    g, cfg, m = encodec48_gold
    x = torch.from_numpy(g["audio_in"])                     # 3413 samples: offsets 0 and 2376 -> frames of 2400 and 1037 samples
    frames = m.encode_frames(x, 48.0)
    assert [f[0].shape[-1] for f in frames] == [100, 44] and all(f[1].shape == (2, 1) for f in frames)
    np.testing.assert_array_equal(frames[0][0].numpy(), g["full_codes"])
    y = m.forward_frames(x, 48.0)["audio"]
    assert y.shape == x.shape
    # exact multiple of the stride plus the overlap: the reference emits one more (24-sample) frame than HF's loop; its
    # 24 samples reach the last conv as ONE frame, whose Pad1d short-input branch (SConv1d.cs:258-272) lengthens it to 4
    assert [f[0].shape[-1] for f in m.encode_frames(x[..., :2400], 48.0)] == [100, 4]
    ref = oenc.linear_overlap_add([m.decode_frame(*f) for f in frames], cfg.segment_stride)[..., :x.shape[-1]]
    np.testing.assert_allclose(y.numpy(), ref.numpy(), atol=1e-6)


def test_ecdc_frames_with_scale_blocks_known_answer_and_round_trip():
    """Segmented stream (EncodecCompressor.cs:116-190,303-400): per frame [int32 BE 1][float32 BE scale][codes packed on their own]."""
    import struct
    cfg = oenc.EncodecConfig.encodec_48khz()
    rng = np.random.default_rng(0)
    al = 2 * 47520 + 12345
    frames = [(rng.integers(0, 1024, (4, t)), np.float32(0.25 * (i + 1))) for i, t in enumerate((150, 150, 39))]
    data = oenc.ecdc_compress_frames(cfg, frames, al, 6.0)
    meta, off = oenc.ecdc_read_header(data)
    assert meta == {"m": "encodec_48khz", "al": al, "nc": 4, "lm": False, "ch": 2, "sr": 48000, "bw": 6}
    per = [8 + (4 * t * 10 + 7) // 8 for t in (150, 150, 39)]
    assert len(data) == off + sum(per)
    assert data[off:off + 8] == struct.pack(">if", 1, 0.25) and data[off + per[0]:off + per[0] + 8] == struct.pack(">if", 1, 0.5)
    # first code bytes of frame 0: values (t0,k0), (t0,k1) at 10 bits each, LSB first
    v0, v1 = int(frames[0][0][0, 0]), int(frames[0][0][1, 0])
    assert data[off + 8] == v0 & 0xFF and data[off + 9] == ((v0 >> 8) | (v1 << 2)) & 0xFF
    back, meta2 = oenc.ecdc_decompress_frames(cfg, data)
    assert meta2 == meta and [f[0].shape for f in back] == [(4, 150), (4, 150), (4, 39)]
    for (c, s), (c2, s2) in zip(frames, back):
        np.testing.assert_array_equal(c, c2)
        assert s2.shape == (1,) and s2[0] == s
    with pytest.raises(EOFError):
        oenc.ecdc_decompress_frames(cfg, data[:-3])
    bad = bytearray(data)
    bad[off:off + 4] = struct.pack(">i", 0)
    with pytest.raises(ValueError, match="Invalid scale count"):
        oenc.ecdc_decompress_frames(cfg, bytes(bad))
