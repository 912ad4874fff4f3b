"""Encodec engine (through the C ABI) against the CPU oracle (24 kHz mono causal preset, 6 kbps)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MAX_ABS, MIN_SNR_DB, NEAR_TIE = 1e-3, 60.0, 1e-6     # north_star tolerances


def snr_db(ref, test):
    ref, test = np.asarray(ref, np.float64), np.asarray(test, np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - test) ** 2).sum(), 1e-300))


def _models(fix, options=None):
    import neuralcodecs_b200 as nc
    from oracle import encodec as oenc
    co, ce, path = fix
    o = oenc.load_safetensors(path, co)
    m = nc.Encodec(ce, options=options)
    m.LoadWeights(path)
    return o, m


def _bad_flips(o, emb, codes_ref, codes):
    """un-cascaded code flips whose scale-normalised margin is not a near-tie"""
    bad = 0
    ct = torch.from_numpy(codes)
    with torch.inference_mode():
        residual = emb.clone()
        tainted = torch.zeros(codes_ref.shape[0], codes_ref.shape[2], dtype=torch.bool)
        for q in range(codes_ref.shape[1]):
            flat = residual.transpose(1, 2).reshape(-1, residual.shape[1])
            dist = o.vq_distances(q, flat).reshape(codes_ref.shape[0], codes_ref.shape[2], -1)
            embed = o.sd[f"quantizer.layers.{q}.codebook.embed"]
            new = (codes_ref[:, q] != ct[:, q]) & ~tainted
            for b, t in new.nonzero().tolist():
                scale = float(residual[b, :, t].pow(2).sum() + embed[ct[b, q, t]].pow(2).sum())
                margin = float(dist[b, t, ct[b, q, t]] - dist[b, t, codes_ref[b, q, t]]) / max(scale, 1e-30)
                bad += abs(margin) >= NEAR_TIE
            tainted |= codes_ref[:, q] != ct[:, q]
            quant, _ = o.vq_forward(q, residual)
            residual = residual - quant
    return bad


def _run(fix, options, batch, length, first=9):
    from oracle import synth
    o, m = _models(fix, options)
    x = synth.synth_audio(batch, length, fix[0].sample_rate, first_clip=first)[:, None, :]
    xt = torch.from_numpy(x)
    ref = o.forward(xt)
    emb = o.encode_latent(xt)
    (codes, scale), = m.Encode(x)
    assert scale is None and codes.dtype == np.int64 and codes.shape == tuple(ref["codes"].shape)
    dec = m.Decode([(ref["codes"].numpy(), None)])
    dref = o.decode(ref["codes"]).numpy()
    assert dec.shape == dref.shape
    return o, m, x, ref, emb, codes, dec, dref


@pytest.mark.parametrize("length", [24000, 12345])
def test_conv_stacks_and_vq_without_lstm_fp32(encodec_nolstm, length):
    o, m, x, ref, emb, codes, dec, dref = _run(encodec_nolstm, {"precision": "fp32"}, 2, length)
    assert _bad_flips(o, emb, ref["codes"], codes) == 0
    np.testing.assert_allclose(dec, dref, atol=5e-6)
    m.Dispose()


def test_conv_stacks_without_lstm_tensor_core(encodec_nolstm):
    o, m, x, ref, emb, codes, dec, dref = _run(encodec_nolstm, None, 3, 30001)
    assert any(v.startswith("tcgen05") for v in m.describe()["layers"].values())
    assert _bad_flips(o, emb, ref["codes"], codes) == 0
    assert np.abs(dec - dref).max() <= MAX_ABS and snr_db(dref, dec) >= MIN_SNR_DB
    m.Dispose()


def test_encodec24k_preset_with_lstm(encodec_24k):
    """BASELINE config #3 shape: 10 s clips -> codes [B, 8, 750]; 8 clips = 6000 frames x 8 codebooks at the DEFAULT policy
    (encoder on tensor cores: 3xTF32 operands, short accumulation chains): every un-cascaded flip must be a near-tie."""
    o, m, x, ref, emb, codes, dec, dref = _run(encodec_24k, None, 8, 240000)
    d = m.describe()
    assert d["encoder_precision"] == "3xtf32" and any("folded" in v for v in d["layers"].values())
    assert codes.shape == (8, 8, 750) and dec.shape == (8, 1, 240000)
    match = float((codes == ref["codes"].numpy()).mean())
    print(f"encodec24k: code match {match:.5f}; decoder max-abs {np.abs(dec - dref).max():.2e} snr {snr_db(dref, dec):.1f} dB")
    assert _bad_flips(o, emb, ref["codes"], codes) == 0
    assert np.abs(dec - dref).max() <= MAX_ABS and snr_db(dref, dec) >= MIN_SNR_DB
    y = m.forward(x)
    assert y.shape == x.shape
    if match == 1.0:
        assert np.abs(y - ref["audio"].numpy()).max() <= MAX_ABS
    m.Dispose()


@pytest.mark.parametrize("length", [100, 319, 320, 321, 641, 1600, 1920, 1921, 2500])
def test_short_clips_take_pad1d_short_input_branch(encodec_24k, length):
    """Clips of fewer than 7 frames (<= 1920 samples): an SConv1d whose input is not longer than its reflect padding
    zero-extends it first and does NOT trim afterwards (Modules/Encodec/SConv1d.cs:258-272), so the frame count grows
    (e.g. 3 frames -> 7).  Shapes, codes and audio must follow the reference through that branch."""
    o, m, x, ref, emb, codes, dec, dref = _run(encodec_24k, None, 2, length)
    assert codes.shape == tuple(ref["codes"].shape) and dec.shape == dref.shape
    if length <= 1920:
        assert codes.shape[-1] >= 7 > -(-length // 320)                 # lengthened by the short-input branch
    assert _bad_flips(o, emb, ref["codes"], codes) == 0
    assert np.abs(dec - dref).max() <= MAX_ABS and snr_db(dref, dec) >= MIN_SNR_DB
    y = m.forward(x)
    assert y.shape == x.shape == tuple(ref["audio"].shape)
    if np.array_equal(codes, ref["codes"].numpy()):
        assert np.abs(y - ref["audio"].numpy()).max() <= MAX_ABS
    m.Dispose()


def test_errors(encodec_nolstm):
    import neuralcodecs_b200 as nc
    _, m = _models(encodec_nolstm, {"precision": "fp32"})
    with pytest.raises(ValueError, match="Expected 1 channels"):
        m.Encode(np.zeros((1, 2, 1000), np.float32))
    with pytest.raises(ValueError, match="No frames provided"):
        m.Decode([])
    with pytest.raises(ValueError, match="Invalid bandwidth"):
        nc.Encodec(nc.EncodecConfig(bandwidth=7.0))
    m.SetTargetBandwidth(3.0)
    (codes, _), = m.Encode(np.zeros((1, 1, 3200), np.float32))
    assert codes.shape == (1, 4, 10)                                     # 3 kbps -> 4 codebooks, ceil(3200/320) frames
    m.Dispose()


# ------------------------------------------------------------------ .ecdc container (no language model)
def test_ecdc_compress_is_byte_exact_against_oracle_packer(encodec_nolstm):
    """Header + bit-packed payload from the device equal the oracle's BinaryIO/BitPacker restatement applied to the codes
    the same engine's Encode returns (so encoder near-ties cannot blur a packing difference)."""
    import neuralcodecs_b200 as nc
    from oracle import encodec as oenc
    from oracle import synth
    co, _, _ = encodec_nolstm
    _, m = _models(encodec_nolstm, {"precision": "fp32"})
    for bw, length, batch in ((6.0, 24000, 3), (1.5, 12345, 2), (24.0, 2241, 1)):
        m.SetTargetBandwidth(bw)
        x = synth.synth_audio(batch, length, co.sample_rate, first_clip=3)[:, None, :]
        (codes, _), = m.Encode(x)
        streams = nc.EncodecCompressor.CompressBatch(m, x)
        assert len(streams) == batch
        for b in range(batch):
            want = oenc.ecdc_compress_codes(co, codes[b], length, bw)
            assert streams[b] == want, f"bw {bw} clip {b}: stream differs from the oracle"
            back, meta = oenc.ecdc_decompress_codes(co, streams[b])
            assert np.array_equal(back, codes[b]) and meta["al"] == length and meta["nc"] == codes.shape[1]
        assert nc.EncodecCompressor.Compress(m, x[0]) == streams[0]                       # [C, L] single-clip form
    m.Dispose()


def test_ecdc_decompress_round_trip(encodec_24k):
    """Decompress(Compress(x)) == forward(x) of the same engine, and == oracle decode of the unpacked codes (tolerance)."""
    import neuralcodecs_b200 as nc
    from oracle import encodec as oenc
    from oracle import synth
    co, _, _ = encodec_24k
    o, m = _models(encodec_24k)
    length = 48123
    x = synth.synth_audio(2, length, co.sample_rate, first_clip=21)[:, None, :]
    streams = nc.EncodecCompressor.CompressBatch(m, x)
    info = nc.EncodecCompressor.ReadHeader(streams[0])
    assert info["al"] == length and info["nc"] == 8 and info["sr"] == 24000 and info["bw"] == 6.0 and not info["lm"]
    wav, sr = nc.EncodecCompressor.DecompressBatch(streams, m)
    assert sr == 24000 and wav.shape == (2, 1, length)
    y = m.forward(x)
    np.testing.assert_array_equal(wav, y)                 # same kernels, same codes -> identical bits
    codes = np.stack([oenc.ecdc_decompress_codes(co, s)[0] for s in streams])
    dref = o.decode(torch.from_numpy(codes)).numpy()[..., :length]
    assert np.abs(wav - dref).max() <= MAX_ABS and snr_db(dref, wav) >= MIN_SNR_DB
    one, _ = nc.EncodecCompressor.Decompress(streams[1], m)
    np.testing.assert_array_equal(one, wav[1])
    m.Dispose()


def test_ecdc_oracle_written_stream_and_errors(encodec_nolstm):
    import neuralcodecs_b200 as nc
    from oracle import encodec as oenc
    co, _, _ = encodec_nolstm
    o, m = _models(encodec_nolstm, {"precision": "fp32"})
    rng = np.random.default_rng(2)
    length, nq = 6400, 16
    codes = rng.integers(0, 1024, size=(nq, 20))
    data = oenc.ecdc_compress_codes(co, codes, length, 12.0)
    wav, sr = nc.EncodecCompressor.Decompress(data, m)
    dref = o.decode(torch.from_numpy(codes[None])).numpy()[0, :, :length]
    np.testing.assert_allclose(wav, dref, atol=5e-6)
    with pytest.raises((ValueError, nc.CodecException), match="Stream ended too soon"):
        nc.EncodecCompressor.Decompress(data[:-3], m)
    lm = oenc.ecdc_header("encodec_24khz", length, nq, True, 1, 24000, 12.0) + data[oenc.ecdc_read_header(data)[1]:]
    with pytest.raises((RuntimeError, nc.CodecException), match="language-model"):
        nc.EncodecCompressor.Decompress(lm, m)
    stereo = oenc.ecdc_header("encodec_24khz", length, nq, False, 2, 24000, 12.0) + data[oenc.ecdc_read_header(data)[1]:]
    with pytest.raises((ValueError, nc.CodecException), match="channels"):
        nc.EncodecCompressor.Decompress(stereo, m)
    with pytest.raises(ValueError, match="shape should be"):
        nc.EncodecCompressor.Compress(m, np.zeros(100, np.float32))
    m.Dispose()
