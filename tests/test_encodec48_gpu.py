"""Encodec 48 kHz preset through the C ABI against the CPU oracle: stereo, non-causal reflect padding, GroupNorm(1, C) after
every conv, 1 s segments with 1 % overlap, per-segment loudness scale, triangular overlap-add
(Config/Encodec/EncodecConfig.cs:37-66, Models/Encodec.cs:213-296,436-489, AudioTools/AudioTensorDSP.cs:161-261)."""
import numpy as np
import pytest
import torch

from test_encodec_gpu import MAX_ABS, MIN_SNR_DB, _bad_flips, _models, snr_db

pytestmark = pytest.mark.gpu

SEG, STRIDE = 48000, 47520


def _stereo(batch, length, first=3):
    from oracle import synth
    x = synth.synth_audio(2 * batch, length, 48000, first_clip=first).reshape(batch, 2, length)
    return np.ascontiguousarray(x * np.linspace(0.2, 1.5, batch, dtype=np.float32).reshape(batch, 1, 1))


def _latents(o, xt):
    """the encoder output of every segment, as the quantiser sees it"""
    out = []
    with torch.inference_mode():
        L = xt.shape[-1]
        seg, stride = o.cfg.segment_length or L, o.cfg.segment_stride or L
        for off in range(0, L, stride):
            fr = xt[:, :, off:min(off + seg, L)]
            if o.cfg.normalize:
                fr = fr / (fr.mean([1], keepdim=True).pow(2).mean([2], keepdim=True).sqrt() + 1e-8)
            out.append(o.encoder(fr))
    return out


def _check_frames(o, m, x, expect_frames):
    xt = torch.from_numpy(x)
    ref = o.encode_frames(xt)
    assert [f[0].shape[-1] for f in ref] == expect_frames
    got = m.Encode(x)
    assert len(got) == len(ref)
    lat = _latents(o, xt)
    for (c, s), (cr, sr), emb in zip(got, ref, lat):
        assert c.dtype == np.int64 and c.shape == tuple(cr.shape)
        if sr is None:
            assert s is None
        else:
            assert s.shape == tuple(sr.shape)
            np.testing.assert_allclose(s, sr.numpy(), rtol=2e-6)
        assert _bad_flips(o, emb, cr, c) == 0
    # Decode of the oracle's frames (identical codes and scales on both sides)
    dref = o.decode_frames(ref).numpy()
    dec = m.Decode([(c.numpy(), None if s is None else s.numpy()) for c, s in ref])
    assert dec.shape == dref.shape
    err = np.abs(dec - dref).max()
    print(f"frames {expect_frames}: code match {np.mean([float((c == cr.numpy()).mean()) for (c, _), (cr, _) in zip(got, ref)]):.5f}; "
          f"decode max-abs {err:.2e} snr {snr_db(dref, dec):.1f} dB")
    assert err <= MAX_ABS * max(1.0, float(np.abs(dref).max())) and snr_db(dref, dec) >= MIN_SNR_DB
    return got, ref, dec, dref


def test_three_segments_with_ragged_tail(encodec_48k):
    """2 full segments + a 12345-sample tail, batch 3 with different loudness per clip: scales, codes (near-tie gate),
    Decode (frames * scale, overlap-add) and forward sliced to the input length."""
    o, m = _models(encodec_48k)
    d = m.describe()
    assert d.get("norm") == "time_group_norm" and d["segment_length"] == SEG and d["segment_stride"] == STRIDE
    x = _stereo(3, 2 * STRIDE + 12345)
    got, ref, dec, dref = _check_frames(o, m, x, [150, 150, 39])
    assert dec.shape == (3, 2, 2 * STRIDE + 39 * 320)
    y = m.forward(x)
    assert y.shape == x.shape
    if all(np.array_equal(c, cr.numpy()) for (c, _), (cr, _) in zip(got, ref)):
        yr = o.forward_frames(torch.from_numpy(x))["audio"].numpy()
        assert np.abs(y - yr).max() <= MAX_ABS * max(1.0, float(np.abs(yr).max()))
    m.Dispose()


def test_truncated_penultimate_segment_and_single_short_frame(encodec_48k):
    o, m = _models(encodec_48k)
    # 47520 + 900 samples: segment 0 is full, segment 1 has 900 samples = 3 frames, which the last encoder conv's Pad1d
    # short-input branch lengthens to 4 (SConv1d.cs:258-272)
    _check_frames(o, m, _stereo(2, STRIDE + 900, first=5), [150, 4])
    # 47900 samples: segment 0 = the whole clip (shorter than a segment), segment 1 = its last 380 samples
    _check_frames(o, m, _stereo(1, 47900, first=6), [150, 4])
    # one short frame: the triangular weights cancel (f * w / w)
    _check_frames(o, m, _stereo(2, 10000, first=7), [32])
    m.Dispose()


def test_last_frame_too_short_for_overlap_add_is_an_error(encodec_48k):
    """With 50 % overlap a 4-frame last segment (1280 samples at offset 24000) ends before segment 0's 48000 samples: the
    reference's narrow() throws in LinearOverlapAdd (AudioTensorDSP.cs:214)."""
    import copy
    import neuralcodecs_b200 as nc
    co, ce, path = encodec_48k
    co, ce = copy.deepcopy(co), copy.deepcopy(ce)
    co.overlap = ce.overlap = 0.5
    o, m = _models((co, ce, path))
    assert m.describe()["segment_stride"] == 24000
    rng = np.random.default_rng(3)
    frames = [(rng.integers(0, 1024, (1, 4, t), dtype=np.int64), np.ones((1, 1), np.float32)) for t in (150, 4)]
    with pytest.raises(RuntimeError):
        o.decode_frames([(torch.from_numpy(c), torch.from_numpy(s)) for c, s in frames])
    with pytest.raises(ValueError, match="overlap-add"):
        m.Decode(frames)
    # three segments cover every sample twice: 48000 + 24000 samples -> offsets 0, 24000, 48000
    _check_frames(o, m, _stereo(1, 72000, first=8), [150, 150, 75])
    m.Dispose()


def test_one_unnormalised_frame_per_clip(encodec_48k_one_frame):
    """Same architecture with Segment = null and Normalize = false: one frame, no scale, no overlap-add (Encodec.cs:220-228)."""
    o, m = _models(encodec_48k_one_frame)
    x = _stereo(2, 30001, first=11)
    got, ref, dec, dref = _check_frames(o, m, x, [94])
    assert got[0][1] is None and dec.shape == (2, 2, 94 * 320)
    with pytest.raises(ValueError, match="single frame"):
        m.Decode([got[0], got[0]])
    m.Dispose()


def test_errors_and_unsupported(encodec_48k):
    import neuralcodecs_b200 as nc
    o, m = _models(encodec_48k)
    with pytest.raises(ValueError, match="channels"):
        m.Encode(np.zeros((1, 1, 1000), np.float32))
    with pytest.raises(ValueError, match="encode_frames"):
        from neuralcodecs_b200 import _lib
        import ctypes as C
        x = _stereo(1, 5000)
        codes = np.zeros((1, 4, 16), np.int64)
        _lib.check(_lib.lib().nc_encodec_encode(m._handle(), x.ctypes.data_as(C.c_void_p), 1, 5000, 6.0, codes.ctypes.data_as(C.c_void_p)))
    cfg = nc.EncodecConfig.Encodec48Khz()
    cfg.use_causal_conv = True
    with pytest.raises(ValueError, match="causal"):
        nc.Encodec(cfg)                                              # NormConv1d.cs:143-147
    cfg = nc.EncodecConfig.Encodec48Khz()
    cfg.channels = 3
    with pytest.raises(ValueError, match="channels"):
        nc.Encodec(cfg)                                              # Encodec.cs:268-271
    m.Dispose()


def test_ecdc_streams_with_scale_blocks(encodec_48k):
    """EncodecCompressor without the language model on the segmented model: the stream is byte-exact against the oracle's
    writer over the same frames; Decompress = Decode of the frames the READER's length formula yields, trimmed to `al`."""
    import neuralcodecs_b200 as nc
    from oracle import encodec as oenc
    o, m = _models(encodec_48k)
    L = 2 * STRIDE + 12345
    x = _stereo(2, L, first=13)
    frames = m.Encode(x)
    streams = nc.EncodecCompressor.CompressBatch(m, x)
    for b, st in enumerate(streams):
        want = oenc.ecdc_compress_frames(o.cfg, [(c[b], s[b]) for c, s in frames], L, 6.0)
        assert st == want
    assert nc.EncodecCompressor.Compress(m, x[1]) == streams[1]
    hdr = nc.EncodecCompressor.ReadHeader(streams[0])
    assert (hdr["al"], hdr["nc"], hdr["ch"], hdr["sr"], hdr["lm"]) == (L, 4, 2, 48000, False)
    wav, sr = nc.EncodecCompressor.DecompressBatch(streams, m)
    assert sr == 48000 and wav.shape == (2, 2, L)
    dec = m.Decode(frames)[..., :L]
    np.testing.assert_allclose(wav, dec, atol=1e-6)
    one, _ = nc.EncodecCompressor.Decompress(streams[1], m)
    np.testing.assert_allclose(one, dec[1], atol=1e-6)
    # an oracle-written stream of the ORACLE's frames decodes to the oracle's audio
    ref = o.encode_frames(torch.from_numpy(x[:1]))
    st = oenc.ecdc_compress_frames(o.cfg, [(c[0].numpy(), s[0].numpy()) for c, s in ref], L, 6.0)
    wav1, _ = nc.EncodecCompressor.Decompress(st, m)
    yr = o.decode_frames(ref).numpy()[0, :, :L]
    assert np.abs(wav1 - yr).max() <= MAX_ABS * max(1.0, float(np.abs(yr).max()))
    # a 900-sample tail: the writer stores the encoder's 4 frames, the reader takes ceil(900*150/48000) = 3 (mirrored as is)
    x2 = _stereo(1, STRIDE + 900, first=14)
    st2 = nc.EncodecCompressor.Compress(m, x2[0])
    f2, _ = oenc.ecdc_decompress_frames(o.cfg, st2)
    assert [f[0].shape[1] for f in f2] == [150, 3]
    wav2, _ = nc.EncodecCompressor.Decompress(st2, m)
    want2 = m.Decode([(f[0][None], f[1].reshape(1, 1)) for f in f2])[0, :, :STRIDE + 900]
    np.testing.assert_allclose(wav2, want2, atol=1e-6)
    # truncated stream / bad scale count
    with pytest.raises(ValueError, match="Stream ended too soon"):
        nc.EncodecCompressor.Decompress(streams[0][:-5], m)
    bad = bytearray(streams[0])
    off = hdr["payload_offset"]
    bad[off:off + 4] = (0).to_bytes(4, "big")
    with pytest.raises(ValueError, match="Invalid scale count"):
        nc.EncodecCompressor.Decompress(bytes(bad), m)
    m.Dispose()


def test_full_size_properties_thirty_second_stereo_clips(encodec_48k):
    """Size-independent properties at a realistic size (4 stereo clips x 30 s = 31 segments each, 124 batch items), where the
    oracle would take minutes: (i) the loudness normalisation makes the codes invariant to the input gain and the scales
    proportional to it; (ii) a clip's frames do not depend on its batch neighbours; (iii) Decode of the frames = forward;
    (iv) segment s of a long clip = the encoding of that segment alone."""
    o, m = _models(encodec_48k)
    L = 30 * 48000
    x = _stereo(4, L, first=21)
    seg, nq, total = m.query_frames(L)
    assert len(seg) == 31 and seg[:-1] == [150] * 30 and nq == 4 and total == 30 * STRIDE + seg[-1] * 320
    f1 = m.Encode(x)
    f2 = m.Encode(4.0 * x)                                            # power-of-two gain: x / scale is the same up to the 1e-8 offset
    assert len(f1) == 31
    match = np.mean([float((a[0] == b[0]).mean()) for a, b in zip(f1, f2)])
    assert match >= 0.9995
    for (_, s1), (_, s2) in zip(f1, f2):
        np.testing.assert_allclose(s2, 4.0 * s1, rtol=1e-6)
    solo = m.Encode(x[2:3])
    for (c, s), (cs, ss) in zip(f1, solo):
        np.testing.assert_array_equal(c[2:3], cs)
        np.testing.assert_array_equal(s[2:3], ss)
    y = m.forward(x)
    dec = m.Decode(f1)
    assert y.shape == x.shape and dec.shape == (4, 2, total)
    np.testing.assert_allclose(y, dec[..., :L], atol=1e-6)
    assert np.isfinite(y).all() and snr_db(x, y) > -20.0              # random weights: only sanity on the level
    s = 17
    alone = m.Encode(np.ascontiguousarray(x[:, :, s * STRIDE:s * STRIDE + SEG]))
    assert len(alone) == 2                                            # the 48000-sample excerpt is itself cut into 47520 + 480
    one = m.Encode(np.ascontiguousarray(x[:, :, s * STRIDE:s * STRIDE + SEG - 480 + 480]))[0]
    np.testing.assert_array_equal(one[0], f1[s][0])
    np.testing.assert_allclose(one[1], f1[s][1], rtol=1e-6)
    m.Dispose()
