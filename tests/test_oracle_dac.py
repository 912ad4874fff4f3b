"""The DAC oracle against (i) committed HF DacModel fixtures and (ii) naive numpy restatements
of each op on small shapes.  CPU only."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import dac as odac
from oracle import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dac_hf_small.npz")


@pytest.fixture(scope="module")
def small():
    cfg = odac.DACConfig(sample_rate=16000, encoder_dim=16, decoder_dim=128, n_codebooks=4, codebook_size=64)
    sd = synth.make_dac_weights_hf(cfg, codebooks="normal")
    model = odac.DACOracle(cfg, odac.convert_hf_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}))
    return cfg, model, np.load(GOLD)


def test_synth_audio_matches_fixture(small):
    _, _, g = small
    x = synth.synth_audio(2, 5000, 16000)
    np.testing.assert_array_equal(x, g["audio_in"][:, 0, :])


def test_encoder_matches_hf_fixture(small):
    _, m, g = small
    with torch.inference_mode():
        z = m.encoder(m.preprocess(torch.from_numpy(g["audio_in"])))
    assert z.shape == g["encoder_out"].shape
    np.testing.assert_allclose(z.numpy(), g["encoder_out"], atol=2e-5, rtol=0)


def test_decoder_matches_hf_fixture(small):
    _, m, g = small
    a = m.decode(torch.from_numpy(g["encoder_out"]))
    assert a.shape == g["decoder_out"].shape          # padded length, not trimmed (DAC.cs:231-234)
    np.testing.assert_allclose(a.numpy(), g["decoder_out"], atol=5e-6, rtol=0)


def test_from_codes_matches_hf_fixture(small):
    _, m, g = small
    z = m.from_codes(torch.from_numpy(g["codes"]))
    np.testing.assert_allclose(z.numpy(), g["from_codes"], atol=5e-6, rtol=0)


def test_preprocess_padding_and_rate_check(small):
    cfg, m, _ = small
    for L in (1, cfg.hop_length - 1, cfg.hop_length, cfg.hop_length + 1, 5000):
        x = torch.ones(1, 1, L)
        p = m.preprocess(x)
        assert p.shape[-1] == math.ceil(L / cfg.hop_length) * cfg.hop_length
        assert float(p[..., L:].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        m.preprocess(torch.ones(1, 1, 10), sample_rate=8000)


def _naive_conv1d(x, w, b, stride, pad, dil):
    B, Cin, T = x.shape
    Cout, _, K = w.shape
    Tout = (T + 2 * pad - dil * (K - 1) - 1) // stride + 1
    y = np.zeros((B, Cout, Tout), np.float64)
    for t in range(Tout):
        for j in range(K):
            ti = t * stride + j * dil - pad
            if 0 <= ti < T:
                y[:, :, t] += x[:, :, ti].astype(np.float64) @ w[:, :, j].astype(np.float64).T
    return y + b[None, :, None]


def _naive_convt1d(x, w, b, stride, pad):
    B, Cin, T = x.shape
    _, Cout, K = w.shape
    Tout = (T - 1) * stride - 2 * pad + K
    y = np.zeros((B, Cout, Tout), np.float64)
    for i in range(T):
        for j in range(K):
            to = i * stride - pad + j
            if 0 <= to < Tout:
                y[:, :, to] += x[:, :, i].astype(np.float64) @ w[:, :, j].astype(np.float64)
    return y + b[None, :, None]


def _fold(v, g):
    n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
    return v / (n + 1e-7) * g


@pytest.mark.parametrize("stride,pad,dil,k", [(1, 3, 1, 7), (1, 9, 3, 7), (1, 27, 9, 7), (2, 1, 1, 4), (8, 4, 1, 16), (1, 0, 1, 1)])
def test_wnconv1d_against_naive(stride, pad, dil, k):
    rng = np.random.default_rng(0)
    v = rng.standard_normal((6, 5, k)).astype(np.float32)
    g = rng.uniform(0.5, 2, (6, 1, 1)).astype(np.float32)
    b = rng.standard_normal(6).astype(np.float32)
    x = rng.standard_normal((2, 5, 64)).astype(np.float32)
    m = odac.DACOracle(odac.DACConfig(), {"c.weight_v": torch.from_numpy(v), "c.weight_g": torch.from_numpy(g),
                                          "c.bias": torch.from_numpy(b)})
    y = m.wnconv1d("c", torch.from_numpy(x), stride=stride, padding=pad, dilation=dil).numpy()
    np.testing.assert_allclose(y, _naive_conv1d(x, _fold(v, g), b, stride, pad, dil), atol=2e-5)


@pytest.mark.parametrize("stride", [2, 4, 8, 5])
def test_wnconvtranspose1d_against_naive(stride):
    rng = np.random.default_rng(1)
    k, pad = 2 * stride, math.ceil(stride / 2)
    v = rng.standard_normal((5, 3, k)).astype(np.float32)
    g = rng.uniform(0.5, 2, (5, 1, 1)).astype(np.float32)     # per-IN-channel gain (WNConvTranspose1d.cs:146-150)
    b = rng.standard_normal(3).astype(np.float32)
    x = rng.standard_normal((2, 5, 17)).astype(np.float32)
    m = odac.DACOracle(odac.DACConfig(), {"c.weight_v": torch.from_numpy(v), "c.weight_g": torch.from_numpy(g),
                                          "c.bias": torch.from_numpy(b)})
    y = m.wnconvtranspose1d("c", torch.from_numpy(x), stride=stride, padding=pad).numpy()
    ref = _naive_convt1d(x, _fold(v, g), b, stride, pad)
    assert y.shape == ref.shape
    assert y.shape[-1] == (stride * 17 if stride % 2 == 0 else stride * 17 - 1)
    np.testing.assert_allclose(y, ref, atol=2e-5)


def test_snake_no_epsilon_and_alpha_zero():
    a = np.array([0.0, 0.5, 2.0], np.float32).reshape(1, 3, 1)
    x = np.linspace(-3, 3, 33, dtype=np.float32).reshape(1, 1, -1).repeat(3, axis=1)
    m = odac.DACOracle(odac.DACConfig(), {"s.alpha": torch.from_numpy(a)})
    y = m.snake("s", torch.from_numpy(x)).numpy()
    np.testing.assert_array_equal(y[0, 0], x[0, 0])                                  # alpha == 0 -> identity
    for c in (1, 2):
        ref = x[0, c].astype(np.float64) + np.sin(a[0, c, 0] * x[0, c].astype(np.float64)) ** 2 / a[0, c, 0]
        np.testing.assert_allclose(y[0, c], ref, atol=1e-6)


def test_vq_is_unnormalised_expanded_form_lowest_index_wins():
    cfg = odac.DACConfig(encoder_dim=8, encoder_rates=[2], n_codebooks=1, codebook_size=4, codebook_dim=8)
    D, Dz = 8, cfg.latent_dim
    rng = np.random.default_rng(3)
    cb = rng.standard_normal((4, D)).astype(np.float32)
    cb[2] = cb[1]                                   # exact tie between entries 1 and 2
    cb[3] = 10 * cb[0]                              # same direction as 0, far away: normalised VQ would tie 0/3
    sd = {"quantizer.quantizers.0.codebook.weight": torch.from_numpy(cb),
          "quantizer.quantizers.0.in_proj.weight_v": torch.eye(D, Dz).reshape(D, Dz, 1).contiguous(),
          "quantizer.quantizers.0.in_proj.weight_g": torch.ones(D, 1, 1),
          "quantizer.quantizers.0.in_proj.bias": torch.zeros(D),
          "quantizer.quantizers.0.out_proj.weight_v": torch.eye(Dz, D).reshape(Dz, D, 1).contiguous(),
          "quantizer.quantizers.0.out_proj.weight_g": torch.ones(Dz, 1, 1),
          "quantizer.quantizers.0.out_proj.bias": torch.zeros(Dz)}
    m = odac.DACOracle(cfg, sd)
    z = torch.zeros(1, Dz, 3)
    z[0, :D, 0] = torch.from_numpy(cb[1])           # frame 0 sits exactly on entries 1 and 2
    z[0, :D, 1] = torch.from_numpy(cb[0]) * 1.01    # frame 1 nearest (un-normalised) to entry 0, not 3
    z[0, :D, 2] = torch.from_numpy(cb[3]) * 0.99
    _, idx, _ = m.vq_forward(0, z)
    assert idx.dtype == torch.int64
    assert idx[0].tolist() == [1, 0, 3]
    e = z[0, :D, 1].double().numpy()
    d = ((e[None, :] - cb.astype(np.float64)) ** 2).sum(1)
    assert int(np.argmin(d)) == 0


def test_rvq_residual_bookkeeping_and_from_codes(small):
    cfg, m, _ = small
    torch.manual_seed(0)
    z = torch.randn(2, cfg.latent_dim, 7)
    zq, codes, latents = m.rvq_forward(z)
    assert codes.shape == (2, cfg.n_codebooks, 7) and codes.dtype == torch.int64
    assert latents.shape == (2, cfg.n_codebooks * cfg.codebook_dim, 7)
    # FromCodes has no straight-through arithmetic: equal up to the two fp32 roundings of :81
    np.testing.assert_allclose(m.from_codes(codes).numpy(), zq.numpy(), atol=1e-5)
    zq2, codes2, _ = m.rvq_forward(z, 2)
    assert codes2.shape[1] == 2
    np.testing.assert_array_equal(codes2.numpy(), codes[:, :2].numpy())


def test_dia_decode_shape(small):
    cfg, m, _ = small
    codes = torch.from_numpy(synth.dia_codes(1, 9, cfg.n_codebooks, cfg.codebook_size))[0]   # [T, nq]
    a = m.dia_decode(codes)
    assert a.shape == (9 * cfg.hop_length,)
