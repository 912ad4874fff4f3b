"""Seeded synthetic audio and random-init weights in the reference's safetensors layouts.

Used by bench.py, the tests and the oracle's data generator (oracle/synth.py adds the
data-fitted codebooks, which need the CPU oracle).  No arithmetic of the codec path lives
here.  Implements SURVEY.md section 8(d):
deterministic tonal+noise clips and per-architecture seeded weights written as
safetensors in the key layouts the reference loads
(DAC: HF ``DacModel`` layout consumed by
``Config/DAC/StateDictNameConverter.cs:40-65,274-376``).

numpy's PCG64 stream is stable across numpy versions, so the same bytes are
produced here and on the GPU box.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Optional

import numpy as np

AUDIO_SEED = 1234
WEIGHT_SEED = 4321


# --------------------------------------------------------------------------- audio
def synth_audio(batch: int, length: int, sample_rate: int, first_clip: int = 0) -> np.ndarray:
    """[batch, length] float32 clips; clip b depends only on (first_clip + b)."""
    out = np.empty((batch, length), dtype=np.float32)
    t = np.arange(length, dtype=np.float64) / float(sample_rate)
    for i in range(batch):
        b = first_clip + i
        f = 110.0 * 2.0 ** ((b % 48) / 12.0)
        phi = 0.37 * b
        rng = np.random.default_rng([AUDIO_SEED, b])
        x = 0.30 * np.sin(2 * np.pi * f * t + phi) + 0.15 * np.sin(2 * np.pi * 3.1 * f * t)
        x += 0.05 * rng.standard_normal(length)
        out[i] = np.clip(x, -1.0, 1.0).astype(np.float32)
    return out


# --------------------------------------------------------------------------- weights
def _rng(name: str, seed: int = WEIGHT_SEED) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def _uniform(name: str, shape, bound: float) -> np.ndarray:
    return _rng(name).uniform(-bound, bound, size=shape).astype(np.float32)


def conv_weight(name: str, cout: int, cin_per_group: int, k: int) -> np.ndarray:
    """U(+-1/sqrt(fan_in)), fan_in = (Cin/groups)*k (the init the ref intends,
    Modules/SNAC/WNConv1d.cs:89-106)."""
    return _uniform(name, (cout, cin_per_group, k), 1.0 / np.sqrt(cin_per_group * k))


def convt_weight(name: str, cin: int, cout: int, k: int) -> np.ndarray:
    return _uniform(name, (cin, cout, k), 1.0 / np.sqrt(cin * k))


def bias(name: str, n: int, fan_in: int) -> np.ndarray:
    return _uniform(name, (n,), 1.0 / np.sqrt(fan_in))


def snake_alpha(name: str, c: int) -> np.ndarray:
    a = 1.0 + 0.1 * _rng(name).standard_normal(c)
    return np.clip(a, 0.5, 2.0).astype(np.float32).reshape(1, c, 1)


# --------------------------------------------------------------------------- DAC
class _Cfg:
    pass


def _norm_cfg(cfg):
    """Accept the oracle's DACConfig or the package's DACConfig (different field names)."""
    c = _Cfg()
    for f in ("sample_rate", "encoder_dim", "encoder_rates", "decoder_dim", "decoder_rates", "codebook_size",
              "codebook_dim"):
        setattr(c, f, getattr(cfg, f))
    c.n_codebooks = getattr(cfg, "n_codebooks", None) or getattr(cfg, "num_codebooks")
    c.latent_dim = getattr(cfg, "resolved_latent_dim", None) or getattr(cfg, "latent_dim")
    return c


def dac_hf_conv_shapes(cfg) -> Dict[str, tuple]:
    """Names and shapes of every tensor in the HF DacModel layout for `cfg`
    (everything except the codebooks, which are data dependent)."""
    cfg = _norm_cfg(cfg)
    shapes: Dict[str, tuple] = {}
    d = cfg.encoder_dim
    shapes["encoder.conv1"] = ("conv", d, 1, 7)
    for i, s in enumerate(cfg.encoder_rates):
        for u, _dil in enumerate((1, 3, 9), start=1):
            p = f"encoder.block.{i}.res_unit{u}"
            shapes[p + ".snake1"] = ("alpha", d)
            shapes[p + ".conv1"] = ("conv", d, d, 7)
            shapes[p + ".snake2"] = ("alpha", d)
            shapes[p + ".conv2"] = ("conv", d, d, 1)
        shapes[f"encoder.block.{i}.snake1"] = ("alpha", d)
        shapes[f"encoder.block.{i}.conv1"] = ("conv", 2 * d, d, 2 * s)
        d *= 2
    shapes["encoder.snake1"] = ("alpha", d)
    shapes["encoder.conv2"] = ("conv", cfg.latent_dim, d, 3)
    for q in range(cfg.n_codebooks):
        p = f"quantizer.quantizers.{q}"
        shapes[p + ".in_proj"] = ("conv", cfg.codebook_dim, cfg.latent_dim, 1)
        shapes[p + ".out_proj"] = ("conv", cfg.latent_dim, cfg.codebook_dim, 1)
    c = cfg.decoder_dim
    shapes["decoder.conv1"] = ("conv", c, cfg.latent_dim, 7)
    for i, s in enumerate(cfg.decoder_rates):
        cin, cout = c // (1 << i), c // (1 << (i + 1))
        shapes[f"decoder.block.{i}.snake1"] = ("alpha", cin)
        shapes[f"decoder.block.{i}.conv_t1"] = ("convt", cin, cout, 2 * s)
        for u in (1, 2, 3):
            p = f"decoder.block.{i}.res_unit{u}"
            shapes[p + ".snake1"] = ("alpha", cout)
            shapes[p + ".conv1"] = ("conv", cout, cout, 7)
            shapes[p + ".snake2"] = ("alpha", cout)
            shapes[p + ".conv2"] = ("conv", cout, cout, 1)
    cl = c // (1 << len(cfg.decoder_rates))
    shapes["decoder.snake1"] = ("alpha", cl)
    shapes["decoder.conv2"] = ("conv", 1, cl, 7)
    return shapes


def make_dac_weights_hf(cfg) -> Dict[str, np.ndarray]:
    """Seeded DAC weights in the HF safetensors layout (folded ``weight``); codebooks are
    N(0,1) rows (nn.Embedding default).  `cfg` needs encoder_dim, encoder_rates, decoder_dim,
    decoder_rates, n_codebooks (or num_codebooks), codebook_size, codebook_dim, latent_dim
    (or resolved_latent_dim).  oracle/synth.py replaces the codebooks by data-fitted ones."""
    cfg = _norm_cfg(cfg)
    sd: Dict[str, np.ndarray] = {}
    for name, spec in dac_hf_conv_shapes(cfg).items():
        kind = spec[0]
        if kind == "alpha":
            sd[name + ".alpha"] = snake_alpha(name + ".alpha", spec[1])
        elif kind == "conv":
            _, cout, cin, k = spec
            sd[name + ".weight"] = conv_weight(name + ".weight", cout, cin, k)
            sd[name + ".bias"] = bias(name + ".bias", cout, cin * k)
        elif kind == "convt":
            _, cin, cout, k = spec
            sd[name + ".weight"] = convt_weight(name + ".weight", cin, cout, k)
            sd[name + ".bias"] = bias(name + ".bias", cout, cin * k)
    K, D = cfg.codebook_size, cfg.codebook_dim
    for q in range(cfg.n_codebooks):
        nm = f"quantizer.quantizers.{q}.codebook.weight"
        sd[nm] = _rng(nm).standard_normal((K, D)).astype(np.float32)
    return sd


def save_safetensors(sd: Dict[str, np.ndarray], path: str, metadata: Optional[dict] = None) -> None:
    from safetensors.numpy import save_file
    save_file({k: np.ascontiguousarray(v) for k, v in sd.items()}, path, metadata=metadata)


def dia_codes(batch: int, frames: int, n_codebooks: int = 9, codebook_size: int = 1024,
              seed: int = 99) -> np.ndarray:
    """Config #5 input: codes[B, T, nq] int64 uniform in [0, K) (post-clamp domain of
    Models/Dia.cs:1039-1044)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, codebook_size, size=(batch, frames, n_codebooks), dtype=np.int64)


# --------------------------------------------------------------------------- SNAC
def _snac_cfg(cfg):
    c = _Cfg()
    for f in ("sample_rate", "encoder_dim", "encoder_rates", "decoder_dim", "decoder_rates", "attn_window_size",
              "codebook_size", "codebook_dim", "vq_strides", "noise", "depthwise"):
        setattr(c, f, getattr(cfg, f))
    c.latent_dim = getattr(cfg, "resolved_latent_dim", None) or getattr(cfg, "latent_dim")
    return c


def snac_layer_specs(cfg) -> Dict[str, tuple]:
    """Reference module-tree names (Modules/SNAC/*.cs registration order) -> layer kind and shape:
    ("conv", cout, cin_per_group, k, has_bias) | ("convt", cin, cout, k) | ("alpha", c) | ("mha", dim)."""
    c = _snac_cfg(cfg)
    specs: Dict[str, tuple] = {}

    def ru(p, dim, groups):
        specs[p + ".block.0"] = ("alpha", dim)
        specs[p + ".block.1"] = ("conv", dim, dim // groups, 7, True)
        specs[p + ".block.2"] = ("alpha", dim)
        specs[p + ".block.3"] = ("conv", dim, dim, 1, True)

    d = c.encoder_dim
    specs["encoder.block.0"] = ("conv", d, 1, 7, True)
    idx = 1
    for s in c.encoder_rates:
        d *= 2
        groups = d // 2 if c.depthwise else 1
        p = f"encoder.block.{idx}"
        for u in range(3):
            ru(f"{p}.block.{u}", d // 2, groups)
        specs[f"{p}.block.3"] = ("alpha", d // 2)
        specs[f"{p}.block.4"] = ("conv", d, d // 2, 2 * s, True)
        idx += 1
    if c.attn_window_size:
        specs[f"encoder.block.{idx}"] = ("mha", d)
        idx += 1
    specs[f"encoder.block.{idx}"] = ("conv", d, 1 if c.depthwise else d, 7, True)
    for q in range(len(c.vq_strides)):
        specs[f"quantizer.quantizers.{q}.in_proj"] = ("conv", c.codebook_dim, c.latent_dim, 1, True)
        specs[f"quantizer.quantizers.{q}.out_proj"] = ("conv", c.latent_dim, c.codebook_dim, 1, True)
    idx = 0
    if c.depthwise:
        specs["decoder.model.0"] = ("conv", c.latent_dim, 1, 7, True)
        specs["decoder.model.1"] = ("conv", c.decoder_dim, c.latent_dim, 1, True)
        idx = 2
    else:
        specs["decoder.model.0"] = ("conv", c.decoder_dim, c.latent_dim, 7, True)
        idx = 1
    if c.attn_window_size:
        specs[f"decoder.model.{idx}"] = ("mha", c.decoder_dim)
        idx += 1
    out_dim = 1
    for i, s in enumerate(c.decoder_rates):
        in_dim, out_dim = c.decoder_dim // (1 << i), c.decoder_dim // (1 << (i + 1))
        groups = out_dim if c.depthwise else 1
        p = f"decoder.model.{idx}"
        specs[f"{p}.block.0"] = ("alpha", in_dim)
        specs[f"{p}.block.1"] = ("convt", in_dim, out_dim, 2 * s)
        b = 2
        if c.noise:
            specs[f"{p}.block.2.linear"] = ("conv", out_dim, out_dim, 1, False)
            b = 3
        for u in range(3):
            ru(f"{p}.block.{b + u}", out_dim, groups)
        idx += 1
    specs[f"decoder.model.{idx}"] = ("alpha", out_dim)
    specs[f"decoder.model.{idx + 1}"] = ("conv", 1, out_dim, 7, True)
    return specs


def make_snac_weights(cfg) -> Dict[str, np.ndarray]:
    """Seeded SNAC weights in the reference's key layout (weight-norm stored as original0 = g = ||W||,
    original1 = v = W); codebooks N(0,1) (oracle/synth.py replaces them by data-fitted ones)."""
    c = _snac_cfg(cfg)
    sd: Dict[str, np.ndarray] = {}
    for name, spec in snac_layer_specs(cfg).items():
        kind = spec[0]
        if kind == "alpha":
            sd[name + ".alpha"] = snake_alpha(name + ".alpha", spec[1])
        elif kind in ("conv", "convt"):
            if kind == "conv":
                _, cout, cin_g, k, has_bias = spec
                w = conv_weight(name + ".weight", cout, cin_g, k)
                fan_in, nb = cin_g * k, cout
            else:
                _, cin, cout, k = spec
                w = convt_weight(name + ".weight", cin, cout, k)
                fan_in, nb, has_bias = cout * k, cout, True
            g = np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
            sd[name + ".parametrizations.weight.original0"] = g
            sd[name + ".parametrizations.weight.original1"] = w
            if has_bias:
                sd[name + ".bias"] = bias(name + ".bias", nb, fan_in)
        elif kind == "mha":
            dim = spec[1]
            sd[name + ".norm.weight"] = (1.0 + 0.1 * _rng(name + ".norm.weight").standard_normal(dim)).astype(np.float32)
            sd[name + ".norm.bias"] = (0.1 * _rng(name + ".norm.bias").standard_normal(dim)).astype(np.float32)
            sd[name + ".to_qkv.weight"] = _uniform(name + ".to_qkv.weight", (3 * dim, dim), 1.0 / np.sqrt(dim))
            sd[name + ".to_out.weight"] = _uniform(name + ".to_out.weight", (dim, dim), 1.0 / np.sqrt(dim))
            # SinusoidalEmbedding.cs:46: inv_freq = 1 / 10000^(arange(0, dim_head, 2) / dim_head), dim_head = 64
            sd[name + ".rel_pos.inv_freq"] = (1.0 / (10000.0 ** (np.arange(0, 64, 2, dtype=np.float32) / 64.0))).astype(np.float32)
    K, D = c.codebook_size, c.codebook_dim
    for q in range(len(c.vq_strides)):
        nm = f"quantizer.quantizers.{q}.codebook.weight"
        sd[nm] = _rng(nm).standard_normal((K, D)).astype(np.float32)
    return sd


def snac_noise(batch: int, lengths, first_clip: int = 0, seed: int = 777):
    """Explicit decoder noise tensors N(0,1), one [batch, 1, T_i] per decoder block."""
    out = []
    for i, t in enumerate(lengths):
        a = np.empty((batch, 1, t), np.float32)
        for b in range(batch):
            a[b, 0] = np.random.default_rng([seed, first_clip + b, i]).standard_normal(t).astype(np.float32)
        out.append(a)
    return out


# --------------------------------------------------------------------------- Encodec
def _enc_cfg(cfg):
    c = _Cfg()
    for f in ("sample_rate", "channels", "num_filters", "hidden_size", "upsampling_ratios", "num_residual_layers",
              "num_lstm_layers", "codebook_size"):
        setattr(c, f, getattr(cfg, f))
    c.num_quantizers = cfg.num_quantizers
    c.norm_type = getattr(cfg, "norm_type", "weight_norm")
    return c


def encodec_layer_specs(cfg) -> Dict[str, tuple]:
    """Reference Sequential indices (SEANetEncoder.cs:60-125, SEANetDecoder.cs:75-145) -> ("conv", cout, cin, k) |
    ("convt", cin, cout, k) | ("lstm", dim, layers)."""
    c = _enc_cfg(cfg)
    specs: Dict[str, tuple] = {}
    nf = c.num_filters

    def resnet(p, dim):
        specs[p + ".block.1"] = ("conv", dim // 2, dim, 3)
        specs[p + ".block.3"] = ("conv", dim, dim // 2, 1)
        specs[p + ".shortcut"] = ("conv", dim, dim, 1)

    mult, idx = 1, 1
    specs["encoder.layers.0"] = ("conv", nf, c.channels, 7)
    for r in reversed(c.upsampling_ratios):
        for _ in range(c.num_residual_layers):
            resnet(f"encoder.layers.{idx}", mult * nf)
            idx += 1
        idx += 1                                   # ELU
        specs[f"encoder.layers.{idx}"] = ("conv", mult * nf * 2, mult * nf, 2 * r)
        idx += 1
        mult *= 2
    if c.num_lstm_layers > 0:
        specs[f"encoder.layers.{idx}"] = ("lstm", mult * nf, c.num_lstm_layers)
        idx += 1
    idx += 1
    specs[f"encoder.layers.{idx}"] = ("conv", c.hidden_size, mult * nf, 7)

    mult = 2 ** len(c.upsampling_ratios)
    specs["decoder.layers.0"] = ("conv", mult * nf, c.hidden_size, 7)
    idx = 1
    if c.num_lstm_layers > 0:
        specs[f"decoder.layers.{idx}"] = ("lstm", mult * nf, c.num_lstm_layers)
        idx += 1
    for r in c.upsampling_ratios:
        idx += 1                                   # ELU
        specs[f"decoder.layers.{idx}"] = ("convt", mult * nf, mult * nf // 2, 2 * r)
        idx += 1
        for _ in range(c.num_residual_layers):
            resnet(f"decoder.layers.{idx}", mult * nf // 2)
            idx += 1
        mult //= 2
    idx += 1
    specs[f"decoder.layers.{idx}"] = ("conv", c.channels, nf, 7)
    return specs


def make_encodec_weights(cfg) -> Dict[str, np.ndarray]:
    """Seeded Encodec weights in the reference's key layout (weight_g = ||W||, weight_v = W; LSTM U(+-1/sqrt(H));
    codebooks N(0,1), inited = 1, cluster_size = 1, embed_avg = embed)."""
    c = _enc_cfg(cfg)
    sd: Dict[str, np.ndarray] = {}
    for name, spec in encodec_layer_specs(cfg).items():
        kind = spec[0]
        if kind == "lstm":
            _, dim, layers = spec
            bnd = 1.0 / np.sqrt(dim)
            for l in range(layers):
                sd[f"{name}.lstm.weight_ih_l{l}"] = _uniform(f"{name}.lstm.weight_ih_l{l}", (4 * dim, dim), bnd)
                sd[f"{name}.lstm.weight_hh_l{l}"] = _uniform(f"{name}.lstm.weight_hh_l{l}", (4 * dim, dim), bnd)
                sd[f"{name}.lstm.bias_ih_l{l}"] = _uniform(f"{name}.lstm.bias_ih_l{l}", (4 * dim,), bnd)
                sd[f"{name}.lstm.bias_hh_l{l}"] = _uniform(f"{name}.lstm.bias_hh_l{l}", (4 * dim,), bnd)
            continue
        if kind == "conv":
            _, cout, cin, k = spec
            w = conv_weight(name + ".conv.weight", cout, cin, k)
            fan_in, nb = cin * k, cout
        else:
            _, cin, cout, k = spec
            w = convt_weight(name + ".conv.weight", cin, cout, k)
            fan_in, nb = cout * k, cout
        if c.norm_type == "time_group_norm":     # plain conv + GroupNorm(1, C) affine (NormConv1d.cs:52-69)
            sd[name + ".conv.weight"] = w
            sd[name + ".norm.weight"] = (1.0 + 0.1 * _rng(name + ".norm.weight").standard_normal(nb)).astype(np.float32)
            sd[name + ".norm.bias"] = (0.1 * _rng(name + ".norm.bias").standard_normal(nb)).astype(np.float32)
        else:
            sd[name + ".conv.weight_g"] = np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True)).astype(np.float32)
            sd[name + ".conv.weight_v"] = w
        sd[name + ".conv.bias"] = bias(name + ".conv.bias", nb, fan_in)
    for q in range(c.num_quantizers):
        p = f"quantizer.layers.{q}.codebook"
        e = _rng(p + ".embed").standard_normal((c.codebook_size, c.hidden_size)).astype(np.float32)
        sd[p + ".embed"] = e
        sd[p + ".embed_avg"] = e.copy()
        sd[p + ".cluster_size"] = np.ones(c.codebook_size, np.float32)
        sd[p + ".inited"] = np.ones(1, np.float32)
    return sd
