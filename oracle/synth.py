"""Seeded synthetic audio and weights shared by the oracle and the engine tests.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Implements SURVEY.md section 8(d):
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
def dac_hf_conv_shapes(cfg) -> Dict[str, tuple]:
    """Names and shapes of every tensor in the HF DacModel layout for `cfg`
    (everything except the codebooks, which are data dependent)."""
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


def make_dac_weights_hf(cfg, codebook_clips: int = 2, codebook_seconds: float = 10.0,
                        codebooks: str = "data") -> Dict[str, np.ndarray]:
    """Seeded DAC weights in the HF safetensors layout (folded ``weight``).

    codebooks="data": stage-i codebook = K seeded-random rows of the oracle's
    stage-i projected residual zE on the synthetic audio (default
    N(0,1) codebooks collapse to a single code, making code parity vacuous).
    codebooks="normal": N(0,1) rows (nn.Embedding default), for structure tests.
    """
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
    if codebooks == "data":
        _fit_dac_codebooks(cfg, sd, codebook_clips, codebook_seconds)
    return sd


def _fit_dac_codebooks(cfg, sd, clips: int, seconds: float) -> None:
    import torch
    from . import dac as dac_oracle

    length = int(round(seconds * cfg.sample_rate))
    K = cfg.codebook_size
    hop = int(np.prod(cfg.encoder_rates))
    # need at least K frames to sample K distinct rows
    while clips * ((length + hop - 1) // hop) < K:
        clips += 1
    audio = torch.from_numpy(synth_audio(clips, length, cfg.sample_rate)).unsqueeze(1)
    model = dac_oracle.DACOracle(cfg, dac_oracle.convert_hf_state_dict(
        {k: torch.from_numpy(v) for k, v in sd.items()}))
    with torch.inference_mode():
        z = model.encoder(model.preprocess(audio))
        residual = z.clone()
        for q in range(cfg.n_codebooks):
            ze = model.vq_in_proj(q, residual)                       # [B, D, T]
            rows = ze.transpose(1, 2).reshape(-1, cfg.codebook_dim)
            nm = f"quantizer.quantizers.{q}.codebook.weight"
            pick = _rng(nm + "/pick").choice(rows.shape[0], size=K, replace=False)
            cb = rows[torch.from_numpy(np.sort(pick))].contiguous().clone()
            sd[nm] = cb.numpy().copy()
            model.set_codebook(q, cb)
            zq_i, _idx, _ze = model.vq_forward(q, residual)
            residual = residual - zq_i


def save_safetensors(sd: Dict[str, np.ndarray], path: str, metadata: Optional[dict] = None) -> None:
    from safetensors.numpy import save_file
    save_file({k: np.ascontiguousarray(v) for k, v in sd.items()}, path, metadata=metadata)


def dia_codes(batch: int, frames: int, n_codebooks: int = 9, codebook_size: int = 1024,
              seed: int = 99) -> np.ndarray:
    """Config #5 input: codes[B, T, nq] int64 uniform in [0, K) (post-clamp domain of
    Models/Dia.cs:1039-1044)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, codebook_size, size=(batch, frames, n_codebooks), dtype=np.int64)
