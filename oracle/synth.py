"""Oracle-side synthetic data: the generic seeded generators live in
neuralcodecs_b200/synthetic.py; this adds DATA-FITTED codebooks, which need the CPU oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  SURVEY.md section 8(d): with nn.Embedding
default N(0,1) codebooks most RVQ stages emit a single code, so code parity would be
vacuous; stage-i codebook = K seeded-random rows of the oracle's stage-i projected residual.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from neuralcodecs_b200.synthetic import (AUDIO_SEED, WEIGHT_SEED, _rng, bias, conv_weight, convt_weight,  # noqa: F401
                                         dac_hf_conv_shapes, dia_codes, save_safetensors, snake_alpha,
                                         synth_audio)
from neuralcodecs_b200 import synthetic as _generic


def make_dac_weights_hf(cfg, codebook_clips: int = 2, codebook_seconds: float = 10.0,
                        codebooks: str = "data") -> Dict[str, np.ndarray]:
    """Seeded DAC weights in the HF safetensors layout (folded ``weight``).

    codebooks="data": fitted on the oracle's projected residuals (see module docstring).
    codebooks="normal": N(0,1) rows (nn.Embedding default), for structure tests."""
    sd = _generic.make_dac_weights_hf(cfg)
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




# --------------------------------------------------------------------------- SNAC
from neuralcodecs_b200.synthetic import snac_layer_specs, snac_noise  # noqa: E402,F401


def make_snac_weights(cfg, codebook_clips: int = 4, codebook_seconds: float = 4.0,
                      codebooks: str = "data") -> Dict[str, np.ndarray]:
    """Seeded SNAC weights in the reference's module-tree key layout; codebooks fitted on the oracle's
    projected residuals (rows re-used with jitter when fewer than K frames are available)."""
    sd = _generic.make_snac_weights(cfg)
    if codebooks == "data":
        _fit_snac_codebooks(cfg, sd, codebook_clips, codebook_seconds)
    return sd


def _fit_snac_codebooks(cfg, sd, clips: int, seconds: float) -> None:
    import torch
    from . import snac as snac_oracle

    K = cfg.codebook_size
    audio = torch.from_numpy(synth_audio(clips, int(round(seconds * cfg.sample_rate)), cfg.sample_rate)).unsqueeze(1)
    model = snac_oracle.SNACOracle(cfg, {k: torch.from_numpy(v) for k, v in sd.items()})
    with torch.inference_mode():
        z = model.encoder(model.preprocess(audio))
        residual = z.clone()
        for q in range(len(cfg.vq_strides)):
            ze = model.vq_in(q, residual)
            rows = ze.transpose(1, 2).reshape(-1, cfg.codebook_dim)
            nm = f"quantizer.quantizers.{q}.codebook.weight"
            rng = _rng(nm + "/pick")
            n = rows.shape[0]
            pick = np.sort(rng.choice(n, size=K, replace=n < K))
            cb = rows[torch.from_numpy(pick)].contiguous().clone()
            if n < K:   # re-used rows: jitter so entries stay distinct
                jit = rng.standard_normal(cb.shape).astype(np.float32) * 0.1 * float(rows.std())
                cb = cb + torch.from_numpy(jit)
            sd[nm] = cb.numpy().copy()
            model.sd[nm] = cb
            zqi, _, _ = model.vq_forward(q, residual)
            residual = residual - zqi


# --------------------------------------------------------------------------- Encodec
from neuralcodecs_b200.synthetic import encodec_layer_specs  # noqa: E402,F401


def make_encodec_weights(cfg, codebook_clips: int = 2, codebook_seconds: float = 10.0,
                         codebooks: str = "data", fit_stages: int = 8) -> Dict[str, np.ndarray]:
    """Seeded Encodec weights; the first `fit_stages` codebooks are fitted on the oracle's residuals."""
    sd = _generic.make_encodec_weights(cfg)
    if codebooks == "data":
        _fit_encodec_codebooks(cfg, sd, codebook_clips, codebook_seconds, fit_stages)
    return sd


def _fit_encodec_codebooks(cfg, sd, clips, seconds, stages) -> None:
    import torch
    from . import encodec as enc_oracle

    K = cfg.codebook_size
    audio = torch.from_numpy(synth_audio(clips * cfg.channels, int(round(seconds * cfg.sample_rate)), cfg.sample_rate))
    audio = audio.reshape(clips, cfg.channels, -1)
    model = enc_oracle.EncodecOracle(cfg, {k: torch.from_numpy(v) for k, v in sd.items()})
    with torch.inference_mode():
        # the latents the quantiser sees: per segment, after the per-segment loudness normalisation (Encodec.cs:259-285,469-480)
        length = audio.shape[-1]
        seg, stride = cfg.segment_length or length, cfg.segment_stride or length
        lat = []
        for off in range(0, length, stride):
            x = audio[:, :, off:min(off + seg, length)]
            if cfg.normalize:
                x = x / (x.mean([1], keepdim=True).pow(2).mean([2], keepdim=True).sqrt() + 1e-8)
            lat.append(model.encoder(x))
        residual = torch.cat(lat, dim=-1).clone()
        for q in range(min(stages, cfg.num_quantizers)):
            rows = residual.transpose(1, 2).reshape(-1, residual.shape[1])
            nm = f"quantizer.layers.{q}.codebook.embed"
            rng = _rng(nm + "/pick")
            n = rows.shape[0]
            pick = np.sort(rng.choice(n, size=K, replace=n < K))
            cb = rows[torch.from_numpy(pick)].contiguous().clone()
            if n < K:
                cb = cb + torch.from_numpy(rng.standard_normal(cb.shape).astype(np.float32) * 0.1 * float(rows.std()))
            sd[nm] = cb.numpy().copy()
            sd[f"quantizer.layers.{q}.codebook.embed_avg"] = sd[nm].copy()
            model.sd[nm] = cb
            quant, _ = model.vq_forward(q, residual)
            residual = residual - quant
