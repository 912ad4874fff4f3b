"""CPU oracle for the reference's SNAC model.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED by the reference (it has no tests); see oracle/__init__.py.

Op-for-op restatement, in the reference's order, of
  Models/SNAC.cs                         (Preprocess / forward / Encode / Decode / ProcessAudio)
  Modules/SNAC/{Encoder,EncoderBlock,ResidualUnit,Snake1d,WNConv1d,WNConvTranspose1d,Decoder,
                DecoderBlock,NoiseBlock,VectorQuantizer,ResidualVectorQuantizer,LocalMHA,
                SinusoidalEmbedding,RotaryEmbedding}.cs
  Config/SNAC/SNACConfig.cs
(paths relative to /root/reference/NeuralCodecs.Torch/).  Weight keys are the reference's
module-tree names: e.g. ``encoder.block.1.block.0.block.1.parametrizations.weight.original1``.

Deviations of the reference that are reproduced here: un-normalised VQ distance
(VectorQuantizer.cs:125-137), weight-norm ``w = (v/||v||) * (g - 1e-7)`` (WNConv1d.cs:132-135),
Snake without epsilon (Snake1d.cs:57), ``Encode``/``forward`` use the padded audio (SNAC.cs:96-98,
142-143), decoder noise is an explicit input here (NoiseBlock.cs:38-45 draws randn).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class SNACConfig:
    """Mirror of Config/SNAC/SNACConfig.cs:11-153."""
    sample_rate: int = 44100
    encoder_dim: int = 64
    encoder_rates: List[int] = field(default_factory=lambda: [2, 3, 8, 8])
    latent_dim_opt: Optional[int] = None
    decoder_dim: int = 1536
    decoder_rates: List[int] = field(default_factory=lambda: [8, 8, 3, 2])
    attn_window_size: Optional[int] = 32
    codebook_size: int = 4096
    codebook_dim: int = 8
    vq_strides: List[int] = field(default_factory=lambda: [8, 4, 2, 1])
    noise: bool = True
    depthwise: bool = True

    @property
    def latent_dim(self) -> int:      # Models/SNAC.cs:37
        return self.latent_dim_opt if self.latent_dim_opt is not None else self.encoder_dim * (1 << len(self.encoder_rates))

    @property
    def hop_length(self) -> int:      # Models/SNAC.cs:38
        return int(math.prod(self.encoder_rates))

    @property
    def pad_multiple(self) -> int:    # Models/SNAC.cs:74-75
        return self.hop_length * math.lcm(self.vq_strides[0], self.attn_window_size or 1)

    @staticmethod
    def snac_24khz() -> "SNACConfig":  # SNACConfig.cs:139-153
        return SNACConfig(sample_rate=24000, encoder_dim=48, encoder_rates=[2, 4, 8, 8], decoder_dim=1024,
                          decoder_rates=[8, 8, 4, 2], attn_window_size=None, vq_strides=[4, 2, 1])

    @staticmethod
    def snac_32khz() -> "SNACConfig":  # SNACConfig.cs:119-133
        return SNACConfig(sample_rate=32000)

    @staticmethod
    def snac_44khz() -> "SNACConfig":  # SNACConfig.cs:113
        return SNACConfig()


class SNACOracle:
    def __init__(self, cfg: SNACConfig, sd: Dict[str, torch.Tensor], dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}

    # ---------------------------------------------------------------- layers
    def _weight(self, name: str):
        """WNConv1d.forward weight (SNAC/WNConv1d.cs:122-135): (v / ||v||_(1,2)) * (g - 1e-7)."""
        v = self.sd[name + ".parametrizations.weight.original1"]
        g = self.sd[name + ".parametrizations.weight.original0"].reshape(-1, 1, 1)
        v_norm = v.contiguous().pow(2).sum([1, 2], keepdim=True, dtype=self.dtype).sqrt()
        return torch.mul(v.div(v_norm), g.sub(1e-7)).contiguous()

    def wnconv1d(self, name, x, stride=1, padding=0, dilation=1, groups=1):
        return F.conv1d(x, self._weight(name), self.sd.get(name + ".bias"), stride, padding, dilation, groups)

    def wnconvtranspose1d(self, name, x, stride, padding, output_padding):
        """SNAC/WNConvTranspose1d.cs:126-143."""
        return F.conv_transpose1d(x, self._weight(name), self.sd.get(name + ".bias"), stride=stride, padding=padding,
                                  output_padding=output_padding, groups=1, dilation=1)

    def snake(self, name, x):
        """SNAC/Snake1d.cs:55-58."""
        alpha = self.sd[name + ".alpha"]
        return torch.where(alpha == 0, x, torch.addcdiv(x, torch.sin(alpha * x).pow_(2), alpha, value=1))

    def residual_unit(self, p, x, dilation, groups):
        """SNAC/ResidualUnit.cs:25-60."""
        pad = (7 - 1) * dilation // 2
        y = self.snake(p + ".block.0", x)
        y = self.wnconv1d(p + ".block.1", y, padding=pad, dilation=dilation, groups=groups)
        y = self.snake(p + ".block.2", y)
        y = self.wnconv1d(p + ".block.3", y)
        d = (x.shape[-1] - y.shape[-1]) // 2
        if d > 0:
            x = x[..., d:-d]
        return x.add(y)

    def local_mha(self, p, x, rotary: bool):
        """SNAC/LocalMHA.cs:80-115 (+ SinusoidalEmbedding.cs:60-106, RotaryEmbedding.cs:16-68)."""
        w = self.cfg.attn_window_size
        B, C, T = x.shape
        H = C // 64
        residual = x
        h = F.layer_norm(x.transpose(1, 2), (C,), self.sd[p + ".norm.weight"], self.sd[p + ".norm.bias"], 1e-5)
        windows = T // w
        q, k, v = F.linear(h, self.sd[p + ".to_qkv.weight"]).chunk(3, dim=-1)

        def rearr(t):
            return t.reshape(B, windows, T // windows, H, C // H).permute(0, 3, 1, 2, 4)

        q, k, v = rearr(q), rearr(k), rearr(v)
        if rotary:
            inv_freq = self.sd[p + ".rel_pos.inv_freq"]
            t = torch.arange(k.shape[-2], dtype=inv_freq.dtype)
            freqs = torch.einsum("i,j->ij", t, inv_freq)
            freqs = torch.cat((freqs, freqs), dim=-1)          # SinusoidalEmbedding.cs:88-92; scale = 1 (use_xpos off)

            def rot_half(t_):
                a, b = t_.chunk(2, dim=-1)
                return torch.cat((-b, a), dim=-1)

            q = q * freqs.cos() + rot_half(q) * freqs.sin()
            k = k * freqs.cos() + rot_half(k) * freqs.sin()
        a = F.scaled_dot_product_attention(q, k, v)
        out = a.permute(0, 2, 3, 1, 4).reshape(B, T, C)
        out = F.linear(out, self.sd[p + ".to_out.weight"])
        return out.transpose(1, 2).add_(residual)

    # ---------------------------------------------------------------- encoder / decoder
    def encoder(self, x):
        """SNAC/Encoder.cs:26-69, EncoderBlock.cs:27-55."""
        c = self.cfg
        x = self.wnconv1d("encoder.block.0", x, padding=3)
        d = c.encoder_dim
        idx = 1
        for s in c.encoder_rates:
            d *= 2
            groups = d // 2 if c.depthwise else 1
            p = f"encoder.block.{idx}"
            for u, dil in enumerate((1, 3, 9)):
                x = self.residual_unit(f"{p}.block.{u}", x, dil, groups)
            x = self.snake(f"{p}.block.3", x)
            x = self.wnconv1d(f"{p}.block.4", x, stride=s, padding=math.ceil(s / 2.0))
            idx += 1
        if c.attn_window_size:
            x = self.local_mha(f"encoder.block.{idx}", x, rotary=True)
            idx += 1
        return self.wnconv1d(f"encoder.block.{idx}", x, padding=3, groups=d if c.depthwise else 1)

    def decoder(self, x, noise: Optional[Sequence[torch.Tensor]] = None):
        """SNAC/Decoder.cs:28-86, DecoderBlock.cs:23-70, NoiseBlock.cs:23-45.  noise[i]: [B,1,T_i] for block i."""
        c = self.cfg
        idx = 0
        if c.depthwise:
            x = self.wnconv1d("decoder.model.0", x, padding=3, groups=c.latent_dim)
            x = self.wnconv1d("decoder.model.1", x)
            idx = 2
        else:
            x = self.wnconv1d("decoder.model.0", x, padding=3)
            idx = 1
        if c.attn_window_size:
            x = self.local_mha(f"decoder.model.{idx}", x, rotary=True)   # Decoder.cs:56: default useRotaryPosEmb
            idx += 1
        out_dim = 1
        for i, s in enumerate(c.decoder_rates):
            out_dim = c.decoder_dim // (1 << (i + 1))
            groups = out_dim if c.depthwise else 1
            p = f"decoder.model.{idx}"
            x = self.snake(f"{p}.block.0", x)
            x = self.wnconvtranspose1d(f"{p}.block.1", x, stride=s, padding=math.ceil(s / 2.0), output_padding=s % 2)
            b = 2
            if c.noise:
                h = self.wnconv1d(f"{p}.block.2.linear", x)
                n = noise[i].to(self.dtype) if noise is not None else torch.zeros(x.shape[0], 1, x.shape[2], dtype=self.dtype)
                x = x + n * h
                b = 3
            for u, dil in enumerate((1, 3, 9)):
                x = self.residual_unit(f"{p}.block.{b + u}", x, dil, groups)
            idx += 1
        x = self.snake(f"decoder.model.{idx}", x)
        x = self.wnconv1d(f"decoder.model.{idx + 1}", x, padding=3)
        return torch.tanh(x)

    def noise_lengths(self, frames: int) -> List[int]:
        """Length of the noise tensor of each decoder block for a latent of `frames` steps."""
        out, t = [], frames
        for s in self.cfg.decoder_rates:
            t = (t - 1) * s - 2 * math.ceil(s / 2.0) + 2 * s + (s % 2)
            out.append(t)
        return out

    # ---------------------------------------------------------------- quantizer
    def vq_distances(self, q, ze):
        D = self.cfg.codebook_dim
        enc = ze.transpose(1, 2).reshape(-1, D).to(self.dtype).contiguous()
        cb = self.sd[f"quantizer.quantizers.{q}.codebook.weight"].to(self.dtype).contiguous()
        e2 = enc.pow(2).sum(1, keepdim=True)
        c2 = cb.pow(2).sum(1, keepdim=True)
        cross = torch.einsum("bd,nd->bn", enc, cb).mul_(2.0)
        return e2 + c2.t() - cross

    def vq_decode_code(self, q, idx):
        cb = self.sd[f"quantizer.quantizers.{q}.codebook.weight"]
        return F.embedding(idx, cb).contiguous().transpose(1, 2).contiguous()

    def vq_in(self, q, z):
        """avg_pool (stride) + in_proj (VectorQuantizer.cs:86-91)."""
        s = self.cfg.vq_strides[q]
        if s > 1:
            z = F.avg_pool1d(z, kernel_size=s, stride=s)
        return self.wnconv1d(f"quantizer.quantizers.{q}.in_proj", z)

    def vq_forward(self, q, z):
        """VectorQuantizer.forward (SNAC/VectorQuantizer.cs:82-103) -> (zQ, indices, zE)."""
        s = self.cfg.vq_strides[q]
        ze = self.vq_in(q, z)
        dist = self.vq_distances(q, ze)
        idx = dist.argmin(1).reshape(ze.shape[0], ze.shape[2])
        zq = self.vq_decode_code(q, idx)
        zq = ze + (zq - ze)
        zq = self.wnconv1d(f"quantizer.quantizers.{q}.out_proj", zq)
        if s > 1:
            zq = zq.repeat_interleave(s, dim=-1)
        return zq, idx, ze

    def rvq_forward(self, z):
        """ResidualVectorQuantizer.forward (SNAC/ResidualVectorQuantizer.cs:69-89)."""
        z = z.contiguous()
        zq = torch.zeros_like(z)
        residual = z.clone()
        codes = []
        for i in range(len(self.cfg.vq_strides)):
            zqi, idx, _ = self.vq_forward(i, residual)
            zq = torch.add(zq, zqi)
            residual = torch.sub(residual, zqi)
            codes.append(idx.clone())
        return zq, codes

    def rvq_from_codes(self, codes: Sequence[torch.Tensor]):
        """ResidualVectorQuantizer.FromCodes (SNAC/ResidualVectorQuantizer.cs:91-131)."""
        if len(codes) != len(self.cfg.vq_strides):
            raise ValueError(f"Expected {len(self.cfg.vq_strides)} codebooks but got {len(codes)}")
        zq = None
        for i, c in enumerate(codes):
            zpi = self.vq_decode_code(i, c)
            zqi = self.wnconv1d(f"quantizer.quantizers.{i}.out_proj", zpi)
            s = self.cfg.vq_strides[i]
            if s > 1:
                zqi = zqi.repeat_interleave(s, dim=-1)
            zq = zqi if zq is None else torch.add(zq, zqi)
        return zq

    # ---------------------------------------------------------------- model surface
    def preprocess(self, audio):
        """SNAC.Preprocess (Models/SNAC.cs:70-80)."""
        length = audio.shape[-1]
        pad_to = self.cfg.pad_multiple
        right = int(math.ceil(length / pad_to) * pad_to) - length
        return F.pad(audio.to(self.dtype), [0, right])

    def encode(self, audio):
        """SNAC.Encode(float[]) / forward's encode half (SNAC.cs:129-150): codes of the PADDED audio."""
        with torch.inference_mode():
            z = self.encoder(self.preprocess(audio))
            _, codes = self.rvq_forward(z)
            return codes

    def encode_latent(self, audio):
        with torch.inference_mode():
            return self.encoder(self.preprocess(audio))

    def decode(self, codes, noise=None):
        """SNAC.Decode (SNAC.cs:157-192): not trimmed."""
        with torch.inference_mode():
            return self.decoder(self.rvq_from_codes(codes), noise)

    def forward(self, audio, noise=None):
        """SNAC.forward (SNAC.cs:91-106): trimmed to the input length."""
        with torch.inference_mode():
            length = audio.shape[-1]
            z = self.encoder(self.preprocess(audio))
            zq, codes = self.rvq_forward(z)
            a = self.decoder(zq, noise)
            return {"audio": a[..., :length], "codes": codes, "z": z, "zq": zq}


def load_safetensors(path: str, cfg: SNACConfig, dtype=torch.float32) -> SNACOracle:
    from safetensors.torch import load_file
    return SNACOracle(cfg, load_file(path), dtype)


# ------------------------------------------------------------------------------------------------------------------
# Input conditioning in front of the model (test infrastructure, numpy float64 = the reference's C# double arithmetic)
def resample_linear(x, src: int, dst: int):
    """SNAC.ResampleAudio (Models/SNAC.cs:284-308) = AudioUtils.ResampleLinear (NeuralCodecs.Core/Utils/AudioUtils.cs:329-352):
    position = i / ratio, index = (int)position, two-point interpolation in double, last sample held."""
    import numpy as np
    x = np.asarray(x, np.float32)
    ratio = float(dst) / float(src)
    n = int(len(x) * ratio)
    pos = np.arange(n, dtype=np.float64) / ratio
    idx = pos.astype(np.int64)
    frac = pos - idx
    last = idx >= len(x) - 1
    i0 = np.minimum(idx, len(x) - 1)
    i1 = np.minimum(idx + 1, len(x) - 1)
    out = (1 - frac) * x[i0].astype(np.float64) + frac * x[i1].astype(np.float64)
    out[last] = x[-1]
    return out.astype(np.float32)


def resample_linear_loop(x, src: int, dst: int):
    """The same, as the reference's scalar loop (pure Python; small inputs only) -- pins the vectorised form."""
    import numpy as np
    ratio = float(dst) / float(src)
    n = int(len(x) * ratio)
    out = np.zeros(n, np.float32)
    for i in range(n):
        position = i / ratio
        index = int(position)
        fraction = position - index
        if index >= len(x) - 1:
            out[i] = x[len(x) - 1]
        else:
            out[i] = np.float32((1 - fraction) * float(x[index]) + fraction * float(x[index + 1]))
    return out


def convert_to_mono(x, channels: int):
    """AudioUtils.ConvertToMono (AudioUtils.cs:45-62): float32 running sum in channel order, divided by the channel count."""
    import numpy as np
    x = np.asarray(x, np.float32)
    frames = len(x) // channels
    s = np.zeros(frames, np.float32)
    for ch in range(channels):
        s = (s + x[ch:frames * channels:channels]).astype(np.float32)
    return (s / np.float32(channels)).astype(np.float32)
