"""CPU oracle for the reference's DAC model.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED by the reference (it has no tests); see oracle/__init__.py.

Op-for-op restatement, in the reference's order, of
  Models/DAC.cs                                (Preprocess/Encode/Decode/FromCodes/forward)
  Modules/DAC/{Encoder,EncoderBlock,ResidualUnit,Snake1d,WNConv1d,WNConvTranspose1d,
               Decoder,DecoderBlock,VectorQuantizer,ResidualVectorQuantizer}.cs
  Config/DAC/{DACConfig,StateDictNameConverter}.cs
using the same ATen ops TorchSharp dispatches to (torch CPU, fp32; fp64 on request
to measure near-tie margins).  Paths are relative to
/root/reference/NeuralCodecs.Torch/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class DACConfig:
    """Mirror of Config/DAC/DACConfig.cs:8-100 (fields the hot path reads)."""
    sample_rate: int = 44100
    encoder_dim: int = 64
    encoder_rates: List[int] = field(default_factory=lambda: [2, 4, 8, 8])
    decoder_dim: int = 1536
    decoder_rates: List[int] = field(default_factory=lambda: [8, 8, 4, 2])
    n_codebooks: int = 9
    codebook_size: int = 1024
    codebook_dim: int = 8
    latent_dim_opt: Optional[int] = None

    @property
    def latent_dim(self) -> int:
        # Models/DAC.cs:64 : config.LatentDim ?? encoderDim * 2^len(rates)
        if self.latent_dim_opt is not None:
            return self.latent_dim_opt
        return self.encoder_dim * (1 << len(self.encoder_rates))

    @property
    def hop_length(self) -> int:
        # Models/DAC.cs:67
        return int(math.prod(self.encoder_rates))

    @staticmethod
    def dac_44khz() -> "DACConfig":  # DACConfig.cs:103
        return DACConfig()

    @staticmethod
    def dac_24khz() -> "DACConfig":  # DACConfig.cs:116-124
        return DACConfig(sample_rate=24000, n_codebooks=32, encoder_rates=[2, 4, 5, 8],
                         decoder_rates=[8, 5, 4, 2])


def convert_hf_state_dict(hf: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """HF ``DacModel`` safetensors keys -> weight_v / weight_g / bias / alpha per layer.

    Follows StateDictNameConverter.ConvertFromSafetensor (StateDictNameConverter.cs:40-65):
    every conv ``.weight`` becomes ``weight_v`` and ``weight_g := sqrt(sum(w^2, dims (1,2)))``
    in fp32 (:48-58, TranslateKey :342-376).  The module-tree renaming (BuildKeyMap :274-340)
    does not affect arithmetic, so the HF prefixes are kept as layer names here.
    """
    out: Dict[str, torch.Tensor] = {}
    for k, v in hf.items():
        if k.endswith(".weight") and ("conv" in k.rsplit(".", 2)[-2] or "_proj" in k):
            w = v.to(torch.float32)
            norm = w.contiguous().pow(2).sum([1, 2], keepdim=True, dtype=torch.float32).sqrt()
            out[k[:-7] + ".weight_v"] = w
            out[k[:-7] + ".weight_g"] = norm
        else:
            out[k] = v
    return out


class DACOracle:
    def __init__(self, cfg: DACConfig, sd: Dict[str, torch.Tensor], dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}

    # ---------------------------------------------------------------- layers
    def wnconv1d(self, name: str, x, stride=1, padding=0, dilation=1, groups=1):
        """Modules/DAC/WNConv1d.cs:140-156 (weight re-normalised every forward)."""
        v, g = self.sd[name + ".weight_v"], self.sd[name + ".weight_g"]
        b = self.sd.get(name + ".bias")
        v_norm = v.contiguous().pow(2).sum([1, 2], keepdim=True, dtype=self.dtype).sqrt()
        normalized = v.div(v_norm.add(1e-7))
        w = torch.mul(normalized, g).contiguous()
        return F.conv1d(x, w, b, stride, padding, dilation, groups)

    def wnconvtranspose1d(self, name: str, x, stride=1, padding=0, output_padding=0):
        """Modules/DAC/WNConvTranspose1d.cs:141-162; norm over dims (1,2) of
        v[Cin,Cout,k] => per-in-channel gain."""
        v, g = self.sd[name + ".weight_v"], self.sd[name + ".weight_g"]
        b = self.sd.get(name + ".bias")
        v_norm = v.contiguous().pow(2).sum([1, 2], keepdim=True, dtype=self.dtype).sqrt()
        normalized = v.div(v_norm.add(1e-7))
        w = torch.mul(normalized, g).contiguous()
        return F.conv_transpose1d(x, w, b, stride=stride, padding=padding,
                                  output_padding=output_padding, groups=1, dilation=1)

    def snake(self, name: str, x):
        """Modules/DAC/Snake1d.cs:49-58: where(a==0, x, addcdiv(x, sin(a*x)^2, a)); no epsilon."""
        alpha = self.sd[name + ".alpha"]
        return torch.where(alpha == 0, x, torch.addcdiv(x, torch.sin(alpha * x).pow_(2), alpha, value=1))

    def residual_unit(self, prefix: str, x, dilation: int):
        """Modules/DAC/ResidualUnit.cs:24-59."""
        pad = (7 - 1) * dilation // 2
        y = self.snake(prefix + ".snake1", x)
        y = self.wnconv1d(prefix + ".conv1", y, padding=pad, dilation=dilation)
        y = self.snake(prefix + ".snake2", y)
        y = self.wnconv1d(prefix + ".conv2", y)
        p = (x.shape[-1] - y.shape[-1]) // 2
        if p > 0:
            x = x[..., p:-p]
        return y.add_(x)

    # ---------------------------------------------------------------- encoder / decoder
    def encoder(self, x):
        """Modules/DAC/Encoder.cs:21-58 + EncoderBlock.cs:20-43."""
        x = self.wnconv1d("encoder.conv1", x, padding=3)
        for i, s in enumerate(self.cfg.encoder_rates):
            for u, dil in enumerate((1, 3, 9), start=1):
                x = self.residual_unit(f"encoder.block.{i}.res_unit{u}", x, dil)
            x = self.snake(f"encoder.block.{i}.snake1", x)
            x = self.wnconv1d(f"encoder.block.{i}.conv1", x, stride=s, padding=math.ceil(s / 2.0))
        x = self.snake("encoder.snake1", x)
        return self.wnconv1d("encoder.conv2", x, padding=1)

    def decoder(self, x):
        """Modules/DAC/Decoder.cs:22-58 + DecoderBlock.cs:20-44."""
        x = self.wnconv1d("decoder.conv1", x, padding=3)
        for i, s in enumerate(self.cfg.decoder_rates):
            x = self.snake(f"decoder.block.{i}.snake1", x)
            x = self.wnconvtranspose1d(f"decoder.block.{i}.conv_t1", x, stride=s,
                                       padding=math.ceil(s / 2.0))
            for u, dil in enumerate((1, 3, 9), start=1):
                x = self.residual_unit(f"decoder.block.{i}.res_unit{u}", x, dil)
        x = self.snake("decoder.snake1", x)
        x = self.wnconv1d("decoder.conv2", x, padding=3)
        return torch.tanh(x)

    # ---------------------------------------------------------------- quantizer
    def set_codebook(self, q: int, cb: torch.Tensor) -> None:
        self.sd[f"quantizer.quantizers.{q}.codebook.weight"] = cb.to(self.dtype)

    def vq_in_proj(self, q: int, z):
        return self.wnconv1d(f"quantizer.quantizers.{q}.in_proj", z)

    def vq_out_proj(self, q: int, z):
        return self.wnconv1d(f"quantizer.quantizers.{q}.out_proj", z)

    def vq_distances(self, q: int, ze):
        """dist[B*T, K] exactly as VectorQuantizer.DecodeLatents (VectorQuantizer.cs:99-118):
        un-normalised, expanded form ||e||^2 + ||c||^2^T - 2 e.c^T."""
        D = self.cfg.codebook_dim
        enc = ze.transpose(1, 2).reshape(-1, D).to(self.dtype).contiguous()
        cb = self.sd[f"quantizer.quantizers.{q}.codebook.weight"].to(self.dtype).contiguous()
        e2 = enc.pow(2).sum(1, keepdim=True)
        c2 = cb.pow(2).sum(1, keepdim=True)
        cross = torch.einsum("bd,nd->bn", enc, cb).mul_(2.0)
        return e2 + c2.t() - cross

    def vq_decode_code(self, q: int, idx):
        """VectorQuantizer.DecodeCode (VectorQuantizer.cs:135-142)."""
        cb = self.sd[f"quantizer.quantizers.{q}.codebook.weight"]
        return F.embedding(idx, cb).contiguous().transpose(-2, -1).contiguous()

    def vq_forward(self, q: int, z):
        """VectorQuantizer.forward (VectorQuantizer.cs:64-91) -> (zQ, indices, zE)."""
        ze = self.vq_in_proj(q, z).to(self.dtype)
        dist = self.vq_distances(q, ze)
        idx = dist.argmin(1).reshape(ze.shape[0], ze.shape[-1]).to(torch.int64)
        zq = self.vq_decode_code(q, idx)
        zq = ze + (zq - ze)                      # straight-through arithmetic kept at inference (:81)
        zq = self.vq_out_proj(q, zq)
        return zq, idx, ze

    def rvq_forward(self, z, n_quantizers: Optional[int] = None):
        """ResidualVectorQuantizer.forward, both overloads
        (ResidualVectorQuantizer.cs:54-103 and :105-206, eval mode: first nQ stages,
        mask all-true)."""
        residual = z.clone()
        zq = torch.zeros_like(z)
        codes, latents = [], []
        nq = self.cfg.n_codebooks if n_quantizers is None else min(n_quantizers, self.cfg.n_codebooks)
        for i in range(nq):
            zqi, idx, ze = self.vq_forward(i, residual)
            zq.add_(zqi)
            residual.sub_(zqi)
            codes.append(idx)
            latents.append(ze)
        return zq, torch.stack(codes, 1), torch.cat(latents, 1)

    def rvq_from_codes(self, codes):
        """ResidualVectorQuantizer.FromCodes (ResidualVectorQuantizer.cs:211-238):
        no straight-through arithmetic; accumulator starts as an int64 zeros(1)."""
        zq = torch.zeros(1, dtype=codes.dtype)
        for i in range(codes.shape[1]):
            zpi = self.vq_decode_code(i, codes[:, i, :])
            zqi = self.vq_out_proj(i, zpi)
            zq = zq.add(zqi)
        return zq

    # ---------------------------------------------------------------- model surface
    def preprocess(self, audio, sample_rate: Optional[int] = None):
        """DAC.Preprocess (Models/DAC.cs:141-154)."""
        sr = self.cfg.sample_rate if sample_rate is None else sample_rate
        if sr != self.cfg.sample_rate:
            raise ValueError(f"Input audio sample rate {sr}Hz does not match model sample rate "
                             f"{self.cfg.sample_rate}Hz")
        length = audio.shape[-1]
        hop = self.cfg.hop_length
        right = int(math.ceil(length / hop) * hop) - length
        return F.pad(audio.to(self.dtype), [0, right])

    def encode(self, audio, n_quantizers: Optional[int] = None, sample_rate: Optional[int] = None):
        """DAC.Encode(Tensor, nQ?, sr?) (Models/DAC.cs:163-181) -> (z, codes, latents)."""
        with torch.inference_mode():
            x = self.preprocess(audio, sample_rate)
            z = self.encoder(x)
            return self.rvq_forward(z, n_quantizers)

    def encode_latent(self, audio):
        """Encoder output before quantisation (not a ref API; used for teacher-forced checks)."""
        with torch.inference_mode():
            return self.encoder(self.preprocess(audio))

    def decode(self, z):
        """DAC.Decode (Models/DAC.cs:231-234); output is NOT trimmed to the input length."""
        with torch.inference_mode():
            return self.decoder(z.to(self.dtype))

    def from_codes(self, codes):
        """DAC.FromCodes (Models/DAC.cs:101-106)."""
        with torch.inference_mode():
            return self.rvq_from_codes(codes)

    def forward(self, audio, n_quantizers: Optional[int] = None):
        """DAC.forward (Models/DAC.cs:262-322)."""
        z, codes, latents = self.encode(audio, n_quantizers)
        return {"audio": self.decode(z), "z": z, "codes": codes, "latents": latents}

    def dia_decode(self, codes_tc):
        """Dia.Decode (Models/Dia.cs:973-981): codes[T,nq] -> FromCodes([1,nq,T]) -> Decode -> squeeze."""
        z = self.from_codes(codes_tc.unsqueeze(0).transpose(1, 2))
        return self.decode(z).squeeze()


def dia_generate_output(model: "DACOracle", generated: torch.Tensor, lengths, delay_pattern=(0, 8, 9, 10, 11, 12, 13, 14, 15),
                        min_valid: int = 0, max_valid: int = 1023):
    """Dia.GenerateOutput's codec stage (Models/Dia.cs:1010-1060 with Modules/Dia/AudioUtils.cs:108-176):
    BuildRevertIndices (t + delay clamped to T-1) -> gather -> where(t_idx >= T, pad, x) (never true after the clamp)
    -> drop the last max(delay) steps -> invalid codes := 0 -> per item [:length] -> Dia.Decode (serial loop)."""
    B, T, Cn = generated.shape
    delay = torch.tensor(delay_pattern, dtype=torch.int64)
    t_idx = torch.minimum(torch.arange(T).view(1, T, 1) + delay.view(1, 1, Cn), torch.tensor(T - 1)).expand(B, T, Cn)
    b_idx = torch.arange(B).view(B, 1, 1).expand(B, T, Cn)
    c_idx = torch.arange(Cn).view(1, 1, Cn).expand(B, T, Cn)
    gathered = generated[b_idx.reshape(-1), t_idx.reshape(-1), c_idx.reshape(-1)].view(B, T, Cn)
    reverted = torch.where(t_idx >= T, torch.tensor(1025, dtype=generated.dtype), gathered)
    codebook = reverted[:, : T - max(delay_pattern), :].clone()
    codebook[(codebook < min_valid) | (codebook > max_valid)] = 0
    return [model.dia_decode(codebook[i, : int(lengths[i]), :]) for i in range(B)]


def load_hf_safetensors(path: str, cfg: DACConfig, dtype=torch.float32) -> DACOracle:
    from safetensors.torch import load_file
    return DACOracle(cfg, convert_hf_state_dict(load_file(path)), dtype)


# ------------------------------------------------------------------ parity accounting
def near_tie_report(model: DACOracle, z, codes_ref, codes_test) -> Dict[str, object]:
    """Classify code mismatches per SURVEY 8(d) 'Parity report'.

    For each frame the first stage that differs is examined teacher-forced from the
    oracle's own residual: margin of the TEST code vs the oracle's best code under both
    normalisations.  Later stages of an already-flipped frame are cascades and excluded.
    """
    with torch.inference_mode():
        B, nq, T = codes_ref.shape
        residual = z.clone().to(model.dtype)
        flipped = torch.zeros(B, T, dtype=torch.bool)
        rows = []
        match_per_stage = []
        for i in range(nq):
            ze = model.vq_in_proj(i, residual)
            dist = model.vq_distances(i, ze).reshape(B, T, -1)
            ref_i, test_i = codes_ref[:, i, :], codes_test[:, i, :]
            match_per_stage.append(float((ref_i == test_i).float().mean()))
            new = (ref_i != test_i) & ~flipped
            if new.any():
                e = ze.transpose(1, 2)
                cb = model.sd[f"quantizer.quantizers.{i}.codebook.weight"]
                for b, t in new.nonzero().tolist():
                    d_ref = float(dist[b, t, ref_i[b, t]])
                    d_test = float(dist[b, t, test_i[b, t]])
                    scale = float(e[b, t].pow(2).sum() + cb[test_i[b, t]].pow(2).sum())
                    rows.append({"b": b, "t": t, "stage": i, "ref": int(ref_i[b, t]),
                                 "test": int(test_i[b, t]),
                                 "margin_scale": (d_test - d_ref) / max(scale, 1e-30),
                                 "margin_d1": (d_test - d_ref) / max(abs(d_ref), 1e-30)})
            flipped |= new
            zqi, _, _ = model.vq_forward(i, residual)
            residual = residual - zqi
        return {"match_per_stage": match_per_stage, "uncascaded_flips": rows,
                "frames_flipped": int(flipped.sum()), "frames": int(B * T)}
