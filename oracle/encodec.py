"""CPU oracle for the reference's Encodec model: the 24 kHz preset (mono, causal, weight-norm, one frame per clip) and the
48 kHz preset (stereo, non-causal, time_group_norm, 1 s segments with 1 % overlap, per-segment loudness scale).
TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED by the reference (it has no tests); see oracle/__init__.py.

Op-for-op restatement of
  Models/Encodec.cs:213-296,436-489            (Encode / Decode / forward / EncodeFrame / DecodeFrame)
  Modules/Encodec/SEANetEncoder.cs:37-148, SEANetDecoder.cs:40-153, SEANetResnetBlock.cs:30-86
  Modules/Encodec/SConv1d.cs:144-173,245-274, SConvTranspose1d.cs:116-139
  Modules/Encodec/WNConv1d.cs:113-127, WNConvTranspose1d.cs:124-156      (w = (v/||v||) * (g - 1e-7))
  Modules/Encodec/SLSTM.cs:24-57
  Modules/Encodec/ResidualVectorQuantizer.cs:107-157, VectorQuantizer.cs:58-115, EuclideanCodebook.cs:74-182
(paths relative to /root/reference/NeuralCodecs.Torch/).  Weight keys are the reference's:
``encoder.layers.{n}.conv.{weight_g,weight_v,bias}``, ``...block.{1,3}.conv.*``, ``...shortcut.conv.*``,
``encoder.layers.13.lstm.{weight_ih,weight_hh,bias_ih,bias_hh}_l{0,1}``, ``quantizer.layers.{i}.codebook.embed``.
48 kHz additions:
  Config/Encodec/EncodecConfig.cs:37-66          (preset), Models/Encodec.cs:190-196 (SegmentLength / SegmentStride)
  Modules/Encodec/NormConv1d.cs:52-100,136-160, NormConvTranspose1d.cs:37-75   (plain conv, then GroupNorm(1, C, eps 1e-5))
  Modules/Encodec/SConv1d.cs:119-128 / SConvTranspose1d.cs:98-106   (keys ``...conv.{weight,bias}``, ``...norm.{weight,bias}``)
  Models/Encodec.cs:213-235,259-296,436-489      (segment loop, per-frame scale, DecodeFrame * scale, LinearOverlapAdd)
  AudioTools/AudioTensorDSP.cs:161-261           (LinearOverlapAdd)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class EncodecConfig:
    """Config/Encodec/EncodecConfig.cs:6-153 (24 kHz preset = defaults)."""
    sample_rate: int = 24000
    channels: int = 1
    num_filters: int = 32
    hidden_size: int = 128                      # SEANet `dimension` and codebook dim
    upsampling_ratios: List[int] = field(default_factory=lambda: [8, 5, 4, 2])
    num_residual_layers: int = 1
    num_lstm_layers: int = 2
    codebook_size: int = 1024
    target_bandwidths: List[float] = field(default_factory=lambda: [1.5, 3.0, 6.0, 12.0, 24.0])
    bandwidth: float = 6.0
    causal: bool = True
    normalize: bool = False
    norm_type: str = "weight_norm"              # "weight_norm" | "time_group_norm"
    chunk_length_s: Optional[float] = None      # `Segment`: None = one frame per clip
    overlap: Optional[float] = None             # Models/Encodec.cs:84: `config.Overlap ?? 0`

    @classmethod
    def encodec_48khz(cls) -> "EncodecConfig":
        """Config/Encodec/EncodecConfig.cs:37-66."""
        return cls(sample_rate=48000, channels=2, target_bandwidths=[3.0, 6.0, 12.0, 24.0], bandwidth=6.0, causal=False,
                   normalize=True, norm_type="time_group_norm", chunk_length_s=1.0, overlap=0.01)

    @property
    def segment_length(self) -> Optional[int]:  # Models/Encodec.cs:190: (int)(_segment * SampleRate), float32 product
        if self.chunk_length_s is None:
            return None
        return int(np.float32(self.chunk_length_s) * np.float32(self.sample_rate))

    @property
    def segment_stride(self) -> Optional[int]:  # Models/Encodec.cs:195-196: max(1, (int)((1 - overlap) * SegmentLength)), float32
        if self.chunk_length_s is None:
            return None
        return max(1, int((np.float32(1) - np.float32(self.overlap or 0.0)) * np.float32(self.segment_length)))

    @property
    def hop_length(self) -> int:
        return int(math.prod(self.upsampling_ratios))

    @property
    def frame_rate(self) -> int:                # Models/Encodec.cs:86
        return int(math.ceil(np.float32(self.sample_rate) / np.float32(self.hop_length)))

    @property
    def num_quantizers(self) -> int:            # Models/Encodec.cs:70-71
        return int(1000 * max(self.target_bandwidths) / (math.ceil(self.sample_rate / float(self.hop_length)) * 10))

    def n_q_for_bandwidth(self, bandwidth: Optional[float] = None) -> int:
        """ResidualVectorQuantizer.Encode (ResidualVectorQuantizer.cs:133-144)."""
        bw = self.bandwidth if bandwidth is None else bandwidth
        bw_per_q = math.log2(self.codebook_size) * self.frame_rate
        n_q = self.num_quantizers
        if bw is not None and bw > 0:
            n_q = int(max(1, math.floor(np.float32(bw) * 1000 / bw_per_q)))
        return n_q


class EncodecOracle:
    def __init__(self, cfg: EncodecConfig, sd: Dict[str, torch.Tensor], dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}

    # ---------------------------------------------------------------- conv wrappers
    def _w(self, p):
        if self.cfg.norm_type != "weight_norm":  # NormConv1d.cs:52-63: plain Conv1d / ConvTranspose1d
            return self.sd[p + ".conv.weight"]
        v, g = self.sd[p + ".conv.weight_v"], self.sd[p + ".conv.weight_g"]
        v_norm = v.contiguous().pow(2).sum([1, 2], keepdim=True, dtype=self.dtype).sqrt()
        return torch.mul(v.div(v_norm), g.sub(1e-7)).contiguous()

    def _norm(self, p, y):
        """NormConv1d.forward (NormConv1d.cs:87-100): time_group_norm = GroupNorm(1, C, eps 1e-5, affine) over (C, T)."""
        if self.cfg.norm_type == "time_group_norm":
            return F.group_norm(y, 1, self.sd[p + ".norm.weight"], self.sd[p + ".norm.bias"], 1e-5)
        return y

    @staticmethod
    def _extra_padding(length, kernel, stride, padding_total):
        """SConv1d.GetExtraPaddingForConv1d (SConv1d.cs:245-250): note the float32 division."""
        n_frames = np.float32(length - kernel + padding_total) / np.float32(stride) + np.float32(1)
        ideal = (int(math.ceil(n_frames)) - 1) * stride + (kernel - padding_total)
        return ideal - length

    @staticmethod
    def _pad1d(x, left, right):
        """SConv1d.Pad1d (SConv1d.cs:252-267): always reflect; small inputs are zero-extended first."""
        if x.shape[-1] <= max(left, right):
            extra = max(left, right) - x.shape[-1] + 1
            x = F.pad(x, [0, extra])
            return F.pad(x, [left, right], mode="reflect")
        return F.pad(x, [left, right], mode="reflect")

    def sconv1d(self, p, x, k, stride=1, dilation=1):
        """SConv1d.forward (SConv1d.cs:144-173)."""
        length = x.shape[2]
        k_eff = (k - 1) * dilation + 1
        padding_total = k_eff - stride
        extra = self._extra_padding(length, k_eff, stride, padding_total)
        if self.cfg.causal:
            padded = self._pad1d(x, padding_total, extra)
        else:
            right = padding_total // 2
            padded = self._pad1d(x, padding_total - right, right + extra)
        return self._norm(p, F.conv1d(padded, self._w(p), self.sd.get(p + ".conv.bias"), stride, 0, dilation, 1))

    def sconvtr1d(self, p, x, k, stride):
        """SConvTranspose1d.forward (SConvTranspose1d.cs:116-139), trim_right_ratio = 1."""
        y = self._norm(p, F.conv_transpose1d(x, self._w(p), self.sd.get(p + ".conv.bias"), stride=stride))   # norm, THEN trim
        padding_total = k - stride
        if self.cfg.causal:
            right = int(math.ceil(padding_total * 1.0))
            left = padding_total - right
        else:
            right = padding_total // 2
            left = padding_total - right
        return y[..., left:y.shape[-1] - right]

    def resnet(self, p, x):
        """SEANetResnetBlock.forward (SEANetResnetBlock.cs:70-86): shortcut conv + [ELU,k3,ELU,k1]."""
        s = self.sconv1d(p + ".shortcut", x, 1)
        y = F.elu(x)
        y = self.sconv1d(p + ".block.1", y, 3)
        y = F.elu(y)
        y = self.sconv1d(p + ".block.3", y, 1)
        return torch.add(s, y)

    def slstm(self, p, x):
        """SLSTM.forward (SLSTM.cs:40-57): [B,C,T] -> [T,B,C]; nn.LSTM (zero state); + skip; back."""
        perm = x.permute(2, 0, 1).contiguous()
        h = perm
        for layer in range(self.cfg.num_lstm_layers):
            w_ih, w_hh = self.sd[f"{p}.lstm.weight_ih_l{layer}"], self.sd[f"{p}.lstm.weight_hh_l{layer}"]
            b_ih, b_hh = self.sd[f"{p}.lstm.bias_ih_l{layer}"], self.sd[f"{p}.lstm.bias_hh_l{layer}"]
            h = torch._VF.lstm(h, (torch.zeros(1, h.shape[1], w_hh.shape[1], dtype=self.dtype),
                                   torch.zeros(1, h.shape[1], w_hh.shape[1], dtype=self.dtype)),
                               [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, False, False, False)[0]
        return h.add(perm).permute(1, 2, 0)

    # ---------------------------------------------------------------- encoder / decoder
    def encoder(self, x):
        c = self.cfg
        x = self.sconv1d("encoder.layers.0", x, 7)
        idx = 1
        for r in reversed(c.upsampling_ratios):
            for _ in range(c.num_residual_layers):
                x = self.resnet(f"encoder.layers.{idx}", x)
                idx += 1
            x = F.elu(x)
            idx += 1
            x = self.sconv1d(f"encoder.layers.{idx}", x, 2 * r, stride=r)
            idx += 1
        if c.num_lstm_layers > 0:
            x = self.slstm(f"encoder.layers.{idx}", x)
            idx += 1
        x = F.elu(x)
        idx += 1
        return self.sconv1d(f"encoder.layers.{idx}", x, 7)

    def decoder(self, x):
        c = self.cfg
        x = self.sconv1d("decoder.layers.0", x, 7)
        idx = 1
        if c.num_lstm_layers > 0:
            x = self.slstm(f"decoder.layers.{idx}", x)
            idx += 1
        for r in c.upsampling_ratios:
            x = F.elu(x)
            idx += 1
            x = self.sconvtr1d(f"decoder.layers.{idx}", x, 2 * r, r)
            idx += 1
            for _ in range(c.num_residual_layers):
                x = self.resnet(f"decoder.layers.{idx}", x)
                idx += 1
        x = F.elu(x)
        idx += 1
        return self.sconv1d(f"decoder.layers.{idx}", x, 7)

    # ---------------------------------------------------------------- quantizer
    def vq_distances(self, q, flat):
        """EuclideanCodebook.Quantize (EuclideanCodebook.cs:155-182): (x^2 + e^2^T) + (-2 x e^T)."""
        embed = self.sd[f"quantizer.layers.{q}.codebook.embed"]
        x2 = flat.pow(2).sum(1, keepdim=True)
        e2 = embed.pow(2).sum(1, keepdim=True).t()
        neg = -2 * flat.matmul(embed.t())
        return x2.add(e2).add(neg)

    def vq_forward(self, q, x):
        """VectorQuantizer.forward (VectorQuantizer.cs:76-115), eval mode -> (quantized [B,D,T], codes [B,T])."""
        xt = x.transpose(1, 2)
        flat = xt.reshape(-1, xt.shape[-1])
        idx = self.vq_distances(q, flat).argmin(dim=-1).view(xt.shape[:-1])
        quant = F.embedding(idx, self.sd[f"quantizer.layers.{q}.codebook.embed"])
        return quant.transpose(1, 2), idx

    def rvq_encode(self, x, bandwidth: Optional[float] = None):
        """ResidualVectorQuantizer.Encode (ResidualVectorQuantizer.cs:133-157) -> codes [B,nq,T] int64."""
        n_q = self.cfg.n_q_for_bandwidth(bandwidth)
        residual = x.clone()
        codes = []
        for i in range(n_q):
            quant, idx = self.vq_forward(i, residual)
            residual = residual - quant
            codes.append(idx)
        return torch.stack(codes, dim=1)

    def rvq_decode(self, codes):
        """ResidualVectorQuantizer.Decode (ResidualVectorQuantizer.cs:107-124)."""
        out = torch.zeros(1, dtype=self.dtype)
        for i in range(codes.shape[1]):
            quant = F.embedding(codes[:, i], self.sd[f"quantizer.layers.{i}.codebook.embed"]).transpose(1, 2)
            out = out + quant
        return out

    # ---------------------------------------------------------------- model surface (24 kHz: one frame = whole clip)
    def encode(self, audio, bandwidth: Optional[float] = None):
        """Encodec.Encode -> EncodeFrame (Encodec.cs:259-285,457-489), Normalize = false -> codes [B,nq,T]."""
        with torch.inference_mode():
            if audio.dim() != 3:
                raise ValueError(f"Expected 3D input tensor [B,C,T], got shape {list(audio.shape)}")
            if audio.shape[1] != self.cfg.channels:
                raise ValueError(f"Expected {self.cfg.channels} channels, got {audio.shape[1]}")
            emb = self.encoder(audio.to(self.dtype))
            return self.rvq_encode(emb, bandwidth)

    def encode_latent(self, audio):
        with torch.inference_mode():
            return self.encoder(audio.to(self.dtype))

    def decode(self, codes):
        """Encodec.Decode -> DecodeFrame (Encodec.cs:213-235,436-455): not trimmed."""
        with torch.inference_mode():
            return self.decoder(self.rvq_decode(codes))

    def forward(self, audio, bandwidth: Optional[float] = None):
        """Encodec.forward (Encodec.cs:292-296): decode(encode(x)) sliced to the input length."""
        codes = self.encode(audio, bandwidth)
        return {"audio": self.decode(codes)[..., :audio.shape[-1]], "codes": codes}


    # ---------------------------------------------------------------- segmented surface (48 kHz; also valid for 24 kHz)
    def encode_frame(self, x, bandwidth: Optional[float] = None):
        """Encodec.EncodeFrame (Encodec.cs:457-489) -> (codes [B,nq,T], scale [B,1] | None)."""
        c = self.cfg
        if c.chunk_length_s is not None and x.shape[-1] / np.float32(c.sample_rate) > c.chunk_length_s + 1e-5:
            raise ValueError("Frame duration exceeds segment size")
        scale = None
        if c.normalize:
            mono = x.mean([1], keepdim=True)
            volume = mono.pow(2).mean([2], keepdim=True).sqrt()
            scale = volume.add(1e-8)
            x = x.div(scale)
            scale = scale.view(-1, 1)
        return self.rvq_encode(self.encoder(x), bandwidth), scale

    def encode_frames(self, audio, bandwidth: Optional[float] = None):
        """Encodec.Encode(Tensor) (Encodec.cs:259-285): one EncodedFrame per `stride` samples, the last ones shorter."""
        with torch.inference_mode():
            if audio.dim() != 3:
                raise ValueError(f"Expected 3D input tensor [B,C,T], got shape {list(audio.shape)}")
            if audio.shape[1] != self.cfg.channels:
                raise ValueError(f"Expected {self.cfg.channels} channels, got {audio.shape[1]}")
            audio = audio.to(self.dtype)
            length = audio.shape[2]
            seg = self.cfg.segment_length or length
            stride = self.cfg.segment_stride or length
            return [self.encode_frame(audio[:, :, off:min(off + seg, length)], bandwidth) for off in range(0, length, stride)]

    def decode_frame(self, codes, scale=None):
        """Encodec.DecodeFrame (Encodec.cs:436-455)."""
        out = self.decoder(self.rvq_decode(codes))
        if scale is not None:
            out = out * scale.view(-1, 1, 1)
        return out

    def decode_frames(self, frames):
        """Encodec.Decode(List<EncodedFrame>) (Encodec.cs:213-235)."""
        with torch.inference_mode():
            if len(frames) == 0:
                raise ValueError("No frames provided to decode")
            if self.cfg.segment_length is None:
                if len(frames) != 1:
                    raise ValueError("Expected single frame when no segmentation is used")
                return self.decode_frame(*frames[0])
            return linear_overlap_add([self.decode_frame(c, s) for c, s in frames], self.cfg.segment_stride)

    def forward_frames(self, audio, bandwidth: Optional[float] = None):
        """Encodec.forward (Encodec.cs:292-296) through the segmented Encode / Decode."""
        frames = self.encode_frames(audio, bandwidth)
        return {"audio": self.decode_frames(frames)[..., :audio.shape[-1]], "frames": frames}


def linear_overlap_add(frames, stride: int):
    """DSP.LinearOverlapAdd (AudioTools/AudioTensorDSP.cs:161-261): triangular weights 0.5 - |linspace(0,1,L0+2)[1:-1] - 0.5|
    sized by the FIRST frame, a slice of them for shorter frames, sum of weighted frames / sum of weights.  A frame that does
    not fit into stride*(n-1) + len(last) raises, as the reference's `narrow` does."""
    if len(frames) == 0:
        raise ValueError("At least one frame is required")
    dtype = frames[0].dtype
    total = stride * (len(frames) - 1) + frames[-1].shape[-1]
    L0 = frames[0].shape[-1]
    t = torch.linspace(0, 1, L0 + 2, dtype=dtype)[1:-1]
    weight = torch.tensor(0.5, dtype=dtype) - (t - torch.tensor(0.5, dtype=dtype)).abs()
    sum_w = torch.zeros(total, dtype=dtype)
    out = torch.zeros(*frames[0].shape[:-1], total, dtype=dtype)
    off = 0
    for f in frames:
        n = f.shape[-1]
        w = weight.narrow(0, 0, n)
        out.narrow(-1, off, n).add_(f.mul(w))
        sum_w.narrow(0, off, n).add_(w)
        off += stride
    if float(sum_w.min()) <= 1e-10:
        sum_w = sum_w.add(1e-10)
    return out.div(sum_w)


def load_safetensors(path: str, cfg: EncodecConfig, dtype=torch.float32) -> EncodecOracle:
    from safetensors.torch import load_file
    return EncodecOracle(cfg, load_file(path), dtype)


# ------------------------------------------------------------------------------------------------------------------
# .ecdc container without the language model: restatement of Modules/Encodec/BinaryIO.cs, BitPacker.cs, BitUnpacker.cs
# and the useLm == false branches of EncodecCompressor.cs.  Pure Python / numpy; test infrastructure only.
# Parity pin: the reference ships no tests or fixtures for this path ("parity unpinned" by reference vectors); the
# restatement is anchored on hand-derived known-answer vectors (tests/test_oracle_snac_encodec.py) and on the published
# facebookresearch/encodec v0.1.1 format (binary.py: struct '!4sBI' header, BitPacker LSB-first), which it follows.
class BitPacker:
    """BitPacker.cs:60-110: `_currentValue |= value << _currentBits`, emit low bytes while >= 8 bits, Flush() emits the rest."""

    def __init__(self, bits: int):
        if not 0 < bits <= 32:
            raise ValueError("Bits must be between 1 and 32")
        self.bits, self.cur, self.nbits, self.out = bits, 0, 0, bytearray()

    def push(self, value: int) -> None:
        self.cur |= (int(value) & ((1 << self.bits) - 1)) << self.nbits
        self.nbits += self.bits
        while self.nbits >= 8:
            self.out.append(self.cur & 0xFF)
            self.cur >>= 8
            self.nbits -= 8

    def flush(self) -> bytes:
        if self.nbits:
            self.out.append(self.cur & 0xFF)
            self.cur, self.nbits = 0, 0
        return bytes(self.out)


class BitUnpacker:
    """BitUnpacker.cs:60-95: refill a byte at a time until `bits` are available, return the low `bits`; None at end of stream."""

    def __init__(self, bits: int, data: bytes):
        self.bits, self.data, self.pos, self.cur, self.nbits = bits, data, 0, 0, 0

    def pull(self):
        while self.nbits < self.bits:
            if self.pos >= len(self.data):
                return None
            self.cur |= self.data[self.pos] << self.nbits
            self.pos += 1
            self.nbits += 8
        v = self.cur & ((1 << self.bits) - 1)
        self.cur >>= self.bits
        self.nbits -= self.bits
        return v


def _json_number(v: float) -> str:   # System.Text.Json: shortest round-trip text, integers without a fraction
    return str(int(v)) if float(v) == int(v) else repr(float(v))


def ecdc_header(model_name: str, audio_length: int, n_codebooks: int, use_lm: bool, channels: int, sample_rate: int,
                bandwidth: Optional[float]) -> bytes:
    """BinaryIO.WriteHeaderAsync (BinaryIO.cs:152-190) over the metadata of EncodecCompressor.cs:98-111."""
    import struct
    j = (f'{{"m":"{model_name}","al":{int(audio_length)},"nc":{int(n_codebooks)},"lm":{"true" if use_lm else "false"},'
         f'"ch":{int(channels)},"sr":{int(sample_rate)}')
    if bandwidth is not None:
        j += f',"bw":{_json_number(bandwidth)}'
    j += "}"
    meta = j.encode("utf-8")
    return b"ECDC" + bytes([0]) + struct.pack(">i", len(meta)) + meta


def ecdc_read_header(data: bytes):
    """BinaryIO.ReadHeaderAsync + ValidateMetadata (BinaryIO.cs:44-146) -> (metadata dict, payload offset)."""
    import json
    import struct
    if len(data) < 9:
        raise EOFError("Stream ended too soon")
    if data[:4] != b"ECDC":
        raise ValueError("File is not in ECDC format")
    if data[4] != 0:
        raise ValueError(f"Version not supported: {data[4]}")
    (n,) = struct.unpack(">i", data[5:9])
    if n <= 0 or len(data) < 9 + n:
        raise EOFError("Stream ended too soon")
    meta = json.loads(data[9:9 + n].decode("utf-8"))
    for k in ("m", "al", "nc", "lm"):
        if k not in meta:
            raise ValueError(f"Missing required metadata key: {k}")
    return meta, 9 + n


def ecdc_compress_codes(cfg: EncodecConfig, codes, audio_length: int, bandwidth: Optional[float]) -> bytes:
    """EncodecCompressor.CompressToStreamAsync useLm=false (:93-190) for one clip's codes [nq, T] (single frame, no scale)."""
    codes = np.asarray(codes)
    nq, T = codes.shape
    bits = int(np.log2(cfg.codebook_size))
    packer = BitPacker(bits)
    for t in range(T):
        for k in range(nq):
            packer.push(int(codes[k, t]))
    name = "encodec_48khz" if cfg.sample_rate == 48000 else "encodec_24khz"
    return ecdc_header(name, audio_length, nq, False, cfg.channels, cfg.sample_rate, bandwidth) + packer.flush()


def ecdc_decompress_codes(cfg: EncodecConfig, data: bytes):
    """DecompressFromStreamAsync useLm=false (:236-398) up to the code tensor: -> (codes [nq, T] int64, metadata)."""
    import math
    meta, off = ecdc_read_header(data)
    if str(meta["lm"]).lower() == "true":
        raise NotImplementedError("lm streams")
    al, nq = int(meta["al"]), int(meta["nc"])
    T = int(math.ceil(al * cfg.frame_rate / cfg.sample_rate))
    un = BitUnpacker(int(np.log2(cfg.codebook_size)), data[off:])
    codes = np.zeros((nq, T), np.int64)
    for t in range(T):
        for k in range(nq):
            v = un.pull()
            if v is None:
                raise EOFError("Stream ended too soon")
            codes[k, t] = v
    return codes, meta


def ecdc_compress_frames(cfg: EncodecConfig, frames, audio_length: int, bandwidth: Optional[float]) -> bytes:
    """EncodecCompressor.CompressToStreamAsync useLm=false (:93-190) for ONE waveform's frames [(codes [nq, T_s], scale | None)]:
    header, then per frame an optional scale block (int32 BE count = 1, float32 BE value, :116-139) and the frame's codes in a
    BitPacker of its own (t outer, k inner; Flush pads to a byte, :177-187)."""
    import struct
    nq = int(np.asarray(frames[0][0]).shape[0])
    bits = int(np.log2(cfg.codebook_size))
    name = "encodec_48khz" if cfg.sample_rate == 48000 else "encodec_24khz"
    out = bytearray(ecdc_header(name, audio_length, nq, False, cfg.channels, cfg.sample_rate, bandwidth))
    for codes, scale in frames:
        codes = np.asarray(codes)
        if scale is not None:
            out += struct.pack(">i", 1) + struct.pack(">f", float(np.asarray(scale, np.float32).reshape(-1)[0]))
        packer = BitPacker(bits)
        for t in range(codes.shape[1]):
            for k in range(codes.shape[0]):
                packer.push(int(codes[k, t]))
        out += packer.flush()
    return bytes(out)


def ecdc_decompress_frames(cfg: EncodecConfig, data: bytes):
    """DecompressFromStreamAsync useLm=false (:236-400) up to the frame list: -> ([(codes [nq, T_s] int64, scale | None)], metadata).
    Frame lengths follow the READER's formula ceil(segment samples * frame_rate / sample_rate) (:306-309)."""
    import math
    import struct
    meta, pos = ecdc_read_header(data)
    if str(meta["lm"]).lower() == "true":
        raise NotImplementedError("lm streams")
    al, nq = int(meta["al"]), int(meta["nc"])
    seg, stride = cfg.segment_length or al, cfg.segment_stride or al
    bits = int(np.log2(cfg.codebook_size))
    frames = []
    for off in range(0, al, stride):
        T = int(math.ceil(min(al - off, seg) * cfg.frame_rate / float(cfg.sample_rate)))
        scale = None
        if cfg.normalize:
            if len(data) < pos + 4:
                raise EOFError("Stream ended too soon")
            (n,) = struct.unpack(">i", data[pos:pos + 4])
            if n <= 0 or n > 1000:
                raise ValueError(f"Invalid scale count: {n}")
            scale = np.array(struct.unpack(f">{n}f", data[pos + 4:pos + 4 + 4 * n]), np.float32)
            pos += 4 + 4 * n
        nbytes = (T * nq * bits + 7) // 8
        un = BitUnpacker(bits, data[pos:pos + nbytes])
        codes = np.zeros((nq, T), np.int64)
        for t in range(T):
            for k in range(nq):
                v = un.pull()
                if v is None:
                    raise EOFError("Stream ended too soon")
                codes[k, t] = v
        pos += nbytes
        frames.append((codes, scale))
    return frames, meta
