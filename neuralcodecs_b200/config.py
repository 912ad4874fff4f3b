"""Config classes of the Cuda backend: same fields, defaults, presets and JSON property
names as the reference's config classes (paths under /root/reference/NeuralCodecs.Torch/):
  DACConfig      Config/DAC/DACConfig.cs:8-136
  SNACConfig     Config/SNAC/SNACConfig.cs:11-153
  EncodecConfig  Config/Encodec/EncodecConfig.cs:6-153
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class DeviceConfiguration:
    """NeuralCodecs.Core/Configuration/DeviceConfiguration.cs:6-28 (CUDA only here)."""
    type: str = "CUDA"
    index: int = 0

    @staticmethod
    def CUDA(index: int = 0) -> "DeviceConfiguration":
        return DeviceConfiguration("CUDA", index)


@dataclass
class DACConfig:
    device: DeviceConfiguration = field(default_factory=DeviceConfiguration)
    architecture: str = "dac"
    sample_rate: int = 44100                                   # "sampling_rate"
    encoder_dim: int = 64                                      # "encoder_hidden_size"
    encoder_rates: List[int] = field(default_factory=lambda: [2, 4, 8, 8])   # "downsampling_ratios"
    decoder_dim: int = 1536                                    # "decoder_hidden_size"
    decoder_rates: List[int] = field(default_factory=lambda: [8, 8, 4, 2])   # "upsampling_ratios"
    num_codebooks: int = 9                                     # "n_codebooks"
    codebook_size: int = 1024
    codebook_dim: int = 8
    latent_dim: Optional[int] = None
    quantizer_dropout: float = 0.0
    version: str = "0.0.1"

    _JSON = {"sampling_rate": "sample_rate", "encoder_hidden_size": "encoder_dim",
             "downsampling_ratios": "encoder_rates", "decoder_hidden_size": "decoder_dim",
             "upsampling_ratios": "decoder_rates", "n_codebooks": "num_codebooks",
             "codebook_size": "codebook_size", "codebook_dim": "codebook_dim",
             "latent_dim": "latent_dim", "quantizer_dropout": "quantizer_dropout",
             "model_type": "architecture"}

    @property
    def hop_length(self) -> int:
        return int(math.prod(self.encoder_rates))

    @property
    def resolved_latent_dim(self) -> int:
        # Models/DAC.cs:64
        return self.latent_dim if self.latent_dim else self.encoder_dim * (1 << len(self.encoder_rates))

    @classmethod
    def from_json(cls, text: str) -> "DACConfig":
        cfg = cls()
        for k, v in json.loads(text).items():
            if k in cls._JSON and v is not None:
                setattr(cfg, cls._JSON[k], v)
        return cfg

    @classmethod
    def FromWeights(cls, path: str) -> "DACConfig":
        """DACUnpickler.LoadWithConfig / CreateConfigFromMetadata (Config/DAC/DACUnpickler.cs:383-424): the config stored in an
        official DAC `.pth` checkpoint ({"metadata": {"kwargs": {...}}}); keys and defaults as the reference reads them
        (the reference looks the keys up at the top level of the metadata; descript-audio-codec stores them under
        "kwargs", so both places are searched)."""
        from . import inspect_weights
        meta = inspect_weights(path)["metadata"] or {}
        kw = dict(meta.get("kwargs") or {})
        kw.update({k: v for k, v in meta.items() if k != "kwargs"})
        cfg = cls()
        for key, attr in (("sample_rate", "sample_rate"), ("sampling_rate", "sample_rate"), ("encoder_dim", "encoder_dim"),
                          ("latent_dim", "latent_dim"), ("encoder_rates", "encoder_rates"), ("decoder_rates", "decoder_rates"),
                          ("decoder_dim", "decoder_dim"), ("n_codebooks", "num_codebooks"), ("codebook_size", "codebook_size"),
                          ("codebook_dim", "codebook_dim"), ("quantizer_dropout", "quantizer_dropout")):
            if kw.get(key) is not None:
                v = kw[key]
                setattr(cfg, attr, list(v) if isinstance(v, (list, tuple)) else v)
        return cfg

    # presets: DACConfig.cs:102-136
    @classmethod
    def DAC44kHz(cls) -> "DACConfig":
        return cls()

    @classmethod
    def DAC44kHz_16kbps(cls) -> "DACConfig":
        return cls(num_codebooks=18, latent_dim=128, version="1.0.0")

    @classmethod
    def DAC24kHz(cls) -> "DACConfig":
        return cls(sample_rate=24000, num_codebooks=32, encoder_rates=[2, 4, 5, 8],
                   decoder_rates=[8, 5, 4, 2], version="0.0.4")

    @classmethod
    def DAC16kHz(cls) -> "DACConfig":
        return cls(sample_rate=16000, num_codebooks=12, encoder_rates=[2, 4, 5, 8],
                   decoder_rates=[8, 5, 4, 2], version="0.0.5")


@dataclass
class SNACConfig:
    """Config/SNAC/SNACConfig.cs:11-153 (same defaults = the 44 kHz preset; JSON names in comments)."""
    device: DeviceConfiguration = field(default_factory=DeviceConfiguration)
    sample_rate: int = 44100                                                 # "sampling_rate"
    encoder_dim: int = 64
    encoder_rates: List[int] = field(default_factory=lambda: [2, 3, 8, 8])
    latent_dim: Optional[int] = None
    decoder_dim: int = 1536
    decoder_rates: List[int] = field(default_factory=lambda: [8, 8, 3, 2])
    attn_window_size: Optional[int] = 32
    codebook_size: int = 4096
    codebook_dim: int = 8
    vq_strides: List[int] = field(default_factory=lambda: [8, 4, 2, 1])
    noise: bool = True
    depthwise: bool = True

    _JSON = {"sampling_rate": "sample_rate", "encoder_dim": "encoder_dim", "encoder_rates": "encoder_rates",
             "latent_dim": "latent_dim", "decoder_dim": "decoder_dim", "decoder_rates": "decoder_rates",
             "attn_window_size": "attn_window_size", "codebook_size": "codebook_size", "codebook_dim": "codebook_dim",
             "vq_strides": "vq_strides", "noise": "noise", "depthwise": "depthwise"}

    @property
    def hop_length(self) -> int:
        return int(math.prod(self.encoder_rates))

    @property
    def resolved_latent_dim(self) -> int:      # Models/SNAC.cs:37
        return self.latent_dim if self.latent_dim else self.encoder_dim * (1 << len(self.encoder_rates))

    @classmethod
    def from_json(cls, text: str) -> "SNACConfig":
        cfg = cls()
        for k, v in json.loads(text).items():
            if k in cls._JSON:
                setattr(cfg, cls._JSON[k], v)
        return cfg

    @classmethod
    def SNAC44kHz(cls) -> "SNACConfig":        # SNACConfig.cs:113
        return cls()

    @classmethod
    def SNAC32kHz(cls) -> "SNACConfig":        # SNACConfig.cs:119-133
        return cls(sample_rate=32000)

    @classmethod
    def SNAC24kHz(cls) -> "SNACConfig":        # SNACConfig.cs:139-153
        return cls(sample_rate=24000, encoder_dim=48, encoder_rates=[2, 4, 8, 8], decoder_dim=1024,
                   decoder_rates=[8, 8, 4, 2], attn_window_size=None, vq_strides=[4, 2, 1])


@dataclass
class EncodecConfig:
    """Config/Encodec/EncodecConfig.cs:6-153 (defaults = Encodec24Khz; JSON names follow the HF config)."""
    device: DeviceConfiguration = field(default_factory=DeviceConfiguration)
    sample_rate: int = 24000                    # "sampling_rate"
    channels: int = 1                           # "audio_channels"
    num_filters: int = 32
    hidden_size: int = 128
    upsampling_ratios: List[int] = field(default_factory=lambda: [8, 5, 4, 2])
    num_residual_layers: int = 1
    num_lstm_layers: int = 2
    codebook_size: int = 1024
    codebook_dim: int = 128
    target_bandwidths: List[float] = field(default_factory=lambda: [1.5, 3.0, 6.0, 12.0, 24.0])
    bandwidth: Optional[float] = 6.0
    use_causal_conv: bool = True
    normalize: bool = False
    chunk_length_s: Optional[float] = None      # `Segment`
    overlap: Optional[float] = None             # Models/Encodec.cs:84: `config.Overlap ?? 0`
    norm_type: str = "weight_norm"

    _JSON = {"sampling_rate": "sample_rate", "audio_channels": "channels", "num_filters": "num_filters",
             "hidden_size": "hidden_size", "upsampling_ratios": "upsampling_ratios",
             "num_residual_layers": "num_residual_layers", "num_lstm_layers": "num_lstm_layers",
             "codebook_size": "codebook_size", "codebook_dim": "codebook_dim", "target_bandwidths": "target_bandwidths",
             "use_causal_conv": "use_causal_conv", "normalize": "normalize", "chunk_length_s": "chunk_length_s",
             "overlap": "overlap", "norm_type": "norm_type"}

    @property
    def hop_length(self) -> int:
        return int(math.prod(self.upsampling_ratios))

    @property
    def num_quantizers(self) -> int:            # Models/Encodec.cs:70-71
        return int(1000 * max(self.target_bandwidths) / (math.ceil(self.sample_rate / float(self.hop_length)) * 10))

    @classmethod
    def from_json(cls, text: str) -> "EncodecConfig":
        cfg = cls()
        for k, v in json.loads(text).items():
            if k in cls._JSON:
                setattr(cfg, cls._JSON[k], v)
        return cfg

    @property
    def segment_length(self) -> Optional[int]:  # Models/Encodec.cs:190 (float32 product)
        if self.chunk_length_s is None:
            return None
        import numpy as np
        return int(np.float32(self.chunk_length_s) * np.float32(self.sample_rate))

    @property
    def segment_stride(self) -> Optional[int]:  # Models/Encodec.cs:195-196
        if self.chunk_length_s is None:
            return None
        import numpy as np
        return max(1, int((np.float32(1) - np.float32(self.overlap or 0.0)) * np.float32(self.segment_length)))

    @classmethod
    def Encodec24Khz(cls) -> "EncodecConfig":   # EncodecConfig.cs:9-34
        return cls()

    @classmethod
    def Encodec48Khz(cls) -> "EncodecConfig":   # EncodecConfig.cs:37-66
        return cls(sample_rate=48000, channels=2, target_bandwidths=[3.0, 6.0, 12.0, 24.0], bandwidth=6.0,
                   use_causal_conv=False, normalize=True, chunk_length_s=1.0, overlap=0.01, norm_type="time_group_norm")
