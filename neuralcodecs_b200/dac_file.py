"""`.dac` code container of the reference (NeuralCodecs.Torch/AudioTools/DACFile.cs:10-105): the file a caller writes
after `DAC.Encode` and reads before `DAC.Decode` / `FromCodes`.  Pure host I/O over code arrays, so it lives in the
host layer of the backend (the C# twin is bindings/NeuralCodecs.Cuda/CudaDACFile.cs); the codes come from / go to the
device through the codec's own entry points.

Layout (little-endian, as System.IO.BinaryWriter writes it, DACFile.cs:74-103):
    int32   configJson.Length            -- UTF-16 code units of the JSON text (read and ignored by LoadAsync, :33)
    string  configJson                   -- BinaryWriter.Write(string): 7-bit-encoded UTF-8 byte count, then the bytes
    int32   number of code tensors
    per tensor:  int32 rank | int64 dims[rank] | int32 element count | int32 values[count]   (codes cast to int32, :95)
The JSON is System.Text.Json's serialisation of DACConfig (Config/DAC/DACConfig.cs): properties in declaration
order under their [JsonPropertyName], [JsonIgnore] members left out.
"""
from __future__ import annotations

import json
import struct
from typing import List, Sequence

import numpy as np

from .config import DACConfig


def config_to_json(cfg: DACConfig) -> str:
    """JsonSerializer.Serialize(DACConfig) with default options: compact, declaration order (DACConfig.cs:10-99)."""
    d = {
        "Metadata": None,
        "model_type": cfg.architecture,
        "codebook_dim": cfg.codebook_dim,
        "codebook_loss_weight": 1,
        "codebook_size": cfg.codebook_size,
        "commitment_loss_weight": 0.25,
        "decoder_hidden_size": cfg.decoder_dim,
        "upsampling_ratios": list(cfg.decoder_rates),
        "encoder_hidden_size": cfg.encoder_dim,
        "downsampling_ratios": list(cfg.encoder_rates),
        "hop_length": cfg.hop_length,
        "n_codebooks": cfg.num_codebooks,
        "quantizer_dropout": int(cfg.quantizer_dropout) if float(cfg.quantizer_dropout).is_integer() else cfg.quantizer_dropout,
        "sampling_rate": cfg.sample_rate,
        "torch_dtype": "float32",
        "transformers_version": None,
        "hidden_size": 1024,
        "latent_dim": cfg.latent_dim,
        "ChunkSeconds": 10,
    }
    return json.dumps(d, separators=(",", ":"))


def _write_7bit(n: int) -> bytes:
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _read_7bit(buf: memoryview, pos: int):
    n = shift = 0
    while True:
        if pos >= len(buf) or shift > 35:
            raise ValueError("bad 7-bit encoded string length")
        b = buf[pos]
        pos += 1
        n |= (b & 0x7F) << shift
        if not b & 0x80:
            return n, pos
        shift += 7


class DACFile:
    """DACFile(codes, config) (DACFile.cs:14-21): `codes` is a list of integer arrays of any shape (DAC.Encode's
    codes tensor [B, n_codebooks, T], or one array per chunk)."""

    def __init__(self, codes: Sequence[np.ndarray], config: DACConfig):
        self.Codes: List[np.ndarray] = [np.asarray(c) for c in codes]
        self.Config = config

    # DACFile.SaveAsync (DACFile.cs:72-103)
    def Save(self, path: str) -> None:
        js = config_to_json(self.Config)
        utf8 = js.encode("utf-8")
        parts = [struct.pack("<i", len(js.encode("utf-16-le")) // 2), _write_7bit(len(utf8)), utf8,
                 struct.pack("<i", len(self.Codes))]
        for code in self.Codes:
            parts.append(struct.pack("<i", code.ndim))
            parts.append(struct.pack(f"<{code.ndim}q", *code.shape))
            data = np.ascontiguousarray(code).astype("<i4", casting="unsafe").reshape(-1)   # code.to(int32), row-major
            parts.append(struct.pack("<i", data.size))
            parts.append(data.tobytes())
        with open(path, "wb") as f:
            f.write(b"".join(parts))

    # DACFile.LoadAsync (DACFile.cs:27-62)
    @classmethod
    def Load(cls, path: str) -> "DACFile":
        with open(path, "rb") as f:
            buf = memoryview(f.read())
        pos = 4                                                     # configLength: read and not used (:33)
        if len(buf) < 4:
            raise EOFError("Unable to read beyond the end of the stream")
        n, pos = _read_7bit(buf, pos)
        if pos + n > len(buf):
            raise EOFError("Unable to read beyond the end of the stream")
        cfg = DACConfig.from_json(bytes(buf[pos:pos + n]).decode("utf-8"))
        pos += n

        def take(fmt):
            nonlocal pos
            size = struct.calcsize(fmt)
            if pos + size > len(buf):
                raise EOFError("Unable to read beyond the end of the stream")
            v = struct.unpack_from(fmt, buf, pos)
            pos += size
            return v

        codes = []
        (count,) = take("<i")
        for _ in range(count):
            (rank,) = take("<i")
            if rank < 0 or rank > 16:
                raise ValueError("bad tensor rank in .dac file")
            shape = take(f"<{rank}q")
            (m,) = take("<i")
            if m < 0 or pos + 4 * m > len(buf):
                raise EOFError("Unable to read beyond the end of the stream")
            data = np.frombuffer(buf, dtype="<i4", count=m, offset=pos).astype(np.int64)   # torch.tensor(int[]) -> indices
            pos += 4 * m
            codes.append(data.reshape(shape))                       # reshape(shape): raises when the counts disagree
        return cls(codes, cfg)
