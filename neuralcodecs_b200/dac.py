"""Host-side mirror of the reference's DAC model class over the C ABI.

Same public method names, argument meaning and error behaviour as
/root/reference/NeuralCodecs.Torch/Models/DAC.cs (Encode :163, EncodeAudio :188,
Decode :231, FromCodes :101, forward :262, LoadWeights :345, Dispose :328), with numpy
arrays in place of TorchSharp tensors.  All arithmetic happens in libneuralcodecs_cuda.so.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib
from .config import DACConfig


def _f32(a, name: str) -> np.ndarray:
    if a is None:
        raise TypeError(f"{name} is null")            # ArgumentNullException
    return np.ascontiguousarray(a, dtype=np.float32)


class DAC:
    """INeuralCodec implementation backed by the sm_100a engine (Config, LoadWeights, Dispose)."""

    def __init__(self, config: DACConfig, *, options: Optional[Dict[str, str]] = None):
        if config is None:
            raise TypeError("config is null")          # Models/DAC.cs:53
        self._config = config
        c = _lib.nc_dac_config()
        c.struct_size = C.sizeof(_lib.nc_dac_config)
        c.sample_rate = config.sample_rate
        c.encoder_dim = config.encoder_dim
        er = config.encoder_rates or [2, 4, 8, 8]     # Models/DAC.cs:57
        dr = config.decoder_rates or [8, 8, 4, 2]     # Models/DAC.cs:59
        if len(er) > _lib.NC_MAX_RATES or len(dr) > _lib.NC_MAX_RATES:
            raise ValueError("too many rates")
        c.n_encoder_rates = len(er)
        for i, r in enumerate(er):
            c.encoder_rates[i] = r
        c.decoder_dim = config.decoder_dim
        c.n_decoder_rates = len(dr)
        for i, r in enumerate(dr):
            c.decoder_rates[i] = r
        c.n_codebooks = config.num_codebooks
        c.codebook_size = config.codebook_size
        c.codebook_dim = config.codebook_dim
        c.latent_dim = config.latent_dim or 0
        self._h = C.c_void_p()
        _lib.check(_lib.lib().nc_create(_lib.NC_CODEC_DAC, C.byref(c), C.sizeof(c), config.device.index,
                                        C.byref(self._h)), "DAC", "Create")
        for k, v in (options or {}).items():
            self.set_option(k, v)

    # ------------------------------------------------------------------ INeuralCodec
    @property
    def Config(self) -> DACConfig:
        return self._config

    def LoadWeights(self, path: str) -> None:
        """Models/DAC.cs:345-389: FileNotFoundException / InvalidOperationException on failure."""
        _lib.check(_lib.lib().nc_load_weights(self._handle(), str(path).encode()), "DAC", "LoadWeights")

    def Dispose(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().nc_destroy(self._h)
            self._h = C.c_void_p()

    close = Dispose

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.Dispose()

    def __del__(self):
        try:
            self.Dispose()
        except Exception:
            pass

    # ------------------------------------------------------------------ engine controls
    def set_option(self, key: str, value) -> None:
        _lib.check(_lib.lib().nc_set_option(self._handle(), key.encode(), str(value).encode()), "DAC", "SetOption")

    def set_tensor(self, name: str, array: np.ndarray) -> None:
        a = np.ascontiguousarray(array)
        if a.dtype == np.int64:
            dt = 1
        else:
            a, dt = np.ascontiguousarray(a, dtype=np.float32), 0
        shape = (C.c_int64 * a.ndim)(*a.shape)
        _lib.check(_lib.lib().nc_set_tensor(self._handle(), name.encode(), dt, a.ndim, shape,
                                            a.ctypes.data_as(C.c_void_p)), "DAC", "SetTensor")

    def finalize_weights(self) -> None:
        _lib.check(_lib.lib().nc_finalize_weights(self._handle()), "DAC", "LoadWeights")

    def launch_count(self) -> int:
        return int(_lib.lib().nc_launch_count(self._handle()))

    def profile_report(self) -> dict:
        buf = C.create_string_buffer(1 << 18)
        _lib.check(_lib.lib().nc_profile_report(self._handle(), buf, len(buf)), "DAC", "Profile")
        return json.loads(buf.value.decode())

    def describe(self) -> dict:
        buf = C.create_string_buffer(1 << 18)
        _lib.check(_lib.lib().nc_describe(self._handle(), buf, len(buf)), "DAC", "Describe")
        return json.loads(buf.value.decode())

    def precision_summary(self) -> str:
        d = self.describe()
        wide = d.get("decoder_wide_precision", d["decoder_precision"])
        dec = d["decoder_precision"] if wide == d["decoder_precision"] else \
            f"{wide} on layers wider than 128 channels, {d['decoder_precision']} elsewhere"
        return (f"encoder {d['encoder_precision']}, decoder {dec}"
                + (" (+3xtf32 on narrow 1x1 / final conv)" if d.get("decoder_boost") and d["decoder_precision"] == "tf32" else ""))

    def stream_ptr(self) -> int:
        """cudaStream_t the *_dev calls enqueue on (wrap with torch.cuda.ExternalStream to time them)."""
        s = C.c_void_p()
        _lib.check(_lib.lib().nc_get_stream(self._handle(), C.byref(s)), "DAC", "GetStream")
        return s.value or 0

    def query_shapes(self, length: int) -> Tuple[int, int]:
        """(padded length, frames) for an input of `length` samples (DAC.Preprocess, DAC.cs:141-154)."""
        pl, fr = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().nc_dac_query_shapes(self._handle(), length, C.byref(pl), C.byref(fr), None, None, None))
        return pl.value, fr.value

    # ------------------------------------------------------------------ model surface
    def Encode(self, audioData, nQuantizers: Optional[int] = None, sampleRate: Optional[int] = None,
               *, want_latents: bool = True):
        """DAC.Encode(Tensor, int?, int?) -> (z [B,latent,T], codes [B,nq,T] int64, latents [B,nq*D,T]).

        audioData: [B,1,L] (or [B,L] / [L]).  A 1-D float array is also what the reference's
        Encode(float[]) overload takes (it returns z only; use EncodeAudio for that shape)."""
        a = _f32(audioData, "audioData")
        a = a.reshape(1, -1) if a.ndim == 1 else a.reshape(a.shape[0], -1)
        B, L = a.shape
        cfg = self._config
        nq = cfg.num_codebooks if nQuantizers is None else min(int(nQuantizers), cfg.num_codebooks)
        if nQuantizers is not None and nQuantizers < 1:
            raise ValueError("nQuantizers must be positive")
        _, T = self.query_shapes(L)
        z = np.empty((B, cfg.resolved_latent_dim, T), np.float32)
        codes = np.empty((B, nq, T), np.int64)
        latents = np.empty((B, nq * cfg.codebook_dim, T), np.float32) if want_latents else None
        frames = C.c_int64()
        _lib.check(_lib.lib().nc_dac_encode(
            self._handle(), a.ctypes.data_as(C.c_void_p), B, L, 0 if sampleRate is None else int(sampleRate), nq,
            z.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p),
            latents.ctypes.data_as(C.c_void_p) if latents is not None else None, C.byref(frames)), "DAC", "Encoding")
        return z, codes, latents

    def EncodeAudio(self, audioData) -> np.ndarray:
        """DAC.EncodeAudio / Encode(float[]) (DAC.cs:188-224): quantised latent z only."""
        z, _, _ = self.Encode(audioData, want_latents=False)
        return z

    def Decode(self, qAudio) -> np.ndarray:
        """DAC.Decode (DAC.cs:231-253): z [B,latent,T] (or flat, reshaped to [1,latent,-1]) -> audio
        [B,1,T*hop]; not trimmed to the input length, as in the reference."""
        z = _f32(qAudio, "qAudio")
        D = self._config.resolved_latent_dim
        z = z.reshape(1, D, -1) if z.ndim == 1 else z
        if z.ndim != 3 or z.shape[1] != D:
            raise ValueError(f"expected z of shape [B,{D},T]")
        B, _, T = z.shape
        out_len = T * self._config.hop_length
        audio = np.empty((B, 1, self._decoded_length(T, out_len)), np.float32)
        _lib.check(_lib.lib().nc_dac_decode(self._handle(), z.ctypes.data_as(C.c_void_p), B, T,
                                            audio.ctypes.data_as(C.c_void_p)), "DAC", "Decoding")
        return audio

    def FromCodes(self, codes) -> np.ndarray:
        """DAC.FromCodes (DAC.cs:101-106): codes [B,nq,T] int64 -> z [B,latent,T]."""
        if codes is None:
            raise TypeError("codes is null")
        c = np.ascontiguousarray(codes, dtype=np.int64)
        if c.ndim != 3:
            raise ValueError("expected codes of shape [B,nq,T]")
        B, nq, T = c.shape
        z = np.empty((B, self._config.resolved_latent_dim, T), np.float32)
        _lib.check(_lib.lib().nc_dac_from_codes(self._handle(), c.ctypes.data_as(C.c_void_p), B, nq, T,
                                                z.ctypes.data_as(C.c_void_p)), "DAC", "Decoding")
        return z

    def DecodeCodes(self, codes) -> np.ndarray:
        """Batched Dia.Decode (Models/Dia.cs:973-981): FromCodes + Decode fused on device."""
        if codes is None:
            raise TypeError("codes is null")
        c = np.ascontiguousarray(codes, dtype=np.int64)
        if c.ndim != 3:
            raise ValueError("expected codes of shape [B,nq,T]")
        B, nq, T = c.shape
        audio = np.empty((B, 1, self._decoded_length(T, T * self._config.hop_length)), np.float32)
        _lib.check(_lib.lib().nc_dac_decode_codes(self._handle(), c.ctypes.data_as(C.c_void_p), B, nq, T,
                                                  audio.ctypes.data_as(C.c_void_p)), "DAC", "Decoding")
        return audio

    DIA_DELAY_PATTERN = (0, 8, 9, 10, 11, 12, 13, 14, 15)      # Config/Dia/DataConfig.cs:56

    def DecodeDia(self, generatedCodes, lengths, delayPattern=DIA_DELAY_PATTERN):
        """Batched Dia.GenerateOutput codec stage (Models/Dia.cs:1010-1060): generatedCodes [B,T,C] int64 (delayed, as
        Dia emits them), lengths [B] -> list of per-item audio arrays of lengths[b]*hop samples."""
        if generatedCodes is None or lengths is None:
            raise TypeError("generatedCodes / lengths is null")
        g = np.ascontiguousarray(generatedCodes, dtype=np.int64)
        if g.ndim != 3:
            raise ValueError("expected generatedCodes of shape [B,T,C]")
        B, T, Cc = g.shape
        ln = np.ascontiguousarray(lengths, dtype=np.int64).reshape(-1)
        if ln.shape[0] != B or len(delayPattern) != Cc:
            raise ValueError("lengths / delayPattern do not match generatedCodes")
        dl = (C.c_int32 * Cc)(*[int(d) for d in delayPattern])
        stride = max(int(ln.max()) * self._config.hop_length, 1)
        audio = np.zeros((B, stride), np.float32)
        _lib.check(_lib.lib().nc_dac_decode_dia(self._handle(), g.ctypes.data_as(C.c_void_p), B, T, Cc, dl,
                                                ln.ctypes.data_as(C.POINTER(C.c_int64)), audio.ctypes.data_as(C.c_void_p),
                                                stride), "DAC", "Decoding")
        return [audio[b, : int(ln[b]) * self._config.hop_length].copy() for b in range(B)]

    def forward(self, audioData, sampleRate: Optional[int] = None, nQuantizers: Optional[int] = None) -> dict:
        """DAC.forward (DAC.cs:262-322): {"audio","z","codes"} (losses are constant zeros in the ref)."""
        a = _f32(audioData, "audioData")
        a = a.reshape(1, -1) if a.ndim == 1 else a.reshape(a.shape[0], -1)
        if sampleRate is not None and sampleRate != self._config.sample_rate:
            raise ValueError(f"Input audio sample rate {sampleRate}Hz does not match model sample rate "
                             f"{self._config.sample_rate}Hz")
        B, L = a.shape
        cfg = self._config
        nq = cfg.num_codebooks if nQuantizers is None else min(int(nQuantizers), cfg.num_codebooks)
        Lp, T = self.query_shapes(L)
        audio = np.empty((B, 1, self._decoded_length(T, Lp)), np.float32)
        z = np.empty((B, cfg.resolved_latent_dim, T), np.float32)
        codes = np.empty((B, nq, T), np.int64)
        _lib.check(_lib.lib().nc_dac_forward(
            self._handle(), a.ctypes.data_as(C.c_void_p), B, L, nq, audio.ctypes.data_as(C.c_void_p),
            codes.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), None), "DAC", "Encoding")
        return {"audio": audio, "z": z, "codes": codes}

    # ------------------------------------------------------------------ device-pointer variants
    def forward_dev(self, audio_ptr: int, batch: int, length: int, audio_out_ptr: int, codes_ptr: int,
                    z_ptr: int = 0, n_quantizers: int = 0) -> int:
        """Zero-copy variant: raw CUDA device pointers (e.g. torch.Tensor.data_ptr())."""
        frames = C.c_int64()
        _lib.check(_lib.lib().nc_dac_forward_dev(self._handle(), audio_ptr, batch, length, n_quantizers,
                                                 audio_out_ptr or None, codes_ptr or None, z_ptr or None,
                                                 C.byref(frames)), "DAC", "Encoding")
        return frames.value

    def decode_codes_dev(self, codes_ptr: int, batch: int, n_quantizers: int, frames: int, audio_ptr: int) -> None:
        _lib.check(_lib.lib().nc_dac_decode_codes_dev(self._handle(), codes_ptr, batch, n_quantizers, frames,
                                                      audio_ptr), "DAC", "Decoding")

    # ------------------------------------------------------------------ raw host-pointer variants
    def forward_host(self, audio_ptr: int, batch: int, length: int, audio_out_ptr: int, codes_ptr: int,
                     z_ptr: int = 0, n_quantizers: int = 0) -> int:
        """nc_dac_forward on caller-owned HOST buffers given as raw addresses (e.g. pinned torch
        tensors' data_ptr()); H2D / D2H copies happen inside the call."""
        frames = C.c_int64()
        _lib.check(_lib.lib().nc_dac_forward(self._handle(), audio_ptr, batch, length, n_quantizers,
                                             audio_out_ptr or None, codes_ptr or None, z_ptr or None,
                                             C.byref(frames)), "DAC", "Encoding")
        return frames.value

    def decode_codes_host(self, codes_ptr: int, batch: int, n_quantizers: int, frames: int, audio_ptr: int) -> None:
        _lib.check(_lib.lib().nc_dac_decode_codes(self._handle(), codes_ptr, batch, n_quantizers, frames,
                                                  audio_ptr), "DAC", "Decoding")

    # ------------------------------------------------------------------ helpers
    def _handle(self):
        if not self._h.value:
            raise RuntimeError("DAC has been disposed")  # ObjectDisposedException
        return self._h

    def _decoded_length(self, T: int, default: int) -> int:
        # even strides: T*hop.  Odd strides (24 kHz / 16 kHz presets) lose one sample per odd
        # transposed conv; follow conv_transpose1d's length formula.
        t = T
        for s in (self._config.decoder_rates or [8, 8, 4, 2]):
            p = -(-s // 2)
            t = (t - 1) * s - 2 * p + 2 * s
        return t
