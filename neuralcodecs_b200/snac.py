"""Host-side mirror of the reference's SNAC model class over the C ABI.

Same public method names and argument meaning as /root/reference/NeuralCodecs.Torch/Models/SNAC.cs
(forward :91, Encode :113/:129, Decode :157/:173, LoadWeights :200, ProcessAudio :255), numpy arrays in
place of TorchSharp tensors.  All arithmetic happens in libneuralcodecs_cuda.so.
"""
from __future__ import annotations

import ctypes as C
import os
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .config import SNACConfig



def _seed(seed: Optional[int]) -> int:
    """NoiseBlock noise (Modules/SNAC/NoiseBlock.cs:38-45) is a fresh randn on every call in the reference: with no
    explicit seed each call draws a new 64-bit seed; an explicit seed gives a reproducible realisation that does not
    depend on how the batch is split (the device counter is (clip index, time step))."""
    if seed is None:
        return int.from_bytes(os.urandom(8), "little")
    return int(seed) & 0xFFFFFFFFFFFFFFFF

class SNAC:
    def __init__(self, config: SNACConfig, *, options: Optional[Dict[str, str]] = None):
        if config is None:
            raise TypeError("config is null")
        self._config = config
        c = _lib.nc_snac_config()
        c.struct_size = C.sizeof(_lib.nc_snac_config)
        c.sample_rate, c.encoder_dim, c.decoder_dim = config.sample_rate, config.encoder_dim, config.decoder_dim
        for name, vals in (("encoder_rates", config.encoder_rates), ("decoder_rates", config.decoder_rates),
                           ("vq_strides", config.vq_strides)):
            if len(vals) > _lib.NC_MAX_RATES:
                raise ValueError(f"too many {name}")
            arr = getattr(c, name)
            for i, v in enumerate(vals):
                arr[i] = v
        c.n_encoder_rates, c.n_decoder_rates, c.n_vq_strides = len(config.encoder_rates), len(config.decoder_rates), len(config.vq_strides)
        c.latent_dim = config.latent_dim or 0
        c.attn_window_size = config.attn_window_size or 0
        c.codebook_size, c.codebook_dim = config.codebook_size, config.codebook_dim
        c.noise, c.depthwise = int(config.noise), int(config.depthwise)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().nc_create(_lib.NC_CODEC_SNAC, C.byref(c), C.sizeof(c), config.device.index, C.byref(self._h)),
                   "SNAC", "Create")
        for k, v in (options or {}).items():
            self.set_option(k, v)

    # ------------------------------------------------------------------ INeuralCodec
    @property
    def Config(self) -> SNACConfig:
        return self._config

    def LoadWeights(self, path: str) -> None:
        """Models/SNAC.cs:200-246."""
        _lib.check(_lib.lib().nc_load_weights(self._handle(), str(path).encode()), "SNAC", "LoadWeights")

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
        _lib.check(_lib.lib().nc_set_option(self._handle(), key.encode(), str(value).encode()), "SNAC", "SetOption")

    def set_tensor(self, name: str, array: np.ndarray) -> None:
        a = np.ascontiguousarray(array)
        a, dt = (a, 1) if a.dtype == np.int64 else (np.ascontiguousarray(a, dtype=np.float32), 0)
        shape = (C.c_int64 * a.ndim)(*a.shape)
        _lib.check(_lib.lib().nc_set_tensor(self._handle(), name.encode(), dt, a.ndim, shape, a.ctypes.data_as(C.c_void_p)),
                   "SNAC", "SetTensor")

    def finalize_weights(self) -> None:
        _lib.check(_lib.lib().nc_finalize_weights(self._handle()), "SNAC", "LoadWeights")

    def launch_count(self) -> int:
        return int(_lib.lib().nc_launch_count(self._handle()))

    def profile_report(self) -> dict:
        buf = C.create_string_buffer(1 << 18)
        _lib.check(_lib.lib().nc_profile_report(self._handle(), buf, len(buf)), "SNAC", "Profile")
        return json.loads(buf.value.decode())

    def describe(self) -> dict:
        buf = C.create_string_buffer(1 << 18)
        _lib.check(_lib.lib().nc_describe(self._handle(), buf, len(buf)), "SNAC", "Describe")
        return json.loads(buf.value.decode())

    def stream_ptr(self) -> int:
        s = C.c_void_p()
        _lib.check(_lib.lib().nc_get_stream(self._handle(), C.byref(s)), "SNAC", "GetStream")
        return s.value or 0

    def query_shapes(self, length: int) -> Tuple[int, int, List[int], List[int]]:
        """(padded length, frames, code lengths per stage, noise lengths per decoder block); SNAC.cs:70-80."""
        pl, fr, ns, nn = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        cl, nl = (C.c_int64 * _lib.NC_MAX_RATES)(), (C.c_int64 * _lib.NC_MAX_RATES)()
        _lib.check(_lib.lib().nc_snac_query_shapes(self._handle(), length, C.byref(pl), C.byref(fr), C.byref(ns), cl,
                                                   C.byref(nn), nl))
        return pl.value, fr.value, [cl[i] for i in range(ns.value)], [nl[i] for i in range(nn.value)]

    # ------------------------------------------------------------------ model surface
    @staticmethod
    def _audio2d(audioData) -> np.ndarray:
        if audioData is None:
            raise TypeError("audioData is null")          # ArgumentNullException (SNAC.cs:131)
        a = np.ascontiguousarray(audioData, dtype=np.float32)
        return a.reshape(1, -1) if a.ndim == 1 else a.reshape(a.shape[0], -1)

    def _ptr_array(self, arrays: Optional[Sequence[Optional[np.ndarray]]], n: int):
        arr = (C.c_void_p * max(n, 1))()
        for i in range(n):
            arr[i] = arrays[i].ctypes.data if arrays is not None and arrays[i] is not None else None
        return arr

    def Encode(self, audioData) -> List[np.ndarray]:
        """SNAC.Encode (SNAC.cs:113-150): one int64 code array [B, T_i] per VQ stage (the float[] overload of the
        reference returns the same values cast to float32)."""
        a = self._audio2d(audioData)
        B, L = a.shape
        _, _, clens, _ = self.query_shapes(L)
        codes = [np.empty((B, n), np.int64) for n in clens]
        _lib.check(_lib.lib().nc_snac_encode(self._handle(), a.ctypes.data_as(C.c_void_p), B, L,
                                             self._ptr_array(codes, len(codes))), "SNAC", "Encoding")
        return codes

    def Decode(self, codes, noise: Optional[Sequence[np.ndarray]] = None, seed: Optional[int] = None) -> np.ndarray:
        """SNAC.Decode (SNAC.cs:157-192): audio [B,1,frames*hop], not trimmed.  `noise`: explicit N(0,1) tensors
        [B,1,T_i] per decoder block (None = drawn on the device from `seed`)."""
        if codes is None:
            raise TypeError("codes is null")
        if len(codes) == 0 or any(c is None for c in codes):
            raise ValueError("Codes list cannot be empty or contain null arrays")       # SNAC.cs:177-180
        cs = [np.ascontiguousarray(np.asarray(c).reshape(1, -1) if np.asarray(c).ndim == 1 else c, dtype=np.int64) for c in codes]
        if len(cs) != len(self._config.vq_strides):
            raise ValueError(f"Expected {len(self._config.vq_strides)} codebooks but got {len(cs)}")   # RVQ.FromCodes
        B = cs[0].shape[0]
        frames = cs[-1].shape[1] * self._config.vq_strides[-1]
        _, _, _, nlens = self.query_shapes(frames * self._config.hop_length)
        ns = None
        if noise is not None:
            ns = [np.ascontiguousarray(n, dtype=np.float32).reshape(B, -1) for n in noise]
        out_len = nlens[-1] if nlens else frames * self._config.hop_length
        audio = np.empty((B, 1, out_len), np.float32)
        _lib.check(_lib.lib().nc_snac_decode(self._handle(), self._ptr_array(cs, len(cs)), B, frames,
                                             self._ptr_array(ns, len(nlens)) if ns is not None else None, _seed(seed),
                                             audio.ctypes.data_as(C.c_void_p)), "SNAC", "Decoding")
        return audio

    def forward(self, audioData, noise: Optional[Sequence[np.ndarray]] = None, seed: Optional[int] = None):
        """SNAC.forward (SNAC.cs:91-106) -> (audio [B,1,L] trimmed to the input length, codes)."""
        a = self._audio2d(audioData)
        B, L = a.shape
        _, _, clens, nlens = self.query_shapes(L)
        codes = [np.empty((B, n), np.int64) for n in clens]
        ns = [np.ascontiguousarray(n, dtype=np.float32).reshape(B, -1) for n in noise] if noise is not None else None
        audio = np.empty((B, 1, L), np.float32)
        _lib.check(_lib.lib().nc_snac_forward(self._handle(), a.ctypes.data_as(C.c_void_p), B, L,
                                              self._ptr_array(ns, len(nlens)) if ns is not None else None, _seed(seed),
                                              audio.ctypes.data_as(C.c_void_p), self._ptr_array(codes, len(codes))),
                   "SNAC", "Encoding")
        return audio, codes

    def ProcessAudio(self, audioData, sampleRate: int, noise=None, seed: Optional[int] = None) -> np.ndarray:
        """SNAC.ProcessAudio (SNAC.cs:255-282): linear resample to the model rate if needed (:284-308, on the device),
        forward, flat array of the (resampled) input length.  A [B, L] array is processed as one batch -> [B, L']."""
        if audioData is None or len(audioData) == 0:
            raise ValueError("Audio data cannot be empty")
        a = np.ascontiguousarray(audioData, dtype=np.float32)
        flat = a.ndim == 1
        a = a.reshape(1, -1) if flat else a.reshape(a.shape[0], -1)
        B, L = a.shape
        n = C.c_int64()
        lib = _lib.lib()
        _lib.check(lib.nc_snac_process_audio(self._handle(), None, B, L, int(sampleRate), None, 0, None, 0, C.byref(n)),
                   "SNAC", "ProcessAudio")
        out = np.empty((B, n.value), np.float32)
        ns = [np.ascontiguousarray(z, dtype=np.float32).reshape(B, -1) for z in noise] if noise is not None else None
        nz = self._ptr_array(ns, len(ns)) if ns is not None else None
        _lib.check(lib.nc_snac_process_audio(self._handle(), a.ctypes.data_as(C.c_void_p), B, L, int(sampleRate), nz, _seed(seed),
                                             out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)), "SNAC", "ProcessAudio")
        return out.reshape(-1) if flat else out

    def ResampleAudio(self, x, src: int, dst: int) -> np.ndarray:
        """SNAC.ResampleAudio (SNAC.cs:284-308) = AudioUtils.ResampleLinear (Core/Utils/AudioUtils.cs:329-352):
        linear interpolation in double, last sample held; runs on the model's device."""
        from .audio_utils import ResampleLinear
        return ResampleLinear(self, x, src, dst)

    # ------------------------------------------------------------------ raw host-pointer variant
    def forward_host(self, audio_ptr: int, batch: int, length: int, audio_out_ptr: int, code_ptrs: Optional[Sequence[int]] = None,
                     seed: Optional[int] = None) -> None:
        """nc_snac_forward on caller-owned HOST buffers given as raw addresses (e.g. pinned torch tensors' data_ptr());
        H2D / D2H copies happen inside the call.  Noise is drawn on the device from `seed`."""
        cp = (C.c_void_p * max(len(code_ptrs), 1))(*[p or None for p in code_ptrs]) if code_ptrs else None
        _lib.check(_lib.lib().nc_snac_forward(self._handle(), audio_ptr, batch, length, None, _seed(seed), audio_out_ptr, cp),
                   "SNAC", "Encoding")

    # ------------------------------------------------------------------ device-pointer variant
    def forward_dev(self, audio_ptr: int, batch: int, length: int, audio_out_ptr: int, code_ptrs: Sequence[int],
                    noise_ptrs: Optional[Sequence[int]] = None, seed: Optional[int] = None) -> None:
        cp = (C.c_void_p * max(len(code_ptrs), 1))(*[p or None for p in code_ptrs])
        npz = (C.c_void_p * max(len(noise_ptrs), 1))(*[p or None for p in noise_ptrs]) if noise_ptrs else None
        _lib.check(_lib.lib().nc_snac_forward_dev(self._handle(), audio_ptr, batch, length, npz, _seed(seed), audio_out_ptr or None,
                                                  cp if code_ptrs else None), "SNAC", "Encoding")

    def _handle(self):
        if not self._h.value:
            raise RuntimeError("SNAC has been disposed")
        return self._h
