"""Host-side mirror of the reference's Encodec model class over the C ABI (24 kHz and 48 kHz presets).

Same public method names and argument meaning as /root/reference/NeuralCodecs.Torch/Models/Encodec.cs
(Encode :243/:259, Decode :213, forward :292, LoadWeights :348), numpy arrays in place of TorchSharp tensors;
an EncodedFrame is the pair (codes [B,nq,T] int64, scale or None) as in Modules/Encodec/EncodedFrame.cs:8.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .config import EncodecConfig

EncodedFrame = Tuple[np.ndarray, Optional[np.ndarray]]


class Encodec:
    def __init__(self, config: EncodecConfig, *, options: Optional[Dict[str, str]] = None):
        if config is None:
            raise TypeError("config is null")
        if config.bandwidth is None or float(config.bandwidth) not in [float(b) for b in config.target_bandwidths]:
            raise ValueError(f"Invalid bandwidth {config.bandwidth}. Select one of {config.target_bandwidths}")   # Encodec.cs:47-53
        self._config = config
        c = _lib.nc_encodec_config()
        c.struct_size = C.sizeof(_lib.nc_encodec_config)
        c.sample_rate, c.channels, c.n_filters, c.dimension = config.sample_rate, config.channels, config.num_filters, config.hidden_size
        c.n_ratios = len(config.upsampling_ratios)
        for i, r in enumerate(config.upsampling_ratios):
            c.ratios[i] = r
        c.n_residual_layers, c.lstm_layers = config.num_residual_layers, config.num_lstm_layers
        c.codebook_size, c.n_quantizers, c.causal = config.codebook_size, config.num_quantizers, int(config.use_causal_conv)
        norms = {"weight_norm": 0, "time_group_norm": 1}
        if config.norm_type not in norms:
            raise ValueError(f"Unsupported normalization: {config.norm_type}")              # NormConv1d.cs:156-157
        c.norm_type, c.normalize = norms[config.norm_type], int(bool(config.normalize))
        c.segment_s = float(config.chunk_length_s) if config.chunk_length_s else 0.0
        c.overlap = float(config.overlap or 0.0)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().nc_create(_lib.NC_CODEC_ENCODEC, C.byref(c), C.sizeof(c), config.device.index, C.byref(self._h)),
                   "Encodec", "Create")
        for k, v in (options or {}).items():
            self.set_option(k, v)

    # ------------------------------------------------------------------ INeuralCodec
    @property
    def Config(self) -> EncodecConfig:
        return self._config

    @property
    def Bandwidth(self) -> float:
        return float(self._config.bandwidth)

    def SetTargetBandwidth(self, bandwidth: float) -> None:
        """Encodec.SetTargetBandwidth (Encodec.cs:409-420)."""
        if float(bandwidth) not in [float(b) for b in self._config.target_bandwidths]:
            raise ValueError(f"This model doesn't support the bandwidth {bandwidth}. Select one of {self._config.target_bandwidths}")
        self._config.bandwidth = float(bandwidth)

    def LoadWeights(self, path: str) -> None:
        _lib.check(_lib.lib().nc_load_weights(self._handle(), str(path).encode()), "Encodec", "LoadWeights")

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
        _lib.check(_lib.lib().nc_set_option(self._handle(), key.encode(), str(value).encode()), "Encodec", "SetOption")

    def set_tensor(self, name: str, array: np.ndarray) -> None:
        a = np.ascontiguousarray(array)
        a, dt = (a, 1) if a.dtype == np.int64 else (np.ascontiguousarray(a, dtype=np.float32), 0)
        shape = (C.c_int64 * a.ndim)(*a.shape)
        _lib.check(_lib.lib().nc_set_tensor(self._handle(), name.encode(), dt, a.ndim, shape, a.ctypes.data_as(C.c_void_p)),
                   "Encodec", "SetTensor")

    def finalize_weights(self) -> None:
        _lib.check(_lib.lib().nc_finalize_weights(self._handle()), "Encodec", "LoadWeights")

    def launch_count(self) -> int:
        return int(_lib.lib().nc_launch_count(self._handle()))

    def profile_report(self) -> dict:
        buf = C.create_string_buffer(1 << 18)
        _lib.check(_lib.lib().nc_profile_report(self._handle(), buf, len(buf)), "Encodec", "Profile")
        return json.loads(buf.value.decode())

    def describe(self) -> dict:
        buf = C.create_string_buffer(1 << 18)
        _lib.check(_lib.lib().nc_describe(self._handle(), buf, len(buf)), "Encodec", "Describe")
        return json.loads(buf.value.decode())

    def stream_ptr(self) -> int:
        s = C.c_void_p()
        _lib.check(_lib.lib().nc_get_stream(self._handle(), C.byref(s)), "Encodec", "GetStream")
        return s.value or 0

    def query_shapes(self, length: int) -> Tuple[int, int, int]:
        """(frames, n_q at the configured bandwidth, decoded length)."""
        fr, nq, dl = C.c_int64(), C.c_int32(), C.c_int64()
        _lib.check(_lib.lib().nc_encodec_query_shapes(self._handle(), length, float(self._config.bandwidth), C.byref(fr),
                                                      C.byref(nq), C.byref(dl)))
        return fr.value, nq.value, dl.value

    # ------------------------------------------------------------------ model surface
    def _audio3d(self, x) -> np.ndarray:
        if x is None:
            raise TypeError("audioData is null")
        a = np.ascontiguousarray(x, dtype=np.float32)
        if a.ndim == 1:
            a = a.reshape(1, self._config.channels, -1)                    # Encode(float[]) (Encodec.cs:247-250)
        if a.ndim != 3:
            raise ValueError(f"Expected 3D input tensor [B,C,T], got shape {list(a.shape)}")   # Encodec.cs:493-497
        if a.shape[1] != self._config.channels:
            raise ValueError(f"Expected {self._config.channels} channels, got {a.shape[1]}")   # Encodec.cs:499-503
        return a

    def query_frames(self, length: int) -> Tuple[List[int], int, int]:
        """(code frames of each segment, n_q at the configured bandwidth, length Decode returns)."""
        n, nq, tot, dl = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        bw = float(self._config.bandwidth)
        _lib.check(_lib.lib().nc_encodec_query_frames(self._handle(), length, bw, C.byref(n), None, 0, C.byref(tot), C.byref(nq),
                                                      C.byref(dl)), "Encodec", "Encoding")
        seg = (C.c_int64 * n.value)()
        _lib.check(_lib.lib().nc_encodec_query_frames(self._handle(), length, bw, C.byref(n), seg, n.value, None, None, None),
                   "Encodec", "Encoding")
        return list(seg), nq.value, dl.value

    def Encode(self, audioData) -> List[EncodedFrame]:
        """Encodec.Encode (Encodec.cs:243-285): one EncodedFrame (codes [B,nq,T_s], scale [B,1] | None) per segment; the
        24 kHz preset has no segmenting -> one frame for the whole clip.  All segments run in one batched device call."""
        a = self._audio3d(audioData)
        B, _, L = a.shape
        seg, nq, _ = self.query_frames(L)
        codes = np.empty((B, nq, sum(seg)), np.int64)
        scales = np.empty((B, len(seg)), np.float32) if self._config.normalize else None
        _lib.check(_lib.lib().nc_encodec_encode_frames(self._handle(), a.ctypes.data_as(C.c_void_p), B, L, float(self._config.bandwidth),
                                                       codes.ctypes.data_as(C.c_void_p),
                                                       scales.ctypes.data_as(C.c_void_p) if scales is not None else None),
                   "Encodec", "Encoding")
        frames, col = [], 0
        for s, t in enumerate(seg):
            frames.append((np.ascontiguousarray(codes[:, :, col:col + t]), scales[:, s:s + 1].copy() if scales is not None else None))
            col += t
        return frames

    def Decode(self, encodedFrames: List[EncodedFrame]) -> np.ndarray:
        """Encodec.Decode (Encodec.cs:213-235): DecodeFrame (* scale) of every frame, then LinearOverlapAdd for segmented
        models; audio [B, channels, stride*(n-1) + len(last)], not trimmed."""
        if encodedFrames is None or len(encodedFrames) == 0:
            raise ValueError("No frames provided to decode")
        if self._config.chunk_length_s is None and len(encodedFrames) != 1:
            raise ValueError("Expected single frame when no segmentation is used")
        if any(f[0] is None for f in encodedFrames):
            raise ValueError("Invalid frame codes in Encodec Decode")
        cs = [np.ascontiguousarray(f[0], dtype=np.int64) for f in encodedFrames]
        B, nq = cs[0].shape[0], cs[0].shape[1]
        codes = np.ascontiguousarray(np.concatenate(cs, axis=2))
        seg = (C.c_int64 * len(cs))(*[c.shape[2] for c in cs])
        with_scale = [f[1] is not None for f in encodedFrames]
        if any(with_scale) and not all(with_scale):
            raise ValueError("frames with and without a scale cannot be mixed in one call")
        scales = None
        if all(with_scale):
            scales = np.ascontiguousarray(np.stack([np.asarray(f[1], np.float32).reshape(B) for f in encodedFrames], axis=1))
        total = C.c_int64()
        _lib.check(_lib.lib().nc_encodec_query_decoded(self._handle(), seg, len(cs), C.byref(total)), "Encodec", "Decoding")
        total = total.value
        audio = np.empty((B, self._config.channels, total), np.float32)
        _lib.check(_lib.lib().nc_encodec_decode_frames(self._handle(), codes.ctypes.data_as(C.c_void_p),
                                                       scales.ctypes.data_as(C.c_void_p) if scales is not None else None, B, nq,
                                                       seg, len(cs), audio.ctypes.data_as(C.c_void_p)), "Encodec", "Decoding")
        return audio

    def forward(self, x) -> np.ndarray:
        """Encodec.forward (Encodec.cs:292-296): decode(encode(x)) sliced to the input length."""
        a = self._audio3d(x)
        B, Cn, L = a.shape
        out = np.empty((B, Cn, L), np.float32)
        _lib.check(_lib.lib().nc_encodec_forward(self._handle(), a.ctypes.data_as(C.c_void_p), B, L, float(self._config.bandwidth),
                                                 out.ctypes.data_as(C.c_void_p), None), "Encodec", "Encoding")
        return out

    def forward_host(self, audio_ptr: int, batch: int, length: int, audio_out_ptr: int, codes_ptr: int = 0) -> None:
        """nc_encodec_forward on caller-owned HOST buffers given as raw addresses (e.g. pinned torch tensors' data_ptr())."""
        _lib.check(_lib.lib().nc_encodec_forward(self._handle(), audio_ptr, batch, length, float(self._config.bandwidth),
                                                 audio_out_ptr, codes_ptr or None), "Encodec", "Encoding")

    def forward_dev(self, audio_ptr: int, batch: int, length: int, audio_out_ptr: int, codes_ptr: int = 0) -> None:
        _lib.check(_lib.lib().nc_encodec_forward_dev(self._handle(), audio_ptr, batch, length, float(self._config.bandwidth),
                                                     audio_out_ptr or None, codes_ptr or None), "Encodec", "Encoding")

    def _handle(self):
        if not self._h.value:
            raise RuntimeError("Encodec has been disposed")
        return self._h


class EncodecCompressor:
    """Mirror of the reference's static EncodecCompressor (Modules/Encodec/EncodecCompressor.cs) for the path without the
    language model: Compress :26-39 / CompressToStreamAsync :60-200, Decompress :46-52 / DecompressFromStreamAsync :236-420.
    The reference builds the decoding model from the stream's "m" key through its model factories (:275-279); here the
    caller passes the loaded model.  The *Batch variants take [B,1,L] audio / a list of equal-metadata streams and run as
    one device batch."""

    @staticmethod
    def Compress(model: "Encodec", wav, useLm: bool = False) -> bytes:
        if useLm:
            raise NotImplementedError("the language-model entropy coder is not built (NC_UNSUPPORTED)")
        a = np.ascontiguousarray(wav, dtype=np.float32)
        if a.ndim != 2:
            raise ValueError("Only single waveform can be encoded (shape should be [C, L])")        # EncodecCompressor.cs:67-70
        if a.shape[0] != model.Config.channels:
            raise ValueError(f"Expected {model.Config.channels} channels, got {a.shape[0]}")        # :74-77
        return EncodecCompressor.CompressBatch(model, a[None])[0]

    @staticmethod
    def CompressBatch(model: "Encodec", wav) -> List[bytes]:
        a = model._audio3d(wav)
        B, _, L = a.shape
        bw = float(model.Config.bandwidth)
        hb, nb = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().nc_encodec_ecdc_size(model._handle(), L, bw, C.byref(hb), C.byref(nb)), "Encodec", "Compress")
        out = np.zeros((B, nb.value), np.uint8)
        wrote = C.c_int64()
        _lib.check(_lib.lib().nc_encodec_compress(model._handle(), a.ctypes.data_as(C.c_void_p), B, L, bw,
                                                  out.ctypes.data_as(C.c_void_p), nb.value, C.byref(wrote)), "Encodec", "Compress")
        return [out[b, :wrote.value].tobytes() for b in range(B)]

    @staticmethod
    def ReadHeader(compressed: bytes) -> dict:
        """BinaryIO.ReadHeaderAsync (BinaryIO.cs:44-100) -> {m-less} metadata + payload offset."""
        buf = np.frombuffer(compressed, np.uint8)
        al, off = C.c_int64(), C.c_int64()
        nq, ch, sr, lm = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        bw = C.c_float()
        _lib.check(_lib.lib().nc_encodec_ecdc_info(buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(al), C.byref(nq), C.byref(ch),
                                                   C.byref(sr), C.byref(bw), C.byref(lm), C.byref(off)), "Encodec", "Decompress")
        return {"al": al.value, "nc": nq.value, "ch": ch.value, "sr": sr.value, "bw": bw.value, "lm": bool(lm.value),
                "payload_offset": off.value}

    @staticmethod
    def Decompress(compressed: bytes, model: "Encodec") -> Tuple[np.ndarray, int]:
        """-> (wav [C, al], sample rate), as DecompressFromStreamAsync returns (wav[0], model.SampleRate)."""
        wav, sr = EncodecCompressor.DecompressBatch([compressed], model)
        return wav[0], sr

    @staticmethod
    def DecompressBatch(streams: List[bytes], model: "Encodec") -> Tuple[np.ndarray, int]:
        if streams is None or len(streams) == 0:
            raise ValueError("No frames provided to decode")
        n = len(streams[0])
        if any(len(s) != n for s in streams):
            raise ValueError("batched decompress needs streams with identical metadata")
        buf = np.frombuffer(b"".join(streams), np.uint8).reshape(len(streams), n)
        al, sr = C.c_int64(), C.c_int32()
        lib = _lib.lib()
        _lib.check(lib.nc_encodec_decompress(model._handle(), buf.ctypes.data_as(C.c_void_p), len(streams), n, n, None, 0,
                                             C.byref(al), C.byref(sr)), "Encodec", "Decompress")
        audio = np.empty((len(streams), model.Config.channels, al.value), np.float32)
        _lib.check(lib.nc_encodec_decompress(model._handle(), buf.ctypes.data_as(C.c_void_p), len(streams), n, n,
                                             audio.ctypes.data_as(C.c_void_p), al.value, C.byref(al), C.byref(sr)),
                   "Encodec", "Decompress")
        return audio, sr.value
