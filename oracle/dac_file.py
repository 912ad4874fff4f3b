"""Oracle restatement of the reference's `.dac` container (AudioTools/DACFile.cs:27-103).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED by the reference (no fixtures): anchored on System.IO.BinaryWriter's documented encodings
(little-endian Int32 / Int64, string = 7-bit-encoded UTF-8 byte count + bytes) and on hand-derived known answers in
tests/test_dac_file.py.  Written stream-wise, one BinaryWriter / BinaryReader call per line of the C#."""
from __future__ import annotations

import io
import json
import struct
from typing import List, Tuple

import numpy as np


def _w_i32(s, v): s.write(struct.pack("<i", v))                      # BinaryWriter.Write(int)
def _w_i64(s, v): s.write(struct.pack("<q", v))                      # BinaryWriter.Write(long)


def _w_str(s, text):                                                  # BinaryWriter.Write(string)
    b = text.encode("utf-8")
    n = len(b)
    while n >= 0x80:
        s.write(bytes([(n | 0x80) & 0xFF]))
        n >>= 7
    s.write(bytes([n]))
    s.write(b)


def save(config_json: str, codes: List[np.ndarray]) -> bytes:
    """DACFile.SaveAsync (:72-103) given the already serialised config JSON."""
    s = io.BytesIO()
    _w_i32(s, len(config_json))                                       # writer.Write(configJson.Length)   :80
    _w_str(s, config_json)                                            # writer.Write(configJson)          :81
    _w_i32(s, len(codes))                                             # writer.Write(Codes.Count)         :84
    for code in codes:
        _w_i32(s, code.ndim)                                          # :89
        for dim in code.shape:
            _w_i64(s, int(dim))                                       # :92
        data = [int(np.int32(v)) for v in np.asarray(code).reshape(-1)]   # code.cpu().to(int32)           :96
        _w_i32(s, len(data))                                          # :97
        for v in data:
            _w_i32(s, v)                                              # :100
    return s.getvalue()


def load(blob: bytes) -> Tuple[dict, List[np.ndarray]]:
    """DACFile.LoadAsync (:27-62) -> (parsed config JSON, code arrays)."""
    s = io.BytesIO(blob)

    def r_i32(): return struct.unpack("<i", s.read(4))[0]
    def r_i64(): return struct.unpack("<q", s.read(8))[0]

    def r_str():
        n = shift = 0
        while True:
            b = s.read(1)[0]
            n |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        return s.read(n).decode("utf-8")

    r_i32()                                                           # configLength (unused)               :33
    cfg = json.loads(r_str())                                         # :34-35
    codes = []
    for _ in range(r_i32()):                                          # :39
        shape = [r_i64() for _ in range(r_i32())]                     # :43-47
        data = [r_i32() for _ in range(r_i32())]                      # :49-54
        codes.append(np.asarray(data, dtype=np.int64).reshape(shape))  # tensor(data).reshape(shape)         :56
    return cfg, codes
