"""`.apr` v1 writer used to serialise random-init models for tests and benchmarks.

Byte layout follows AprWriter::to_bytes / AprWriterInt8::to_bytes of the reference
(src/format/mod.rs:1082-1151, 1290-1359; header :218-245; descriptor :434-458; filterbank
section :961-975; CRC-32 src/format/checksum.rs).  Int4 (quantization byte 3) is this
project's documented extension: the Int8 layout with size_bytes = ceil(n/2) and the nibble
packing of src/model/quantized.rs:1908-1945.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

Q_F32, Q_F16, Q_INT8, Q_INT4 = 0, 1, 2, 3


def _round_half_away(v: np.ndarray) -> np.ndarray:
    return np.sign(v) * np.floor(np.abs(v) + np.float32(0.5))


def quantize_int8(x: np.ndarray):
    """Per-tensor symmetric int8, scale = absmax/127, clamp +-127 (format/mod.rs:849-871)."""
    x = np.asarray(x, np.float32).ravel()
    absmax = np.float32(np.abs(x).max()) if x.size else np.float32(0)
    scale = np.float32(absmax / np.float32(127.0)) if absmax > 0 else np.float32(1.0)
    q = np.clip(_round_half_away((x / scale).astype(np.float32)), -127, 127).astype(np.int8)
    return q, scale


def quantize_int4(x: np.ndarray):
    """Per-tensor symmetric int4, scale = absmax/7, clamp -8..7, even index -> low nibble (quantized.rs:1908-1945)."""
    x = np.asarray(x, np.float32).ravel()
    if x.size == 0:
        return np.zeros(0, np.uint8), np.float32(1.0)
    absmax = np.float32(np.abs(x).max())
    scale = np.float32(1.0) if absmax < np.float32(1e-10) else np.float32(absmax / np.float32(7.0))
    q = np.clip(_round_half_away((x / scale).astype(np.float32)), -8, 7).astype(np.int8)
    nib = q.astype(np.uint8) & 0x0F
    if nib.size % 2:
        nib = np.concatenate([nib, np.zeros(1, np.uint8)])
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8), scale


def write_apr(cfg, tensors, quant: int = Q_F32, filterbank: np.ndarray | None = None) -> bytes:
    """cfg: object with model_type, n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer,
    n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels.  tensors: ordered (name, f32 array)."""
    head = bytearray(48)
    struct.pack_into("<HBBB", head, 0, 1, cfg.model_type, quant, 0)
    struct.pack_into("<H", head, 5, len(tensors))
    head[7] = (1 << 1) if filterbank is not None else 0
    struct.pack_into("<10I", head, 8, cfg.n_vocab, cfg.n_audio_ctx, cfg.n_audio_state, cfg.n_audio_head, cfg.n_audio_layer,
                     cfg.n_text_ctx, cfg.n_text_state, cfg.n_text_head, cfg.n_text_layer, cfg.n_mels)
    index, scales, blobs, offset = [], [], [], 0
    for name, arr in tensors:
        arr = np.asarray(arr, np.float32)
        if quant == Q_F32:
            blob = arr.astype("<f4").tobytes()
        elif quant == Q_INT8:
            q, s = quantize_int8(arr)
            blob = q.tobytes()
            scales.append(struct.pack("<f", float(s)))
        elif quant == Q_INT4:
            q, s = quantize_int4(arr)
            blob = q.tobytes()
            scales.append(struct.pack("<f", float(s)))
        else:
            raise ValueError("quantization not writable")
        d = bytearray(96)
        nb = name.encode()[:47]
        d[: len(nb)] = nb
        struct.pack_into("<QQQ", d, 48, offset, len(blob), arr.size)
        for i, dim in enumerate(arr.shape[:4]):
            struct.pack_into("<I", d, 72 + 4 * i, dim)
        d[88] = min(arr.ndim, 4)
        index.append(bytes(d))
        blobs.append(blob)
        offset += len(blob)
    parts = [b"APR1", bytes(head)] + index + scales + blobs
    if filterbank is not None:
        fb = np.asarray(filterbank, np.float32)
        body = struct.pack("<II", fb.shape[0], fb.shape[1]) + fb.astype("<f4").tobytes()
        parts.append(struct.pack("<I", len(body)) + body)
    crc = 0
    for p in parts:
        crc = zlib.crc32(p, crc)
    parts.append(struct.pack("<I", crc & 0xFFFFFFFF))
    return b"".join(parts)
