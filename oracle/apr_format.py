"""Oracle: `.apr` v1 container exactly as whisper.apr reads/writes it (CPU, numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to the reference checkout):
  src/format/mod.rs:64,1520-1530   magic "APR1"
  src/format/mod.rs:162-245        AprHeader::{parse,to_bytes} (48 bytes)
  src/format/mod.rs:352-458        TensorDescriptor::{new,parse,to_bytes} (96 bytes)
  src/format/mod.rs:484-522        AprReader::new (index, Int8 scale table offset)
  src/format/mod.rs:610-672        load_tensor / read_int8_tensor_dequantized
  src/format/mod.rs:736-780        read_mel_filterbank
  src/format/mod.rs:849-871        QuantizedTensorData::from_f32 (per-tensor int8, +-127)
  src/format/mod.rs:961-1004       MelFilterbankData::{to_bytes,from_bytes}
  src/format/mod.rs:1082-1151      AprWriter::to_bytes (f32)
  src/format/mod.rs:1290-1359      AprWriterInt8::to_bytes
  src/format/checksum.rs:20-120    CRC-32 (IEEE 802.3, reflected 0xEDB88320) == zlib.crc32
  src/model/quantized.rs:1887-1969 packed int4 (scale=absmax/7, clamp -8..7, even index -> low nibble)

Int4 (quantization byte 3) has NO reader/writer in the reference
(format/mod.rs:619-628 treats anything != Int8 as f32).  This build DEFINES it as
the Int8 layout (index, f32 scale table, data) with size_bytes = ceil(n/2) and the
nibble packing of model/quantized.rs -- a documented extension.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAGIC = b"APR1"
HEADER_SIZE = 48
DESC_SIZE = 96
Q_F32, Q_F16, Q_INT8, Q_INT4 = 0, 1, 2, 3


def crc32(data: bytes) -> int:
    return zlib.crc32(data) & 0xFFFFFFFF


def header_bytes(cfg, quant: int, n_tensors: int, has_vocab=False, has_filterbank=False, compressed=False) -> bytes:
    b = bytearray(HEADER_SIZE)
    struct.pack_into("<HBBB", b, 0, 1, cfg.model_type, quant, int(compressed))
    struct.pack_into("<H", b, 5, n_tensors)
    b[7] = int(has_vocab) | (int(has_filterbank) << 1)
    struct.pack_into("<10I", b, 8, cfg.n_vocab, cfg.n_audio_ctx, cfg.n_audio_state, cfg.n_audio_head,
                     cfg.n_audio_layer, cfg.n_text_ctx, cfg.n_text_state, cfg.n_text_head, cfg.n_text_layer, cfg.n_mels)
    return bytes(b)


def parse_header(b: bytes) -> dict:
    if len(b) < HEADER_SIZE:
        raise ValueError("header too short")
    version, model_type, quant, compressed = struct.unpack_from("<HBBB", b, 0)
    if version > 1:
        raise ValueError(f"unsupported format version: {version}")
    if quant > 3:
        raise ValueError("invalid quantization")
    (n_tensors,) = struct.unpack_from("<H", b, 5)
    flags = b[7]
    vals = struct.unpack_from("<10I", b, 8)
    keys = ["n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
            "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer", "n_mels"]
    h = dict(zip(keys, vals))
    h.update(version=version, model_type=model_type, quantization=quant, compressed=bool(compressed),
             n_tensors=n_tensors, has_vocab=bool(flags & 1), has_filterbank=bool(flags & 2))
    return h


def desc_bytes(name: str, shape, offset: int, size: int, n_elements=None) -> bytes:
    b = bytearray(DESC_SIZE)
    nb = name.encode()[:47]
    b[: len(nb)] = nb
    if n_elements is None:
        n_elements = int(np.prod(shape)) if len(shape) else 1
    struct.pack_into("<QQQ", b, 48, offset, size, n_elements)
    for i, dim in enumerate(list(shape)[:4]):
        struct.pack_into("<I", b, 72 + 4 * i, dim)
    b[88] = min(len(shape), 4)
    return bytes(b)


def parse_desc(b: bytes) -> dict:
    if len(b) < DESC_SIZE:
        raise ValueError("tensor descriptor too short")
    name = b[:48].split(b"\0", 1)[0].decode("utf-8", "replace")
    offset, size, n_elements = struct.unpack_from("<QQQ", b, 48)
    shape = struct.unpack_from("<4I", b, 72)
    n_dims = b[88]
    return dict(name=name, offset=offset, size=size, n_elements=n_elements, shape=shape[:n_dims], n_dims=n_dims)


def quantize_int8(x: np.ndarray):
    """QuantizedTensorData::from_f32 (format/mod.rs:849-871)."""
    x = np.asarray(x, np.float32).ravel()
    absmax = np.float32(np.abs(x).max()) if x.size else np.float32(0)
    scale = np.float32(absmax / np.float32(127.0)) if absmax > 0 else np.float32(1.0)
    v = (x / scale).astype(np.float32)
    q = np.sign(v) * np.floor(np.abs(v) + np.float32(0.5))  # Rust f32::round = half away from zero
    return np.clip(q, -127, 127).astype(np.int8), scale


def quantize_int4(x: np.ndarray):
    """quantize_f32_to_i4_packed (model/quantized.rs:1908-1945)."""
    x = np.asarray(x, np.float32).ravel()
    if x.size == 0:
        return np.zeros(0, np.uint8), np.float32(1.0)
    absmax = np.float32(np.abs(x).max())
    scale = np.float32(1.0) if absmax < np.float32(1e-10) else np.float32(absmax / np.float32(7.0))
    v = (x / scale).astype(np.float32)
    q = np.clip(np.sign(v) * np.floor(np.abs(v) + np.float32(0.5)), -8, 7).astype(np.int8)
    nib = (q.astype(np.uint8) & 0x0F)
    if nib.size % 2:
        nib = np.concatenate([nib, np.zeros(1, np.uint8)])
    packed = (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)
    return packed, scale


def dequantize_int4(packed: np.ndarray, scale, n: int) -> np.ndarray:
    """dequantize_i4_packed_to_f32 (model/quantized.rs:1949-1969)."""
    packed = np.asarray(packed, np.uint8)
    lo = (packed & 0x0F).astype(np.int8)
    hi = (packed >> 4).astype(np.int8)
    lo = np.where(lo >= 8, lo - 16, lo)
    hi = np.where(hi >= 8, hi - 16, hi)
    q = np.empty(packed.size * 2, np.int8)
    q[0::2], q[1::2] = lo, hi
    return q[:n].astype(np.float32) * np.float32(scale)


def filterbank_section(filters: np.ndarray) -> bytes:
    f = np.asarray(filters, np.float32)
    body = struct.pack("<II", f.shape[0], f.shape[1]) + f.astype("<f4").tobytes()
    return struct.pack("<I", len(body)) + body


def write_apr(cfg, tensors, quant: int = Q_F32, filterbank=None, vocab: bytes | None = None) -> bytes:
    """AprWriter::to_bytes (f32) / AprWriterInt8::to_bytes (int8) / the defined Int4 layout.

    tensors: ordered list of (name, np.ndarray f32)."""
    out = bytearray(MAGIC)
    out += header_bytes(cfg, quant, len(tensors), vocab is not None, filterbank is not None)
    blobs, scales, offset = [], [], 0
    for name, arr in tensors:
        arr = np.asarray(arr, np.float32)
        if quant == Q_F32:
            blob = arr.astype("<f4").tobytes()
        elif quant == Q_INT8:
            q, s = quantize_int8(arr)
            blob, _ = q.tobytes(), scales.append(s)
        elif quant == Q_INT4:
            q, s = quantize_int4(arr)
            blob, _ = q.tobytes(), scales.append(s)
        else:
            raise ValueError("unsupported quantization for writing")
        out += desc_bytes(name, arr.shape, offset, len(blob), arr.size)
        blobs.append(blob)
        offset += len(blob)
    for s in scales:
        out += struct.pack("<f", float(s))
    for blob in blobs:
        out += blob
    if vocab is not None:
        out += struct.pack("<I", len(vocab)) + vocab
    if filterbank is not None:
        out += filterbank_section(filterbank)
    out += struct.pack("<I", crc32(bytes(out)))
    return bytes(out)


class AprReader:
    """AprReader::{new,load_tensor,read_mel_filterbank} (format/mod.rs:484-522,610-672,736-780)."""

    def __init__(self, data: bytes):
        if len(data) < 4 or data[:4] != MAGIC:
            raise ValueError("invalid magic")
        self.data = data
        self.header = parse_header(data[4:])
        n = self.header["n_tensors"]
        idx0 = 4 + HEADER_SIZE
        if n > 0 and len(data) < idx0 + n * DESC_SIZE:
            raise ValueError("file too short for tensor index")
        self.tensors = [parse_desc(data[idx0 + i * DESC_SIZE: idx0 + (i + 1) * DESC_SIZE]) for i in range(n)]
        self.scale_table = idx0 + n * DESC_SIZE
        quant = self.header["quantization"]
        self.data_offset = self.scale_table + (4 * n if quant in (Q_INT8, Q_INT4) else 0)

    def names(self):
        return [t["name"] for t in self.tensors]

    def load_tensor(self, name: str) -> np.ndarray:
        for i, t in enumerate(self.tensors):
            if t["name"] == name:
                break
        else:
            raise KeyError(f"tensor not found: {name}")
        quant = self.header["quantization"]
        start = self.data_offset + t["offset"]
        n = t["n_elements"]
        if quant == Q_INT8:
            (scale,) = struct.unpack_from("<f", self.data, self.scale_table + 4 * i)
            if start + n > len(self.data):
                raise ValueError("tensor data out of bounds")
            q = np.frombuffer(self.data, np.int8, n, start)
            return q.astype(np.float32) * np.float32(scale)
        if quant == Q_INT4:
            (scale,) = struct.unpack_from("<f", self.data, self.scale_table + 4 * i)
            nb = (n + 1) // 2
            if start + nb > len(self.data):
                raise ValueError("tensor data out of bounds")
            return dequantize_int4(np.frombuffer(self.data, np.uint8, nb, start), scale, n)
        if start + 4 * n > len(self.data):
            raise ValueError("tensor data out of bounds")
        return np.frombuffer(self.data, "<f4", n, start).astype(np.float32)

    def read_mel_filterbank(self):
        if not self.header["has_filterbank"]:
            return None
        pos = self.data_offset + sum(t["size"] for t in self.tensors)
        if self.header["has_vocab"]:
            if pos + 4 > len(self.data):
                return None
            (vs,) = struct.unpack_from("<I", self.data, pos)
            pos += 4 + vs
        if pos + 4 > len(self.data):
            return None
        (fs,) = struct.unpack_from("<I", self.data, pos)
        body = self.data[pos + 4: pos + 4 + fs]
        if len(body) < fs or fs < 8:
            return None
        n_mels, n_freqs = struct.unpack_from("<II", body, 0)
        if len(body) < 8 + 4 * n_mels * n_freqs:
            return None
        return np.frombuffer(body, "<f4", n_mels * n_freqs, 8).reshape(n_mels, n_freqs).astype(np.float32)

    def load_all(self) -> dict:
        return {t["name"]: self.load_tensor(t["name"]) for t in self.tensors}
