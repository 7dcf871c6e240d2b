"""Pins the `.apr` oracle (and the product's Python writer) against the reference's format known-answer tests (CPU only).

KATs: src/format/checksum.rs:140-185 (CRC-32 check values), src/format/mod.rs:1625-1637 (header round trip),
:1774-1807 (descriptor byte layout), :495-501,1326-1329 (Int8 scale table between index and data),
src/model/quantized.rs:2596-2660 (int8 quantiser), :1887-1969 (int4 packing).
"""
import struct

import numpy as np
import pytest

from oracle import apr_format as F
from oracle import encoder as E
from whisper_apr_b200 import apr_writer, synth


@pytest.mark.parametrize("data,crc", [(b"", 0x00000000), (b"Hello, World!", 0xEC4AC3D0), (b"123456789", 0xCBF43926),
                                      (b"a", 0xE8B7BE43), (b"\x00", 0xD202EF8D), (b"\x00" * 32, 0x190A55AD),
                                      (b"\xff" * 32, 0xFF6CAB0B)])
def test_crc32_check_values(data, crc):
    assert F.crc32(data) == crc


def test_header_roundtrip_and_layout():
    cfg = E.CONFIGS["tiny"]
    b = F.header_bytes(cfg, F.Q_INT8, 67, has_vocab=False, has_filterbank=True)
    assert len(b) == 48
    assert b[0:2] == b"\x01\x00" and b[3] == 2 and b[4] == 0 and b[5:7] == struct.pack("<H", 67) and b[7] == 0b10
    h = F.parse_header(b)
    assert (h["n_vocab"], h["n_audio_ctx"], h["n_audio_state"], h["n_audio_head"], h["n_audio_layer"], h["n_mels"]) == \
        (51865, 1500, 384, 6, 4, 80)
    assert h["quantization"] == 2 and h["has_filterbank"] and not h["has_vocab"] and h["n_tensors"] == 67
    with pytest.raises(ValueError):
        F.parse_header(b[:10])                                   # "header too short"
    bad = bytearray(b)
    bad[0] = 9
    with pytest.raises(ValueError):
        F.parse_header(bytes(bad))                               # unsupported version


def test_descriptor_layout_kat():
    raw = bytearray(96)
    raw[:10] = b"encoder.pe"
    struct.pack_into("<QQQ", raw, 48, 1000, 2000, 500)
    raw[88] = 2
    d = F.parse_desc(bytes(raw))
    assert (d["name"], d["offset"], d["size"], d["n_elements"], d["n_dims"]) == ("encoder.pe", 1000, 2000, 500, 2)
    b = F.desc_bytes("test.weight", (384, 512), 0, 768000)
    assert len(b) == 96 and b.startswith(b"test.weight\0")
    assert F.parse_desc(b)["n_elements"] == 384 * 512 and F.parse_desc(b)["shape"] == (384, 512)
    long_name = "x" * 60
    assert F.parse_desc(F.desc_bytes(long_name, (1,), 0, 4))["name"] == "x" * 47     # 47 chars + NUL
    with pytest.raises(ValueError):
        F.parse_desc(bytes(32))


def test_int8_quantiser_kats():
    q, s = F.quantize_int8(np.zeros(10, np.float32))
    assert (q == 0).all() and s == 1.0
    q, s = F.quantize_int8(np.array([127.0], np.float32))
    assert q[0] == 127 and s == 1.0
    q, s = F.quantize_int8(np.array([-127.0], np.float32))
    assert q[0] == -127 and s == 1.0
    q, s = F.quantize_int8(np.array([0.1, -0.1, 0.05, -0.05], np.float32))
    assert abs(s - 0.1 / 127.0) < 1e-6 and list(q) == [127, -127, 64, -64]          # round half away from zero
    x = np.array([1.0, -1.0, 0.5, -0.5, 0.0], np.float32)
    q, s = F.quantize_int8(x)
    assert np.abs(q.astype(np.float32) * s - x).max() < 0.02


def test_int4_packing_kats():
    x = np.array([7.0, -7.0, 1.0, -8.0, 3.0], np.float32)       # absmax 8 -> scale 8/7
    p, s = F.quantize_int4(x)
    assert abs(s - 8.0 / 7.0) < 1e-6 and p.size == 3
    q = np.clip(np.round(x / s), -8, 7)
    assert p[0] & 0x0F == int(q[0]) & 0x0F and p[0] >> 4 == int(q[1]) & 0x0F       # even -> low nibble, odd -> high
    back = F.dequantize_int4(p, s, 5)
    assert np.abs(back - x).max() <= s / 2 + 1e-6
    p0, s0 = F.quantize_int4(np.zeros(4, np.float32))
    assert s0 == 1.0 and (p0 == 0).all()


@pytest.mark.parametrize("quant", [F.Q_F32, F.Q_INT8, F.Q_INT4])
def test_apr_roundtrip_and_writer_cross_check(quant, fb80):
    ocfg, pcfg = E.CONFIGS["tiny"], synth.CONFIGS["tiny"]
    tensors = synth.random_encoder_tensors(pcfg, seed=3)[:9]
    data = F.write_apr(ocfg, tensors, quant, fb80)
    assert data == apr_writer.write_apr(pcfg, tensors, quant, fb80)                 # two independent writers agree
    assert struct.unpack("<I", data[-4:])[0] == F.crc32(data[:-4])                   # format/mod.rs:1149-1151
    r = F.AprReader(data)
    assert r.header["n_tensors"] == 9 and r.names() == [n for n, _ in tensors]
    n = len(tensors)
    assert r.data_offset == 4 + 48 + 96 * n + (4 * n if quant else 0)                # scale table sits before the data
    assert np.array_equal(r.read_mel_filterbank(), fb80)
    for name, arr in tensors:
        got = r.load_tensor(name)
        assert got.size == arr.size
        if quant == F.Q_F32:
            assert np.array_equal(got, arr.ravel())
        else:
            levels = 127.0 if quant == F.Q_INT8 else 7.0
            assert np.abs(got - arr.ravel()).max() <= np.abs(arr).max() / levels * 0.5 + 1e-7
    with pytest.raises(KeyError):
        r.load_tensor("nope")


def test_reader_errors():
    with pytest.raises(ValueError):
        F.AprReader(b"XXXX" + bytes(100))                        # invalid magic
    data = F.write_apr(E.CONFIGS["tiny"], synth.random_encoder_tensors(synth.CONFIGS["tiny"])[:2], 0, None)
    with pytest.raises(ValueError):
        F.AprReader(data[:60])                                   # file too short for tensor index
    r = F.AprReader(data[:4 + 48 + 2 * 96 + 100])
    with pytest.raises(ValueError):
        r.load_tensor("encoder.conv1.weight")                    # tensor data out of bounds
    assert r.read_mel_filterbank() is None
