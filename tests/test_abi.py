"""CPU-side checks of the C ABI: the library loads without a GPU, exports every symbol include/whisper_b200.h declares, its
host-only entry points agree with the oracle, and compute calls fail loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import mel as M
from whisper_apr_b200 import WhisperApr, WhisperError, _lib, split_into_chunks, synth, to_padded_tensor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/whisper_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert b"sm_100a" in lib.wb_version()


def test_fft_factorisation_matches_dft():
    lib = _lib.lib()
    rng = np.random.default_rng(0)
    for trial in range(8):
        y = (rng.standard_normal(400) * 10 ** rng.uniform(-3, 1)).astype(np.float32)
        if trial == 0:
            y[:] = 0
            y[7] = 1.0
        p = np.zeros(201, np.float32)
        lib.wb_debug_fft400_power_host(y.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p))
        ref = np.abs(np.fft.rfft(y.astype(np.float64))) ** 2
        assert np.abs(p - ref).max() <= 2e-6 * max(ref.max(), 1e-30)


@pytest.mark.parametrize("n,chunk,overlap", [(10, 4, 1), (10, 4, 0), (10, 3, 5), (0, 4, 1), (10, 0, 0), (160000, 80000, 8000), (7, 10, 2)])
def test_split_into_chunks_matches_oracle(n, chunk, overlap):
    s = np.arange(n, dtype=np.float32)
    got, ref = split_into_chunks(s, chunk, overlap), M.split_into_chunks(s, chunk, overlap)
    assert len(got) == len(ref) and all(np.array_equal(a, b) for a, b in zip(got, ref))


def test_to_padded_tensor_matches_oracle():
    rng = np.random.default_rng(1)
    mels = [rng.standard_normal((f, 80)).astype(np.float32) for f in (5, 9, 1)]
    assert np.array_equal(to_padded_tensor(mels, 80), M.to_padded_tensor(mels, 80))
    assert to_padded_tensor([], 80).shape == (0, 80, 0)


def test_format_errors_need_no_gpu():
    for bad, msg in [(b"XXXX" + bytes(64), "invalid magic"), (b"APR1" + bytes(10), "header too short")]:
        with pytest.raises(WhisperError) as e:
            WhisperApr.load_from_apr(bad)
        assert e.value.kind == "Format" and msg in str(e.value)
    data, _ = synth.random_model_apr(synth.CONFIGS["tiny"])
    hdr = bytearray(data[:200])
    hdr[4] = 9                                      # version 9
    with pytest.raises(WhisperError) as e:
        WhisperApr.load_from_apr(bytes(hdr))
    assert "unsupported format version" in str(e.value)
    with pytest.raises(WhisperError) as e:
        WhisperApr.load_from_apr(data[:4 + 48 + 96])   # index truncated
    assert "file too short for tensor index" in str(e.value)


def test_no_cpu_fallback():
    if _lib.lib().wb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    data, _ = synth.random_model_apr(synth.CONFIGS["tiny"])
    with pytest.raises(WhisperError) as e:
        WhisperApr.load_from_apr(data)
    assert e.value.kind == "Cuda" and "no CPU fallback" in str(e.value)
    out = np.zeros((4, 4), np.float32)
    st = _lib.lib().wb_debug_layernorm(0, out.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 4, 4,
                                       out.ctypes.data_as(C.c_void_p))
    assert st == _lib.WB_ERR_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "whisper_apr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "oracle/" not in src or f.endswith(".md"), f"{f} references oracle/"


def test_hostile_tensor_descriptor_is_out_of_bounds():
    """A descriptor whose n_elements / offset would wrap 64-bit arithmetic must read as "tensor data out of bounds"
    (format/mod.rs:610-628), for every payload type; a sane one reports its byte count."""
    import struct
    lib = _lib.lib()
    cfg = synth.ModelConfig("t", 0, 80, 1500, 128, 2, 1, 51865, 448, 128, 2, 1)
    for quant, per in ((0, 4.0), (2, 1.0), (3, 0.5)):
        data = bytearray(synth.random_model_apr(cfg, quant=quant)[0])
        name = b"encoder.conv1.bias"
        buf = (C.c_ubyte * len(data)).from_buffer(data)
        assert lib.wb_debug_apr_tensor_bytes(buf, len(data), name) == int(128 * per)
        assert lib.wb_debug_apr_tensor_bytes(buf, len(data), b"no.such.tensor") == -1
        # descriptor 1 is conv1.bias: offset @ +48, n_elements @ +64
        base = 4 + 48 + 96 * 1
        assert bytes(data[base:base + len(name)]) == name
        for n_elem in (1 << 62, (1 << 64) - 1, 1 << 63, len(data) * 8):
            bad = bytearray(data)
            struct.pack_into("<Q", bad, base + 64, n_elem)
            b2 = (C.c_ubyte * len(bad)).from_buffer(bad)
            assert lib.wb_debug_apr_tensor_bytes(b2, len(bad), name) == -1, (quant, n_elem)
        for off in ((1 << 64) - 8, (1 << 63), len(data)):
            bad = bytearray(data)
            struct.pack_into("<Q", bad, base + 48, off)
            b2 = (C.c_ubyte * len(bad)).from_buffer(bad)
            assert lib.wb_debug_apr_tensor_bytes(b2, len(bad), name) == -1, (quant, off)
    assert lib.wb_debug_apr_tensor_bytes(None, 0, b"x") == -2


def test_unreasonable_header_dimensions_are_refused():
    import struct
    data = bytearray(synth.random_model_apr(synth.ModelConfig("t", 0, 80, 1500, 128, 2, 1, 51865, 448, 128, 2, 1))[0])
    struct.pack_into("<I", data, 4 + 16, 1 << 30)          # n_audio_state
    with pytest.raises(WhisperError) as e:
        WhisperApr.load_from_apr(bytes(data))
    assert e.value.kind == "Format" and "unreasonable model dimensions" in str(e.value)


def test_caller_supplied_out_buffer_is_validated():
    from whisper_apr_b200 import api
    api._check_out(np.empty((2, 1500, 384), np.float32), (2, 1500, 384), np.float32)
    for bad in (np.empty((2, 1500, 384), np.float64), np.empty((2, 1500, 383), np.float32),
                np.empty((2, 1500, 768), np.float32)[:, :, ::2], [1, 2, 3]):
        with pytest.raises((ValueError, TypeError)):
            api._check_out(bad, (2, 1500, 384), np.float32)
    ro = np.empty((2, 1500, 384), np.float32)
    ro.setflags(write=False)
    with pytest.raises(ValueError):
        api._check_out(ro, (2, 1500, 384), np.float32)
