"""ctypes wrapper of oracle/whisper_ref.c (the C restatement used as checker and as the timed CPU baseline).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  -march=native: the shared object is rebuilt on the box it runs on
when the prebuilt one fails to load or is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libwhisper_ref.so")
_lib = None
_fp = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    if force and os.path.exists(SO):
        os.remove(SO)
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building oracle/whisper_ref.c failed:\n" + r.stdout + r.stderr)
    return SO


def lib():
    global _lib
    if _lib is None:
        stamp = os.path.join(_HERE, "_ref", "host.txt")
        host = open("/proc/cpuinfo").read().split("flags", 1)[-1].split("\n", 1)[0] if os.path.exists("/proc/cpuinfo") else ""
        if not os.path.exists(SO) or not os.path.exists(stamp) or open(stamp).read() != host:
            build(force=True)            # -march=native code must match the host it runs on
            open(stamp, "w").write(host)
        _lib = C.CDLL(SO)
        _lib.wref_mel_compute.restype = C.c_long
        _lib.wref_conv1d.restype = C.c_long
        _lib.wref_conv_stem.restype = C.c_long
        _lib.wref_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_fp)


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def max_threads() -> int:
    return int(lib().wref_max_threads())


def mel_compute(audio, filters, hop=160):
    audio, filters = _f(audio).ravel(), _f(filters)
    m = filters.shape[0]
    n = audio.size
    nf = (n - 400) // hop + 1 if (n >= 400 and hop > 0) else 0
    out = np.empty((max(nf, 1), m), np.float32)
    got = lib().wref_mel_compute(_p(audio), C.c_long(n), _p(filters), m, C.c_long(hop), _p(out))
    if got < 0:
        raise ValueError("hop_length must be positive")
    return out[:got]


def compute_mel(audio, filters):
    audio, filters = _f(audio).ravel(), _f(filters)
    out = np.empty((3000, filters.shape[0]), np.float32)
    lib().wref_compute_mel(_p(audio), C.c_long(audio.size), _p(filters), filters.shape[0], _p(out))
    return out


def conv_stem(mel, w, cfg):
    mel = _f(mel).reshape(-1, cfg.n_mels)
    d = cfg.n_audio_state
    T = mel.shape[0]
    x = np.empty(((T - 1) // 2 + 1, d), np.float32)
    from . import encoder as E
    pe = _f(E.positional_embedding(w, cfg)[: x.shape[0]])
    S = lib().wref_conv_stem(_p(mel), C.c_long(T), cfg.n_mels, d, _p(_f(w["encoder.conv1.weight"])), _p(_f(w["encoder.conv1.bias"])),
                             _p(_f(w["encoder.conv2.weight"])), _p(_f(w["encoder.conv2.bias"])), _p(pe), _p(x))
    return x[:S]


class LayerWeights:
    """One encoder block's tensors laid out as the reference keeps them after finalize_weights()."""

    def __init__(self, w, i, d):
        p = f"encoder.layers.{i}"
        z = np.zeros(d, np.float32)
        g = lambda n, dflt=None: _f(w[n]) if n in w else dflt
        self.ln1g, self.ln1b = g(f"{p}.self_attn_layer_norm.weight"), g(f"{p}.self_attn_layer_norm.bias")
        self.ln2g, self.ln2b = g(f"{p}.final_layer_norm.weight"), g(f"{p}.final_layer_norm.bias")
        t = lambda n: _f(_f(w[n]).reshape(d, d).T)           # cached transpose (attention.rs:96-105)
        self.wq, self.wk, self.wv, self.wo = (t(f"{p}.self_attn.{k}_proj.weight") for k in ("q", "k", "v", "out"))
        self.bq, self.bk = g(f"{p}.self_attn.q_proj.bias", z), g(f"{p}.self_attn.k_proj.bias", z)
        self.bv, self.bo = g(f"{p}.self_attn.v_proj.bias", z), g(f"{p}.self_attn.out_proj.bias", z)
        self.w1, self.b1 = g(f"{p}.fc1.weight"), g(f"{p}.fc1.bias")
        self.w2, self.b2 = g(f"{p}.fc2.weight"), g(f"{p}.fc2.bias")


def encoder_layer(x, lw: LayerWeights, n_heads, threads=1):
    x = _f(x).copy()
    S, d = x.shape
    lib().wref_encoder_layer(_p(x), C.c_long(S), d, n_heads, _p(lw.ln1g), _p(lw.ln1b), _p(lw.wq), _p(lw.bq), _p(lw.wk), _p(lw.bk),
                             _p(lw.wv), _p(lw.bv), _p(lw.wo), _p(lw.bo), _p(lw.ln2g), _p(lw.ln2b), _p(lw.w1), _p(lw.b1),
                             _p(lw.w2), _p(lw.b2), threads)
    return x


def mha(x, lw: LayerWeights, n_heads, threads=1, h0=0, h1=None):
    x = _f(x)
    S, d = x.shape
    out = np.empty_like(x)
    lib().wref_mha(_p(x), C.c_long(S), d, n_heads, _p(lw.wq), _p(lw.bq), _p(lw.wk), _p(lw.bk), _p(lw.wv), _p(lw.bv), _p(lw.wo),
                   _p(lw.bo), threads, h0, n_heads if h1 is None else h1, _p(out))
    return out


def ffn(x, lw: LayerWeights):
    x = _f(x)
    rows, d = x.shape
    out = np.empty_like(x)
    lib().wref_ffn(_p(x), C.c_long(rows), d, _p(lw.w1), _p(lw.b1), _p(lw.w2), _p(lw.b2), _p(out))
    return out


def layernorm(x, g, b):
    x = _f(x)
    out = np.empty_like(x)
    lib().wref_layernorm(_p(x), C.c_long(x.shape[0]), x.shape[1], _p(_f(g)), _p(_f(b)), _p(out))
    return out


def forward_mel(mel, w, cfg, threads=1):
    """Encoder::forward_mel through the C restatement."""
    x = conv_stem(mel, w, cfg)
    d = cfg.n_audio_state
    for i in range(cfg.n_audio_layer):
        x = encoder_layer(x, LayerWeights(w, i, d), cfg.n_audio_head, threads)
    return layernorm(x, w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"])
