"""Deterministic synthetic workload: audio chunks and random-init encoder weights.

No datasets or checkpoints are reachable offline, so both the parity tests and bench.py run on
  * audio chunk i = 0.3 sin(2 pi 200 t) + 0.2 sin(2 pi 500 t) + 0.1 sin(2 pi 1000 t) (the reference's own
    "speech-like" test signal, src/audio/mel.rs:1157-1164) with per-chunk frequency jitter and 0.05 N(0,1) noise,
    numpy default_rng(1234 + i);
  * weights from default_rng(seed): linear/conv U(-1/sqrt(fan_in), 1/sqrt(fan_in)), biases N(0, 0.02),
    LayerNorm gamma 1 + N(0, 0.02), beta N(0, 0.02), k_proj.bias omitted (as HF checkpoints do), positional embedding
    = the reference's default table (src/model/encoder.rs:429-441) stored as `encoder.positional_embedding`.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

N_SAMPLES_30S = 480_000
SAMPLE_RATE = 16_000
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass(frozen=True)
class ModelConfig:
    """src/model/mod.rs:35-150."""
    name: str
    model_type: int
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int = 51865
    n_text_ctx: int = 448
    n_text_state: int = 0
    n_text_head: int = 0
    n_text_layer: int = 0


def _cfg(name, mtype, d, h, L, m=80):
    return ModelConfig(name, mtype, m, 1500, d, h, L, 51865, 448, d, h, L)


CONFIGS = {
    "tiny": _cfg("tiny", 0, 384, 6, 4),
    "base": _cfg("base", 2, 512, 8, 6),
    "small": _cfg("small", 4, 768, 12, 12),
    "medium": _cfg("medium", 6, 1024, 16, 24),
    "large": _cfg("large", 8, 1280, 20, 32),
    "large-v3": _cfg("large-v3", 11, 1280, 20, 32, 128),
}


def encoder_gflop_per_chunk(cfg: ModelConfig) -> float:
    """2*3000*3m*d + 2*1500*3d*d + L*(8 S d^2 + 4 S^2 d + 16 S d^2)."""
    d, m, L, S = cfg.n_audio_state, cfg.n_mels, cfg.n_audio_layer, 1500
    return (2 * 3000 * 3 * m * d + 2 * 1500 * 3 * d * d + L * (8 * S * d * d + 4 * S * S * d + 16 * S * d * d)) / 1e9


def load_filterbank(n_mels: int) -> np.ndarray:
    """The slaney filterbank tables the reference ships (data/mel_{80,128}.bin; OpenAI's mel_filters.npz rows), kept as package data."""
    path = os.path.join(_DATA, f"mel_{n_mels}.bin")
    return np.fromfile(path, "<f4").reshape(n_mels, 201).astype(np.float32)


def synth_audio(i: int, n_samples: int = N_SAMPLES_30S) -> np.ndarray:
    rng = np.random.default_rng(1234 + i)
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    j = 1.0 + 0.1 * rng.standard_normal(3)
    x = 0.3 * np.sin(2 * np.pi * 200 * j[0] * t) + 0.2 * np.sin(2 * np.pi * 500 * j[1] * t) + 0.1 * np.sin(2 * np.pi * 1000 * j[2] * t)
    x = x + 0.05 * rng.standard_normal(n_samples)
    return x.astype(np.float32)


def default_positional_embedding(max_len: int, d_model: int) -> np.ndarray:
    pos = np.arange(max_len, dtype=np.float32)[:, None]
    i = np.arange(d_model // 2, dtype=np.float32)[None, :]
    denom = np.power(np.float32(10000.0), np.float32(2.0) * i / np.float32(d_model), dtype=np.float32)
    angle = (pos / denom).astype(np.float32)
    pe = np.zeros((max_len, d_model), np.float32)
    pe[:, 0::2] = np.sin(angle)
    pe[:, 1::2] = np.cos(angle)
    return pe


def random_encoder_tensors(cfg: ModelConfig, seed: int = 0):
    """Ordered (name, f32 array) list with the reference's tensor names (src/lib.rs:769-840, 936-992)."""
    rng = np.random.default_rng(seed)
    d, m = cfg.n_audio_state, cfg.n_mels

    def uni(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return ((rng.random(shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(b)).astype(np.float32)

    def nrm(shape, mean=0.0):
        return (mean + 0.02 * rng.standard_normal(shape, dtype=np.float32)).astype(np.float32)

    out = [
        ("encoder.conv1.weight", uni((d, m, 3), 3 * m)),
        ("encoder.conv1.bias", nrm((d,))),
        ("encoder.conv2.weight", uni((d, d, 3), 3 * d)),
        ("encoder.conv2.bias", nrm((d,))),
        ("encoder.positional_embedding", default_positional_embedding(cfg.n_audio_ctx, d)),
    ]
    for i in range(cfg.n_audio_layer):
        p = f"encoder.layers.{i}"
        out += [
            (f"{p}.self_attn_layer_norm.weight", nrm((d,), 1.0)),
            (f"{p}.self_attn_layer_norm.bias", nrm((d,))),
            (f"{p}.self_attn.q_proj.weight", uni((d, d), d)),
            (f"{p}.self_attn.q_proj.bias", nrm((d,))),
            (f"{p}.self_attn.k_proj.weight", uni((d, d), d)),
            (f"{p}.self_attn.v_proj.weight", uni((d, d), d)),
            (f"{p}.self_attn.v_proj.bias", nrm((d,))),
            (f"{p}.self_attn.out_proj.weight", uni((d, d), d)),
            (f"{p}.self_attn.out_proj.bias", nrm((d,))),
            (f"{p}.final_layer_norm.weight", nrm((d,), 1.0)),
            (f"{p}.final_layer_norm.bias", nrm((d,))),
            (f"{p}.fc1.weight", uni((4 * d, d), d)),
            (f"{p}.fc1.bias", nrm((4 * d,))),
            (f"{p}.fc2.weight", uni((d, 4 * d), 4 * d)),
            (f"{p}.fc2.bias", nrm((d,))),
        ]
    out += [("encoder.layer_norm.weight", nrm((d,), 1.0)), ("encoder.layer_norm.bias", nrm((d,)))]
    return out


def random_decoder_tensors(cfg: ModelConfig, seed: int = 1):
    """Ordered (name, f32 array) list of a random-init decoder, tensor names as load_decoder_weights reads them
    (src/lib.rs:843-929; k_proj has no bias in Whisper checkpoints).  A strong positional term keeps a random decoder from emitting
    one token forever, so the greedy sequence varies along the positions and with the audio."""
    rng = np.random.default_rng(seed)
    d, L = cfg.n_text_state, cfg.n_text_layer
    out = [("decoder.embed_tokens.weight", (0.05 * rng.standard_normal((cfg.n_vocab, d))).astype(np.float32)),
           ("decoder.embed_positions.weight", (0.5 * rng.standard_normal((cfg.n_text_ctx, d))).astype(np.float32))]

    def lin(name, n_out, n_in, bias=True):
        b = 1.0 / np.sqrt(n_in)
        out.append((name + ".weight", rng.uniform(-b, b, (n_out, n_in)).astype(np.float32)))
        if bias:
            out.append((name + ".bias", (0.02 * rng.standard_normal(n_out)).astype(np.float32)))

    def ln(name):
        out.append((name + ".weight", (1.0 + 0.02 * rng.standard_normal(d)).astype(np.float32)))
        out.append((name + ".bias", (0.02 * rng.standard_normal(d)).astype(np.float32)))

    for i in range(L):
        p = f"decoder.layers.{i}"
        ln(p + ".self_attn_layer_norm")
        ln(p + ".encoder_attn_layer_norm")
        ln(p + ".final_layer_norm")
        for a in ("self_attn", "encoder_attn"):
            lin(f"{p}.{a}.q_proj", d, d)
            lin(f"{p}.{a}.k_proj", d, d, bias=False)
            lin(f"{p}.{a}.v_proj", d, d)
            lin(f"{p}.{a}.out_proj", d, d)
        lin(p + ".fc1", 4 * d, d)
        lin(p + ".fc2", d, 4 * d)
    ln("decoder.layer_norm")
    return out


def random_model_apr(cfg: ModelConfig, quant: int = 0, seed: int = 0, with_filterbank: bool = True, with_decoder: bool = False):
    """(apr bytes, tensors list) of a random-init model of the named architecture."""
    from .apr_writer import write_apr
    tensors = random_encoder_tensors(cfg, seed)
    if with_decoder:
        tensors = tensors + random_decoder_tensors(cfg, seed + 1)
    fb = load_filterbank(cfg.n_mels) if with_filterbank else None
    return write_apr(cfg, tensors, quant, fb), tensors
