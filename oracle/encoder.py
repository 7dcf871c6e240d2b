"""Oracle: Whisper encoder forward exactly as whisper.apr computes it (CPU, numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to the reference checkout):
  src/model/encoder.rs:72-110    Conv1d::forward (weight [out][in][k], input [t][in])
  src/model/encoder.rs:161-175   ConvFrontend::forward (conv1+GELU, conv2 stride 2 + GELU)
  src/model/encoder.rs:219-251   LayerNorm::forward (population variance, eps 1e-5)
  src/model/encoder.rs:282-295   FeedForward::forward
  src/model/encoder.rs:314-318   gelu (tanh approximation)
  src/model/encoder.rs:346-361   EncoderBlock::forward (pre-norm residual)
  src/model/encoder.rs:429-441   create_positional_embedding (interleaved sin/cos default)
  src/model/encoder.rs:450-478   Encoder::forward
  src/model/encoder.rs:566-660   forward_mel / forward_batch / forward_batch_padded
  src/model/attention.rs:143-167 LinearWeights::forward (y = x W^T + b, W [out][in])
  src/model/attention.rs:267-406 flash_attention (online softmax, KV block 32)
  src/model/attention.rs:894-935 forward_cross_flash (heads = 64-wide column slices)
  src/model/mod.rs:64-150        ModelConfig::{tiny,base,small,medium,large}

`dtype=np.float64` is the truth the tolerances are quoted against; `np.float32`
reproduces what the Rust produces up to summation order inside matmul.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ModelConfig:
    """model/mod.rs:35-150 (encoder-relevant fields + text fields for the .apr header)."""
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

    @property
    def d(self):
        return self.n_audio_state


def _cfg(name, mtype, d, h, L, m=80):
    return ModelConfig(name, mtype, m, 1500, d, h, L, 51865, 448, d, h, L)


CONFIGS = {
    "tiny": _cfg("tiny", 0, 384, 6, 4),
    "base": _cfg("base", 2, 512, 8, 6),
    "small": _cfg("small", 4, 768, 12, 12),
    "medium": _cfg("medium", 6, 1024, 16, 24),
    "large": _cfg("large", 8, 1280, 20, 32),
    # BASELINE.json "whisper-large-v3 shape": large() with 128 mel bins
    "large-v3": _cfg("large-v3", 11, 1280, 20, 32, 128),
}


def gelu(x):
    """encoder.rs:314-318."""
    c = x.dtype.type(0.7978846)
    k = x.dtype.type(0.044715)
    return x.dtype.type(0.5) * x * (x.dtype.type(1.0) + np.tanh(c * (x + k * x * x * x)))


def conv1d(x, weight, bias, stride: int, padding: int = 1):
    """Conv1d::forward (encoder.rs:72-110). x [T][Cin]; weight [Cout][Cin][K]; -> [T_out][Cout]."""
    T, cin = x.shape
    cout, cin2, K = weight.shape
    assert cin == cin2, "Conv1d input size mismatch"
    t_out = (T + 2 * padding - K) // stride + 1
    xp = np.zeros((T + 2 * padding, cin), x.dtype)
    xp[padding:padding + T] = x
    out = np.broadcast_to(bias.astype(x.dtype), (t_out, cout)).copy()
    for k in range(K):
        rows = xp[k:k + stride * (t_out - 1) + 1:stride]
        out += rows @ weight[:, :, k].astype(x.dtype).T
    return out


def conv_out_len(T: int, stride: int, K: int = 3, padding: int = 1) -> int:
    return (T + 2 * padding - K) // stride + 1


def layer_norm(x, gamma, beta, eps=1e-5):
    """LayerNorm::forward (encoder.rs:219-251)."""
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    inv = x.dtype.type(1.0) / np.sqrt(var + x.dtype.type(eps))
    return (x - mean) * inv * gamma.astype(x.dtype) + beta.astype(x.dtype)


def linear(x, weight, bias):
    """LinearWeights::forward (attention.rs:143-167)."""
    return x @ weight.astype(x.dtype).T + bias.astype(x.dtype)


def flash_attention(q, k, v, block_size: int = 32):
    """flash_attention (attention.rs:360-406 with helpers :267-343), mask=None.

    q [S][dh], k,v [KV][dh].  Online softmax over KV blocks; final O/sum (0 if sum<=1e-10)."""
    S, dh = q.shape
    KV = k.shape[0]
    dt = q.dtype
    scale = dt.type(1.0) / np.sqrt(dt.type(dh))
    out = np.zeros((S, dh), dt)
    row_max = np.full(S, -np.inf, dt)
    row_sum = np.zeros(S, dt)
    for s0 in range(0, KV, block_size):
        s1 = min(s0 + block_size, KV)
        scores = (q @ k[s0:s1].T) * scale
        new_max = np.maximum(row_max, scores.max(axis=1))
        scale_prev = np.exp(row_max - new_max)
        row_sum = row_sum * scale_prev
        out = out * scale_prev[:, None]
        p = np.exp(scores - new_max[:, None])
        row_sum = row_sum + p.sum(axis=1)
        out = out + p @ v[s0:s1]
        row_max = new_max
    inv = np.where(row_sum > 1e-10, dt.type(1.0) / np.where(row_sum > 1e-10, row_sum, 1), dt.type(0.0))
    return out * inv[:, None]


def naive_attention(q, k, v):
    """softmax(q k^T / sqrt(dh)) v -- the identity flash_attention must satisfy (attention.rs:1848-1876)."""
    dt = q.dtype
    s = (q @ k.T) * (dt.type(1.0) / np.sqrt(dt.type(q.shape[1])))
    s = s - s.max(axis=1, keepdims=True)
    p = np.exp(s)
    return (p / p.sum(axis=1, keepdims=True)) @ v


def default_positional_embedding(max_len: int, d_model: int) -> np.ndarray:
    """Encoder::create_positional_embedding (encoder.rs:429-441), interleaved sin/cos, f32."""
    pos = np.arange(max_len, dtype=np.float32)[:, None]
    i = np.arange(d_model // 2, dtype=np.float32)[None, :]
    denom = np.power(np.float32(10000.0), np.float32(2.0) * i / np.float32(d_model), dtype=np.float32)
    angle = (pos / denom).astype(np.float32)
    pe = np.zeros((max_len, d_model), np.float32)
    pe[:, 0::2] = np.sin(angle)
    pe[:, 1::2] = np.cos(angle)
    return pe


def _get(w, name, shape, default=0.0):
    """Loader semantics of lib.rs:769-800: a missing tensor keeps its default."""
    if name in w:
        return np.asarray(w[name]).reshape(shape)
    return np.full(shape, default, np.float32)


def mha(x, w, prefix: str, n_heads: int, attention=flash_attention):
    """MultiHeadAttention::forward -> forward_cross_flash (attention.rs:742-782,894-935)."""
    S, d = x.shape
    dh = d // n_heads
    dt = x.dtype
    q = linear(x, _get(w, f"{prefix}.q_proj.weight", (d, d)), _get(w, f"{prefix}.q_proj.bias", (d,)))
    k = linear(x, _get(w, f"{prefix}.k_proj.weight", (d, d)), _get(w, f"{prefix}.k_proj.bias", (d,)))
    v = linear(x, _get(w, f"{prefix}.v_proj.weight", (d, d)), _get(w, f"{prefix}.v_proj.bias", (d,)))
    concat = np.zeros((S, d), dt)
    for h in range(n_heads):
        sl = slice(h * dh, (h + 1) * dh)
        concat[:, sl] = attention(q[:, sl], k[:, sl], v[:, sl])
    return linear(concat, _get(w, f"{prefix}.out_proj.weight", (d, d)), _get(w, f"{prefix}.out_proj.bias", (d,)))


def encoder_block(x, w, i: int, cfg: ModelConfig, attention=flash_attention):
    """EncoderBlock::forward (encoder.rs:346-361)."""
    d = cfg.d
    p = f"encoder.layers.{i}"
    n = layer_norm(x, _get(w, f"{p}.self_attn_layer_norm.weight", (d,), 1.0), _get(w, f"{p}.self_attn_layer_norm.bias", (d,)))
    x = x + mha(n, w, f"{p}.self_attn", cfg.n_audio_head, attention)
    n = layer_norm(x, _get(w, f"{p}.final_layer_norm.weight", (d,), 1.0), _get(w, f"{p}.final_layer_norm.bias", (d,)))
    hid = gelu(linear(n, _get(w, f"{p}.fc1.weight", (4 * d, d)), _get(w, f"{p}.fc1.bias", (4 * d,))))
    return x + linear(hid, _get(w, f"{p}.fc2.weight", (d, 4 * d)), _get(w, f"{p}.fc2.bias", (d,)))


def conv_frontend(mel, w, cfg: ModelConfig, dtype=np.float64):
    """ConvFrontend::forward (encoder.rs:161-175). mel [T][n_mels] -> [S][d]."""
    d, m = cfg.d, cfg.n_mels
    x = np.asarray(mel, dtype=dtype).reshape(-1, m)
    x = gelu(conv1d(x, _get(w, "encoder.conv1.weight", (d, m, 3)), _get(w, "encoder.conv1.bias", (d,)), 1))
    x = gelu(conv1d(x, _get(w, "encoder.conv2.weight", (d, d, 3)), _get(w, "encoder.conv2.bias", (d,)), 2))
    return x


def positional_embedding(w, cfg: ModelConfig):
    """lib.rs:793-800: embed_positions.weight, else positional_embedding, else the default table."""
    for name in ("encoder.embed_positions.weight", "encoder.positional_embedding"):
        if name in w:
            return np.asarray(w[name], np.float32).reshape(cfg.n_audio_ctx, cfg.d)
    return default_positional_embedding(cfg.n_audio_ctx, cfg.d)


def encoder_forward(x, w, cfg: ModelConfig, attention=flash_attention, n_layers=None):
    """Encoder::forward (encoder.rs:450-478). x [S][d] (after the conv stem)."""
    S, d = x.shape
    if d != cfg.d:
        raise ValueError("input size mismatch")
    if S > cfg.n_audio_ctx:
        raise ValueError(f"sequence length {S} exceeds max {cfg.n_audio_ctx}")
    x = x + positional_embedding(w, cfg)[:S].astype(x.dtype)
    L = cfg.n_audio_layer if n_layers is None else n_layers
    for i in range(L):
        x = encoder_block(x, w, i, cfg, attention)
    return layer_norm(x, _get(w, "encoder.layer_norm.weight", (d,), 1.0), _get(w, "encoder.layer_norm.bias", (d,)))


def forward_mel(mel, w, cfg: ModelConfig, dtype=np.float64, attention=flash_attention):
    """Encoder::forward_mel (encoder.rs:566-581)."""
    mel = np.asarray(mel)
    if mel.size % cfg.n_mels != 0:
        raise ValueError(f"mel size {mel.size} not divisible by n_mels {cfg.n_mels}")
    return encoder_forward(conv_frontend(mel, w, cfg, dtype), w, cfg, attention).astype(np.float32)


def forward_batch(mels, w, cfg: ModelConfig, dtype=np.float64):
    """Encoder::forward_batch (encoder.rs:599-608)."""
    return [forward_mel(m, w, cfg, dtype) for m in mels]


def forward_batch_padded(mels, w, cfg: ModelConfig, dtype=np.float64):
    """Encoder::forward_batch_padded (encoder.rs:625-660) -> (features [B][max_S][d], seq_lengths)."""
    enc = forward_batch(mels, w, cfg, dtype)
    lens = [e.shape[0] for e in enc]
    mx = max(lens, default=0)
    out = np.zeros((len(enc), mx, cfg.d), np.float32)
    for b, e in enumerate(enc):
        out[b, : e.shape[0]] = e
    return out, lens
