"""Oracle: Whisper decoder + greedy decoding exactly as whisper.apr runs them (CPU, numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The decoder is OUT of the accelerated path; it is
restated here only so that the north star's third parity gate -- "identical greedy tokens on the tiny/base configs" -- can be
checked: the same decoder is fed the oracle's encoder states and the GPU's encoder states and must emit the same tokens.

Follows (paths relative to the reference checkout):
  src/model/decoder.rs:2125-2172  Decoder::forward_one (token + positional embedding, blocks with KV cache, ln_post, logits)
  src/model/decoder.rs:2241-2325  forward_block_cached (pre-norm self-attn over the cache, cross-attn over the encoder states
                                  with K/V computed once, FFN; residual adds)
  src/model/decoder.rs:2414-2459  compute_attention_cached (per head, 64-wide column slices, softmax(q k^T / sqrt(d_head)) v)
  src/model/decoder.rs:1794-1806  project_to_vocab (logits = x . token_embedding^T)
  src/lib.rs:455-481              get_initial_tokens ([SOT, lang, task, no_timestamps] for multilingual vocabularies)
  src/lib.rs:529-598              decode (forward_one per unseen token, WhisperTokenSuppressor, GreedyDecoder)
  src/inference/processors.rs:60-147  WhisperTokenSuppressor::{new, apply}
  src/inference/greedy.rs:83-146  argmax (first maximum wins) and the decode loop (stop after EOT or at max_tokens)
  src/tokenizer/vocab.rs:43-78    special token ids
No decoder weights ship with the reference: `random_decoder_tensors` draws them the way synth.py draws the encoder's.
"""
from __future__ import annotations

import numpy as np

from .encoder import ModelConfig, gelu, layer_norm, linear

EOT, SOT, LANG_BASE, TRANSLATE, TRANSCRIBE = 50257, 50258, 50259, 50358, 50359
SPEAKER_TURN, PREV, NO_SPEECH, NO_TIMESTAMPS, TIMESTAMP_BASE = 50360, 50361, 50362, 50363, 50364


def initial_tokens(language_offset: int = 0, translate: bool = False):
    """lib.rs:455-481 for a multilingual vocabulary (n_vocab 51865): [SOT, lang, task, no_timestamps]."""
    return [SOT, LANG_BASE + language_offset, TRANSLATE if translate else TRANSCRIBE, NO_TIMESTAMPS]


def suppressed_ids(n_vocab: int = 51865, suppress_timestamps: bool = True):
    """processors.rs:60-147: SOT, NO_SPEECH, TRANSLATE, TRANSCRIBE, PREV, SPEAKER_TURN, NO_TIMESTAMPS, every language token
    (LANG_BASE..TRANSLATE), and -- unless word timestamps are requested -- every id from TIMESTAMP_BASE up."""
    ids = [SOT, NO_SPEECH, TRANSLATE, TRANSCRIBE, PREV, SPEAKER_TURN, NO_TIMESTAMPS] + list(range(LANG_BASE, TRANSLATE))
    if suppress_timestamps:
        ids += list(range(TIMESTAMP_BASE, n_vocab))
    return np.array(sorted(set(i for i in ids if i < n_vocab)), np.int64)


def random_decoder_tensors(cfg: ModelConfig, seed: int = 1):
    """Random-init decoder of the named architecture, tensor names as load_decoder_weights reads them (lib.rs:843-929)."""
    rng = np.random.default_rng(seed)
    d, L = cfg.n_text_state, cfg.n_text_layer
    w = {"decoder.embed_tokens.weight": (0.05 * rng.standard_normal((cfg.n_vocab, d))).astype(np.float32),
         # a strong positional term keeps a random decoder from emitting one token forever (the emitted sequence then varies
         # along the positions AND with the audio, which is what the token gate needs)
         "decoder.embed_positions.weight": (0.5 * rng.standard_normal((cfg.n_text_ctx, d))).astype(np.float32)}

    def lin(name, n_out, n_in, bias=True):
        b = 1.0 / np.sqrt(n_in)
        w[name + ".weight"] = rng.uniform(-b, b, (n_out, n_in)).astype(np.float32)
        if bias:
            w[name + ".bias"] = (0.02 * rng.standard_normal(n_out)).astype(np.float32)

    def ln(name):
        w[name + ".weight"] = (1.0 + 0.02 * rng.standard_normal(d)).astype(np.float32)
        w[name + ".bias"] = (0.02 * rng.standard_normal(d)).astype(np.float32)

    for i in range(L):
        p = f"decoder.layers.{i}"
        ln(p + ".self_attn_layer_norm")
        ln(p + ".encoder_attn_layer_norm")
        ln(p + ".final_layer_norm")
        for a in ("self_attn", "encoder_attn"):
            lin(f"{p}.{a}.q_proj", d, d)
            lin(f"{p}.{a}.k_proj", d, d, bias=False)          # k_proj has no bias in Whisper checkpoints
            lin(f"{p}.{a}.v_proj", d, d)
            lin(f"{p}.{a}.out_proj", d, d)
        lin(p + ".fc1", 4 * d, d)
        lin(p + ".fc2", d, 4 * d)
    ln("decoder.layer_norm")
    return w


def _proj(x, w, name, d):
    return linear(x, w[name + ".weight"], w.get(name + ".bias", np.zeros(d, x.dtype)))


def _attend(q, k, v, n_heads):
    """decoder.rs:2414-2459: one query row against kv_len cached rows, per 64-wide head slice."""
    d = q.shape[-1]
    dh = d // n_heads
    out = np.empty(d, q.dtype)
    for h in range(n_heads):
        sl = slice(h * dh, (h + 1) * dh)
        s = (k[:, sl] @ q[sl]) / np.sqrt(dh)
        p = np.exp(s - s.max())
        out[sl] = (p / p.sum()) @ v[:, sl]
    return out


class Decoder:
    """Incremental decoder with the reference's cache layout: per layer the self-attention K/V rows seen so far and the
    cross-attention K/V of the encoder states (computed on the first token, decoder.rs:2276-2296)."""

    def __init__(self, w, cfg: ModelConfig, encoder_states, dtype=np.float32):
        self.w = {k: np.asarray(v, dtype) for k, v in w.items()}
        self.cfg, self.d, self.dtype = cfg, cfg.n_text_state, dtype
        self.enc = np.asarray(encoder_states, dtype).reshape(-1, self.d)
        self.k_self = [np.zeros((0, self.d), dtype) for _ in range(cfg.n_text_layer)]
        self.v_self = [np.zeros((0, self.d), dtype) for _ in range(cfg.n_text_layer)]
        self.kv_cross = [None] * cfg.n_text_layer
        self.pos = 0

    def forward_one(self, token: int) -> np.ndarray:
        w, d, H = self.w, self.d, self.cfg.n_text_head
        if self.pos >= self.cfg.n_text_ctx:
            raise ValueError(f"cache position {self.pos} exceeds max {self.cfg.n_text_ctx}")
        if not 0 <= token < self.cfg.n_vocab:
            raise ValueError(f"token {token} out of vocabulary range {self.cfg.n_vocab}")
        x = w["decoder.embed_tokens.weight"][token] + w["decoder.embed_positions.weight"][self.pos]
        for i in range(self.cfg.n_text_layer):
            p = f"decoder.layers.{i}"
            n = layer_norm(x, w[p + ".self_attn_layer_norm.weight"], w[p + ".self_attn_layer_norm.bias"])
            q = _proj(n, w, p + ".self_attn.q_proj", d)
            self.k_self[i] = np.vstack([self.k_self[i], _proj(n, w, p + ".self_attn.k_proj", d)[None]])
            self.v_self[i] = np.vstack([self.v_self[i], _proj(n, w, p + ".self_attn.v_proj", d)[None]])
            x = x + _proj(_attend(q, self.k_self[i], self.v_self[i], H), w, p + ".self_attn.out_proj", d)
            n = layer_norm(x, w[p + ".encoder_attn_layer_norm.weight"], w[p + ".encoder_attn_layer_norm.bias"])
            if self.kv_cross[i] is None:
                self.kv_cross[i] = (_proj(self.enc, w, p + ".encoder_attn.k_proj", d), _proj(self.enc, w, p + ".encoder_attn.v_proj", d))
            q = _proj(n, w, p + ".encoder_attn.q_proj", d)
            x = x + _proj(_attend(q, self.kv_cross[i][0], self.kv_cross[i][1], H), w, p + ".encoder_attn.out_proj", d)
            n = layer_norm(x, w[p + ".final_layer_norm.weight"], w[p + ".final_layer_norm.bias"])
            x = x + _proj(gelu(_proj(n, w, p + ".fc1", 4 * d)), w, p + ".fc2", d)
        self.pos += 1
        x = layer_norm(x, w["decoder.layer_norm.weight"], w["decoder.layer_norm.bias"])
        return w["decoder.embed_tokens.weight"] @ x


def greedy_decode(w, cfg: ModelConfig, encoder_states, max_tokens: int = 32, init=None, dtype=np.float32, return_margins: bool = False):
    """lib.rs:529-598 + greedy.rs:118-146: feed every unseen token through forward_one, suppress, argmax (first maximum),
    stop after EOT or when the sequence holds max_tokens tokens.  Returns the whole sequence (initial tokens included)."""
    dec = Decoder(w, cfg, encoder_states, dtype)
    tokens = list(initial_tokens() if init is None else init)
    sup = suppressed_ids(cfg.n_vocab)
    margins = []
    logits = None
    done = 0
    while len(tokens) < max_tokens:
        for t in tokens[done:]:
            logits = dec.forward_one(t)
        done = len(tokens)
        logits = logits.copy()
        logits[sup] = -np.inf
        nxt = int(np.argmax(logits))
        if return_margins:
            top2 = np.partition(logits, -2)[-2:]
            margins.append(float(top2[1] - top2[0]))
        tokens.append(nxt)
        if nxt == EOT:
            break
    return (tokens, margins) if return_margins else tokens
