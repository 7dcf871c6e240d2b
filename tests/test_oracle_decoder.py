"""Oracle decoder (test infrastructure for the greedy-token gate): structural pins against the reference's own rules."""
import numpy as np

from oracle import decoder as D
from oracle import encoder as E


def _small_cfg():
    return E.ModelConfig("toy", 0, 80, 1500, 128, 2, 2, n_vocab=51865, n_text_ctx=32, n_text_state=128, n_text_head=2, n_text_layer=2)


def test_special_tokens_and_initial_sequence():
    # src/tokenizer/vocab.rs:43-78, src/lib.rs:455-481
    assert (D.EOT, D.SOT, D.LANG_BASE, D.TRANSLATE, D.TRANSCRIBE, D.NO_TIMESTAMPS, D.TIMESTAMP_BASE) == (50257, 50258, 50259, 50358, 50359, 50363, 50364)
    assert D.initial_tokens() == [50258, 50259, 50359, 50363]
    assert D.initial_tokens(language_offset=3, translate=True) == [50258, 50262, 50358, 50363]


def test_suppression_list_matches_whisper_token_suppressor():
    # src/inference/processors.rs:60-147: 7 specials + 99 language tokens + all timestamps; EOT and text tokens stay selectable
    ids = set(D.suppressed_ids(51865).tolist())
    assert {50258, 50362, 50358, 50359, 50361, 50360, 50363} <= ids
    assert set(range(50259, 50358)) <= ids and set(range(50364, 51865)) <= ids
    assert 50257 not in ids and 0 not in ids and 50256 not in ids
    assert len(ids) == 7 + 99 + (51865 - 50364)
    assert 50364 not in set(D.suppressed_ids(51865, suppress_timestamps=False).tolist())


def test_cached_decoder_equals_full_recomputation():
    """forward_one with the KV cache (decoder.rs:2241-2325) == attending over all previous positions from scratch."""
    cfg = _small_cfg()
    w = D.random_decoder_tensors(cfg, seed=3)
    rng = np.random.default_rng(0)
    enc = rng.standard_normal((50, cfg.n_text_state)).astype(np.float32)
    toks = [50258, 50259, 50359, 50363, 11, 4242]
    dec = D.Decoder(w, cfg, enc, dtype=np.float64)
    last = None
    for t in toks:
        last = dec.forward_one(t)
    # independent restatement: full causal self-attention over the whole prefix, last row only
    d, H = cfg.n_text_state, cfg.n_text_head
    W = {k: np.asarray(v, np.float64) for k, v in w.items()}
    x = W["decoder.embed_tokens.weight"][toks] + W["decoder.embed_positions.weight"][:len(toks)]
    for i in range(cfg.n_text_layer):
        p = f"decoder.layers.{i}"
        n = E.layer_norm(x, W[p + ".self_attn_layer_norm.weight"], W[p + ".self_attn_layer_norm.bias"])
        q, k, v = (D._proj(n, W, f"{p}.self_attn.{a}_proj", d) for a in "qkv")
        att = np.stack([D._attend(q[r], k[:r + 1], v[:r + 1], H) for r in range(len(toks))])
        x = x + D._proj(att, W, p + ".self_attn.out_proj", d)
        n = E.layer_norm(x, W[p + ".encoder_attn_layer_norm.weight"], W[p + ".encoder_attn_layer_norm.bias"])
        q = D._proj(n, W, p + ".encoder_attn.q_proj", d)
        ke, ve = D._proj(enc.astype(np.float64), W, p + ".encoder_attn.k_proj", d), D._proj(enc.astype(np.float64), W, p + ".encoder_attn.v_proj", d)
        x = x + D._proj(np.stack([D._attend(q[r], ke, ve, H) for r in range(len(toks))]), W, p + ".encoder_attn.out_proj", d)
        n = E.layer_norm(x, W[p + ".final_layer_norm.weight"], W[p + ".final_layer_norm.bias"])
        x = x + D._proj(E.gelu(D._proj(n, W, p + ".fc1", 4 * d)), W, p + ".fc2", d)
    ref = W["decoder.embed_tokens.weight"] @ E.layer_norm(x[-1], W["decoder.layer_norm.weight"], W["decoder.layer_norm.bias"])
    assert np.abs(last - ref).max() < 1e-9


def test_greedy_loop_rules():
    """greedy.rs:118-146: sequence starts with the initial tokens, never emits a suppressed id, stops at max_tokens or after EOT."""
    cfg = _small_cfg()
    w = D.random_decoder_tensors(cfg, seed=4)
    enc = np.random.default_rng(1).standard_normal((40, cfg.n_text_state)).astype(np.float32)
    toks, margins = D.greedy_decode(w, cfg, enc, max_tokens=12, return_margins=True)
    assert toks[:4] == D.initial_tokens() and len(toks) <= 12 and len(margins) == len(toks) - 4
    sup = set(D.suppressed_ids(cfg.n_vocab).tolist())
    assert not (set(toks[4:]) & sup)
    if D.EOT in toks:
        assert toks.index(D.EOT) == len(toks) - 1
    assert toks == D.greedy_decode(w, cfg, enc, max_tokens=12)          # deterministic
    # different encoder states steer the decoder (cross-attention is live)
    other = D.greedy_decode(w, cfg, 3.0 * enc[::-1].copy(), max_tokens=12)
    assert other[:4] == toks[:4]
