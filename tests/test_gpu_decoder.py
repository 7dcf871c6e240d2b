"""GPU decoder front half (-m gpu): cross-attention K/V precompute, Decoder::forward_one with a KV cache, the token
suppressor and the greedy loop on the device, against the CPU oracle (oracle/decoder.py) on the same random-init decoder.

North-star gate 3: identical greedy tokens on the tiny / base configurations -- now with the GPU decoder on GPU encoder states.
"""
import numpy as np
import pytest

from oracle import decoder as D
from oracle import encoder as E
from oracle import mel as M
from whisper_apr_b200 import WhisperApr, WhisperError, synth

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="module")
def tiny_full():
    cfg = synth.CONFIGS["tiny"]
    data, tensors = synth.random_model_apr(cfg, seed=0, with_decoder=True)
    model = WhisperApr.load_from_apr(data)
    w = dict(tensors)
    yield model, w, E.CONFIGS["tiny"]
    model.close()


def test_decoder_is_loaded_and_encoder_only_files_refuse(tiny_full):
    model = tiny_full[0]
    assert model.has_decoder
    data, _ = synth.random_model_apr(synth.CONFIGS["tiny"], seed=0)              # no decoder tensors in the file
    enc_only = WhisperApr.load_from_apr(data)
    assert not enc_only.has_decoder
    with pytest.raises(WhisperError) as e:
        enc_only.decode_greedy(np.zeros((1500, 384), np.float32), D.initial_tokens(), 8)
    assert e.value.kind == "Model" and "no decoder tensors" in str(e.value)
    enc_only.close()


def test_cross_kv_precompute_matches_oracle(tiny_full):
    """decoder.rs:2276-2296: K = enc . Wk^T (no bias in Whisper checkpoints), V = enc . Wv^T + bv, per decoder layer."""
    model, w, cfg = tiny_full
    rng = np.random.default_rng(5)
    enc = rng.standard_normal((1500, cfg.n_text_state)).astype(np.float32)
    d = cfg.n_text_state
    import torch
    from whisper_apr_b200 import _lib
    t16 = torch.float16 if _lib.lib().wb_operand_format() == b"fp16" else torch.bfloat16
    r = lambda x: torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(t16).to(torch.float64).numpy()
    for layer in (0, cfg.n_text_layer - 1):
        k, v = model.debug_cross_kv(enc, layer)
        p = f"decoder.layers.{layer}.encoder_attn"
        kr = r(enc) @ r(w[p + ".k_proj.weight"]).T
        vr = r(enc) @ r(w[p + ".v_proj.weight"]).T + w[p + ".v_proj.bias"]
        assert np.abs(k - kr).max() <= 2e-2 * max(1.0, np.abs(kr).max() / 4) and _cos(k, kr) > 0.99999
        assert np.abs(v - vr).max() <= 2e-2 * max(1.0, np.abs(vr).max() / 4) and _cos(v, vr) > 0.99999
    # shorter encoder sequences (Encoder::forward_mel of a short mel) go through the same path
    k, v = model.debug_cross_kv(enc[:77], 1)
    assert k.shape == (77, d) and _cos(k, r(enc[:77]) @ r(w["decoder.layers.1.encoder_attn.k_proj.weight"]).T) > 0.99999


@pytest.mark.parametrize("n_tokens", [1, 4, 9])
def test_forward_one_logits_match_oracle(tiny_full, n_tokens):
    """Decoder::forward_one (decoder.rs:2125-2172) after feeding n tokens: logits over the whole vocabulary."""
    model, w, cfg = tiny_full
    rng = np.random.default_rng(7)
    enc = (0.8 * rng.standard_normal((1500, cfg.n_text_state))).astype(np.float32)
    toks = (D.initial_tokens() + [11, 4242, 50000, 7, 300])[:n_tokens]
    got = model.debug_decoder_logits(enc, toks)
    dec = D.Decoder(w, cfg, enc, dtype=np.float64)
    ref = None
    for t in toks:
        ref = dec.forward_one(t)
    assert got.shape == ref.shape == (cfg.n_vocab,)
    # f32 token path; the only reduced-precision inputs are the bf16 cross-attention K/V
    assert np.abs(got - ref).max() <= 2e-2 and _cos(got, ref) > 0.99999
    assert int(np.argmax(got)) == int(np.argmax(ref))


def test_greedy_loop_rules_on_device(tiny_full):
    """greedy.rs:118-146 + processors.rs:126-147 on the device: the sequence starts with the initial tokens, never contains a
    suppressed id, holds at most max_tokens tokens, stops after EOT, is deterministic and batches independently."""
    model, w, cfg = tiny_full
    rng = np.random.default_rng(9)
    enc = (0.8 * rng.standard_normal((3, 1500, cfg.n_text_state))).astype(np.float32)
    init = D.initial_tokens()
    seqs = model.decode_greedy(enc, init, 14)
    sup = set(D.suppressed_ids(cfg.n_vocab).tolist())
    assert len(seqs) == 3
    for s in seqs:
        assert s[:4] == init and len(s) <= 14 and not (set(s[4:]) & sup)
        if D.EOT in s:
            assert s.index(D.EOT) == len(s) - 1
    assert seqs == model.decode_greedy(enc, init, 14)
    assert model.decode_greedy(enc[1], init, 14)[0] == seqs[1]                    # a chunk decodes the same alone and in a batch
    assert model.decode_greedy(enc[:1], init, 4)[0] == init                        # max_tokens == initial length: nothing generated
    ts = model.decode_greedy(enc[:1], init, 10, suppress_timestamps=False)[0]      # with_timestamp_suppression(false): lib.rs:549-551
    assert all(t not in (sup - set(range(D.TIMESTAMP_BASE, cfg.n_vocab))) for t in ts[4:])
    with pytest.raises(WhisperError) as e:
        model.decode_greedy(enc[:1], [D.SOT, 60000], 8)                            # decoder.rs:2141-2146
    assert "out of vocabulary range" in str(e.value)


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_greedy_tokens_identical_gpu_decoder(name, fb80):
    """North-star gate 3 with BOTH halves on the GPU: audio -> mel -> encoder -> cross K/V -> greedy loop, tokens only come back
    (wb_transcribe_tokens_batch), against the oracle's mel + encoder + decoder.  Audio is chosen so that no oracle step is a
    near-tie (top-2 margin >= 0.01): a coin-flip step is not reproducible by any implementation."""
    cfg, ecfg = synth.CONFIGS[name], E.CONFIGS[name]
    data, tensors = synth.random_model_apr(cfg, seed=0, with_decoder=True)
    w = dict(tensors)
    model = WhisperApr.load_from_apr(data)
    chosen = []
    for a in range(7, 30):
        audio = synth.synth_audio(a)
        ref_states = E.forward_mel(M.compute_mel(audio, fb80), w, ecfg, attention=E.naive_attention)
        toks, margins = D.greedy_decode(w, ecfg, ref_states.astype(np.float32), max_tokens=24, return_margins=True)
        if min(margins) >= 0.01:
            chosen.append((audio, ref_states, toks))
        if len(chosen) == 2:
            break
    assert len(chosen) == 2, "no synthetic chunks with a tie-free oracle decode"
    got = model.transcribe_tokens_batch([c[0] for c in chosen], D.initial_tokens(), 24)
    for (audio, ref_states, toks), g in zip(chosen, got):
        assert g == toks
        assert len(set(toks[4:])) >= 3                                              # not a degenerate repeat
        # the two halves separately: GPU decoder on the ORACLE's states, and on the GPU's states
        assert model.decode_greedy(ref_states, D.initial_tokens(), 24)[0] == toks
    model.close()
