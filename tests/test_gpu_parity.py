"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Gates (BASELINE.json north_star): mel <= 1e-4 max-abs (fp32); encoder hidden states <= 2e-2 max-abs and cosine >= 0.9999
(bf16 tensor-core math vs the float64 oracle).  Edge cases follow the reference's own tests.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import apr_format as F
from oracle import cref
from oracle import encoder as E
from oracle import mel as M
from whisper_apr_b200 import WhisperApr, WhisperError, _lib, bf16_bits_to_f32, synth

pytestmark = pytest.mark.gpu

MEL_TOL = 1e-4
ENC_TOL, ENC_COS = 2e-2, 0.9999


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bf16_round(x):
    """Round to the library's 16-bit OPERAND format (fp16 by default, bf16 in a -DWB_OPERANDS_BF16 build)."""
    import torch
    t = torch.float16 if _lib.lib().wb_operand_format() == b"fp16" else torch.bfloat16
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(t).to(torch.float32).numpy()


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="module")
def tiny():
    cfg = synth.CONFIGS["tiny"]
    data, tensors = synth.random_model_apr(cfg, seed=0)
    model = WhisperApr.load_from_apr(data)
    yield model, dict(tensors), E.CONFIGS["tiny"]
    model.close()


@pytest.fixture(scope="module")
def mel0(fb80):
    return M.compute_mel(synth.synth_audio(0), fb80)


# --------------------------------------------------------------------------------------------- mel
def test_mel_golden_vector(tiny, golden_audio, golden_mel, fb80):
    model = tiny[0]
    got = model.mel_filters.compute(golden_audio, 160)
    assert got.shape == (148, 80)
    assert np.abs(got - M.mel_compute(golden_audio, fb80, 160)).max() <= MEL_TOL
    # the in-tree golden (symmetric-window variant) is matched as loosely as the reference's own algorithm matches it
    assert np.abs(got - golden_mel).max() < 0.02 and _cos(got, golden_mel) > 0.99999
    # the C restatement (reference loop structure, f32) agrees too
    assert np.abs(got - cref.mel_compute(golden_audio, fb80)).max() <= MEL_TOL


@pytest.mark.parametrize("n,frames", [(0, 0), (100, 0), (399, 0), (400, 1), (16000, 98), (24000, 148)])
def test_mel_frame_counts(tiny, n, frames):
    out = tiny[0].mel_filters.compute(np.full(n, 0.25, np.float32), 160)     # mel.rs:633-668
    assert out.shape == (frames, 80)


def test_mel_hop_zero_is_audio_error(tiny):
    with pytest.raises(WhisperError) as e:
        tiny[0].mel_filters.compute(np.ones(1000, np.float32), 0)            # mel.rs:690-698
    assert e.value.kind == "Audio" and "hop_length must be positive" in str(e.value)


@pytest.mark.parametrize("hop", [160, 100, 517, 1, 4000])
def test_mel_any_hop(tiny, fb80, hop):
    a = synth.synth_audio(3)[:20000]
    got = tiny[0].mel_filters.compute(a, hop)
    ref = M.mel_compute(a, fb80, hop)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= MEL_TOL


def test_mel_silence_and_determinism(tiny):
    model = tiny[0]
    out = model.mel_filters.compute(np.zeros(16000, np.float32), 160)
    assert (out < 0).all() and np.allclose(out, -1.5)                          # mel.rs:786-799
    a = synth.synth_audio(5)[:48000]
    assert np.array_equal(model.mel_filters.compute(a), model.mel_filters.compute(a))   # mel.rs:1179-1196


@pytest.mark.parametrize("n", [480000, 80000, 500000, 0, 399])
def test_compute_mel_padding_rules(tiny, fb80, n):
    a = np.resize(synth.synth_audio(1), n).astype(np.float32) if n else np.zeros(0, np.float32)
    got = tiny[0].compute_mel(a)
    ref = M.compute_mel(a, fb80)
    assert got.shape == (3000, 80) and (got[2998:] == -1.0).all()             # lib.rs:431-437
    assert np.abs(got - ref).max() <= MEL_TOL


def test_compute_mel_batch_and_128_mels(fb128):
    cfg = synth.ModelConfig("t128", 0, 128, 1500, 384, 6, 1, 51865, 448, 384, 6, 1)
    data, _ = synth.random_model_apr(cfg, seed=1)
    model = WhisperApr.load_from_apr(data)
    audio = np.stack([synth.synth_audio(i) for i in range(3)])
    got = model.compute_mel_batch(audio)
    assert got.shape == (3, 3000, 128)
    for i in range(3):
        assert np.abs(got[i] - M.compute_mel(audio[i], fb128)).max() <= MEL_TOL
    model.close()


def test_mel_htk_fallback_filterbank():
    cfg = synth.CONFIGS["tiny"]
    data, _ = synth.random_model_apr(cfg, seed=0, with_filterbank=False)        # lib.rs:297-298 -> MelFilterbank::new
    model = WhisperApr.load_from_apr(data)
    t = np.arange(16000) / 16000.0
    for f, lo, hi in [(440.0, 10, 35), (4000.0, 40, 79)]:                       # mel.rs:1037-1118
        tone = np.sin(2 * np.pi * f * t).astype(np.float32)
        got = model.mel_filters.compute(tone)
        assert np.abs(got - M.mel_compute(tone, M.htk_filterbank(80))).max() <= MEL_TOL
        assert lo <= got.mean(axis=0).argmax() <= hi
    model.close()


# --------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("M_,N,K", [(128, 128, 64), (1500, 384, 384), (300, 1152, 384), (777, 1280, 1280), (200, 512, 2048), (130, 256, 240)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_gemm_epilogues(M_, N, K, epi):
    rng = np.random.default_rng(M_ + N + K + epi)
    A = rng.standard_normal((M_, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    extra = rng.standard_normal((M_, N)).astype(np.float32)
    out = np.empty((M_, N), np.float32)
    _lib.check(_lib.lib().wb_debug_gemm(0, _p(A), _p(W), _p(bias), _p(extra), M_, N, K, epi, C.c_float(0.5), _p(out)))
    ref = 0.5 * (_bf16_round(A).astype(np.float64) @ _bf16_round(W).astype(np.float64).T) + bias
    if epi == 1:
        ref = E.gelu(ref)
    elif epi == 2:
        ref = extra + ref
    elif epi == 3:
        ref = E.gelu(ref) + extra
    tol = 2e-5 if epi in (2, 3, 4) else 2.5e-2          # f32 outputs: accumulation order only; bf16 outputs: one rounding
    assert np.abs(out - ref).max() <= tol * max(1.0, np.abs(ref).max() / 4)
    assert _cos(out, ref) > 0.99999


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (1, 77, 1), (2, 1500, 6), (1, 1499, 2), (1, 129, 3)])
def test_attention_vs_naive_softmax(B, S, H):
    rng = np.random.default_rng(B * 1000 + S + H)
    d = 64 * H
    qkv = (rng.standard_normal((B, S, 3 * d)) * 1.5).astype(np.float32)
    out = np.empty((B, S, d), np.float32)
    _lib.check(_lib.lib().wb_debug_attention(0, _p(qkv), B, S, d, H, _p(out)))
    r = _bf16_round(qkv).astype(np.float64)
    for b in range(B):
        for h in range(H):
            q, k, v = (r[b, :, o + h * 64: o + (h + 1) * 64] for o in (0, d, 2 * d))
            ref = E.naive_attention(q, k, v)
            got = out[b, :, h * 64:(h + 1) * 64]
            assert np.abs(got - ref).max() <= 2.5e-2 and _cos(got, ref) > 0.99999


@pytest.mark.parametrize("rows,d", [(5, 384), (1500, 1280), (33, 512), (7, 200), (9, 2048)])
def test_layernorm(rows, d):
    rng = np.random.default_rng(rows + d)
    x = (rng.standard_normal((rows, d)) * 3 + 1).astype(np.float32)
    g = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    b = (0.1 * rng.standard_normal(d)).astype(np.float32)
    out = np.empty_like(x)
    _lib.check(_lib.lib().wb_debug_layernorm(0, _p(x), _p(g), _p(b), rows, d, _p(out)))
    assert np.abs(out - E.layer_norm(x.astype(np.float64), g, b)).max() <= 5e-6


# --------------------------------------------------------------------------------------------- encoder
def test_encoder_stagewise_tiny(tiny, mel0):
    model, w, cfg = tiny
    x = E.conv_frontend(mel0, w, cfg) + E.positional_embedding(w, cfg)[:1500]
    got = model.debug_encode(mel0, n_layers=0, ln_post=False)
    assert np.abs(got - x).max() <= 5e-3 and _cos(got, x) > 0.99999           # conv stem + positional embedding
    for i in range(cfg.n_audio_layer):
        x = E.encoder_block(x, w, i, cfg, attention=E.naive_attention)
        got = model.debug_encode(mel0, n_layers=i + 1, ln_post=False)
        assert np.abs(got - x).max() <= ENC_TOL and _cos(got, x) >= ENC_COS
    ref = E.layer_norm(x, w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"])
    got = model.encode(mel0)
    assert got.shape == (1500, 384)
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS
    # reference's own statistical gate (tests/ground_truth_tests.rs:733-742)
    assert abs(got.mean()) < 0.5 and 0.5 < got.std() < 3.0 and np.isfinite(got).all()


@pytest.mark.parametrize("T", [3000, 1001, 70])
def test_layernorm_follower_bit_identical(tiny, mel0, T):
    """The experimental concurrent LayerNorm (a follower kernel fed by the residual GEMM's completion counters, off by default)
    must give bit-identical encoder states to the separate launches: same arithmetic, different schedule -- whichever of the
    follower and the sweep behind the GEMM normalises a row group."""
    import whisper_apr_b200
    model, w, cfg = tiny
    L = whisper_apr_b200.lib()
    ref = model.encode(mel0[:T])
    try:
        whisper_apr_b200._lib.check(L.wb_debug_set_ln_follow(model._h, 1))
        for _ in range(3):                                                     # eager, captured and replayed
            got = model.encode(mel0[:T])
            assert np.array_equal(got, ref)
    finally:
        whisper_apr_b200._lib.check(L.wb_debug_set_ln_follow(model._h, 0))


def test_encoder_matches_c_restatement(tiny, mel0):
    model, w, cfg = tiny
    ref = cref.forward_mel(mel0[:600], w, cfg, threads=4)                      # f32, reference loop structure, flash block 32
    got = model.encode(mel0[:600])
    assert got.shape == ref.shape == (300, 384)
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS


@pytest.mark.parametrize("T", [3000, 2999, 1001, 256, 2, 1])
def test_encoder_variable_length(tiny, mel0, T):
    model, w, cfg = tiny
    got = model.encode(mel0[:T])
    ref = E.forward_mel(mel0[:T], w, cfg, attention=E.naive_attention)
    assert got.shape == ref.shape == ((T - 1) // 2 + 1, 384)                  # encoder.rs:1172-1178
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS


def test_encoder_errors(tiny):
    model = tiny[0]
    with pytest.raises(WhisperError) as e:
        model.encode(np.zeros(81, np.float32))                                  # encoder.rs:568-574
    assert e.value.kind == "Model" and "not divisible by n_mels" in str(e.value)
    with pytest.raises(WhisperError) as e:
        model.encode(np.zeros((3002, 80), np.float32))                          # 1501 positions, encoder.rs:456-461
    assert e.value.kind == "Model" and "exceeds max 1500" in str(e.value)
    assert model.encode(np.zeros((0, 80), np.float32)).shape == (0, 384)


def test_forward_batch_padded(tiny, mel0):
    model, w, cfg = tiny
    mels = [mel0[:40], mel0[:21], mel0[:40], mel0[:3000]]
    o = model.encoder.forward_batch_padded(mels)                                # encoder.rs:1269-1377
    assert o.seq_lengths == [20, 11, 20, 1500] and o.max_seq_len == 1500 and o.batch_size == 4 and o.total_tokens() == 1551
    assert (o.features[1, 11:] == 0).all() and (o.features[0, 20:] == 0).all()
    for i, mm in enumerate(mels[:3]):
        ref = E.forward_mel(mm, w, cfg, attention=E.naive_attention)
        assert np.abs(o.get(i) - ref).max() <= ENC_TOL
    assert o.get(7) is None
    assert np.array_equal(o.get(0), o.get(2))                                   # batching does not change results
    assert model.encoder.forward_batch_padded([]).is_empty()


def test_fused_batch_entry_point(tiny, fb80):
    model, w, cfg = tiny
    audio = [synth.synth_audio(10), synth.synth_audio(11)[:80000], synth.synth_audio(12)[:160000]]
    out = model.mel_encode_batch(audio)                                         # lib.rs:1162-1170
    assert out.shape == (3, 1500, 384)
    for i, a in enumerate(audio):
        ref = E.forward_mel(M.compute_mel(a, fb80), w, cfg, attention=E.naive_attention)
        assert np.abs(out[i] - ref).max() <= ENC_TOL and _cos(out[i], ref) >= ENC_COS
    single = model.encode(model.compute_mel(audio[1]))
    assert np.abs(single - out[1]).max() <= 5e-3        # bf16 mel hand-off (fused) vs f32 mel round trip: one extra rounding
    bf = bf16_bits_to_f32(model.mel_encode_batch(audio[:1], out_dtype="bf16"))
    assert np.abs(bf[0] - out[0]).max() <= 2e-2
    model.set_max_batch(2)                               # micro-batching must not change results
    out2 = model.mel_encode_batch(audio)
    model.set_max_batch(32)
    assert np.array_equal(out, out2)


@pytest.mark.parametrize("quant,levels", [(F.Q_INT8, 127.0), (F.Q_INT4, 7.0)])
def test_quantised_apr_payloads(quant, levels, mel0):
    cfg = synth.CONFIGS["tiny"]
    data, _ = synth.random_model_apr(cfg, quant=quant, seed=0)
    w = F.AprReader(data).load_all()                    # dequantised exactly as format/mod.rs:632-672 / quantized.rs:1949-1969
    model = WhisperApr.load_from_apr(data)
    assert model.config.quantization == quant
    got = model.encode(mel0[:1000])
    ref = E.forward_mel(mel0[:1000], w, E.CONFIGS["tiny"], attention=E.naive_attention)
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS
    model.close()


def test_missing_tensors_keep_defaults(mel0, fb80):
    cfg = synth.CONFIGS["tiny"]
    tensors = [(n, a) for n, a in synth.random_encoder_tensors(cfg, 0) if "layers.2" not in n and "positional" not in n and n != "encoder.conv1.bias"]
    from whisper_apr_b200.apr_writer import write_apr
    data = write_apr(cfg, tensors, 0, fb80)
    model = WhisperApr.load_from_apr(data)              # lib.rs:769-800: absent tensors silently keep their defaults
    got = model.encode(mel0[:500])
    ref = E.forward_mel(mel0[:500], dict(tensors), E.CONFIGS["tiny"], attention=E.naive_attention)
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS
    model.close()


def test_base_config_parity(fb80):
    cfg = synth.CONFIGS["base"]
    data, tensors = synth.random_model_apr(cfg, seed=2)
    model = WhisperApr.load_from_apr(data)
    mel = M.compute_mel(synth.synth_audio(20), fb80)
    got = model.encode(mel)
    ref = E.forward_mel(mel, dict(tensors), E.CONFIGS["base"], dtype=np.float32, attention=E.naive_attention)
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS
    model.close()


def test_full_size_properties_large_v3_shape():
    """BASELINE.json's full shape (d=1280, 128 mel; depth cut to 2 layers to keep the oracle in seconds): parity on one chunk,
    plus size-independent properties on a batch: batch invariance, chunk independence and determinism."""
    cfg = synth.ModelConfig("large-v3-2l", 11, 128, 1500, 1280, 20, 2, 51865, 448, 1280, 20, 2)
    data, tensors = synth.random_model_apr(cfg, seed=4)
    model = WhisperApr.load_from_apr(data)
    audio = [synth.synth_audio(30 + i) for i in range(5)]
    out = model.mel_encode_batch(audio)
    assert out.shape == (5, 1500, 1280) and np.isfinite(out).all()
    ocfg = E.ModelConfig("l2", 11, 128, 1500, 1280, 20, 2)
    ref = E.forward_mel(M.compute_mel(audio[0], synth.load_filterbank(128)), dict(tensors), ocfg, dtype=np.float32, attention=E.naive_attention)
    assert np.abs(out[0] - ref).max() <= ENC_TOL and _cos(out[0], ref) >= ENC_COS
    again = model.mel_encode_batch(audio)
    assert np.array_equal(out, again)                                            # deterministic
    perm = model.mel_encode_batch([audio[3], audio[0]])
    assert np.array_equal(perm[0], out[3]) and np.array_equal(perm[1], out[0])   # chunks are independent units
    model.close()


@pytest.mark.gpu
def test_async_pipeline_matches_sync(tiny):
    """wb_mel_encode_batch_async over both staging slots (4 calls in flight) returns what the synchronous call returns."""
    model, _, _ = tiny
    d = model.config.n_audio_state
    batches = [[synth.synth_audio(10 * k + i)[: 480000 - 1000 * i].copy() for i in range(2)] for k in range(4)]
    want = [model.mel_encode_batch(b) for b in batches]
    outs = [np.zeros((2, 1500, d), np.float32) for _ in range(4)]
    for b, o in zip(batches, outs):
        model.mel_encode_batch_async(b, o)
    model.sync()
    for w, o in zip(want, outs):
        assert np.array_equal(w, o)


@pytest.mark.gpu
def test_streaming_chunks_int4(fb80):
    """BASELINE config 5 at test size: a long stream cut by split_into_chunks(5 s, 0.5 s overlap) (batch.rs:219-240), every chunk
    zero-padded to 30 s by compute_mel (lib.rs:413-420), int4 `.apr` payload, through the fused batch entry point."""
    cfg = synth.CONFIGS["tiny"]
    data, _ = synth.random_model_apr(cfg, quant=F.Q_INT4, seed=0)
    w = F.AprReader(data).load_all()
    model = WhisperApr.load_from_apr(data)
    stream = np.concatenate([synth.synth_audio(40), synth.synth_audio(41)])[:200000]
    from whisper_apr_b200 import api
    chunks = api.split_into_chunks(stream, 80000, 8000)
    assert [c.size for c in chunks] == [80000, 80000, 56000]
    assert np.array_equal(chunks[1], stream[72000:152000]) and np.array_equal(chunks[2], stream[144000:])
    out = model.mel_encode_batch(chunks)
    for i, c in enumerate(chunks):
        ref = E.forward_mel(M.compute_mel(c, fb80), w, E.CONFIGS["tiny"], attention=E.naive_attention)
        assert np.abs(out[i] - ref).max() <= ENC_TOL and _cos(out[i], ref) >= ENC_COS
    model.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "base"])
def test_greedy_tokens_identical(name, fb80):
    """North-star gate 3: the reference's greedy decode (oracle/decoder.py: forward_one + WhisperTokenSuppressor + argmax) emits the
    same tokens from the GPU's encoder states as from the oracle's.  Audio is the first synthetic chunk whose oracle decode has no
    near-tie step (top-2 logit margin >= 0.01): a coin-flip step is not reproducible by any implementation, the reference's own
    scalar and SIMD paths included."""
    from oracle import decoder as D
    cfg, ecfg = synth.CONFIGS[name], E.CONFIGS[name]
    data, tensors = synth.random_model_apr(cfg, seed=0)
    w = dict(tensors)
    dw = D.random_decoder_tensors(ecfg, seed=1)
    model = WhisperApr.load_from_apr(data)
    chosen = None
    for a in range(7, 19):
        audio = synth.synth_audio(a)
        ref_states = E.forward_mel(M.compute_mel(audio, fb80), w, ecfg, attention=E.naive_attention)
        toks, margins = D.greedy_decode(dw, ecfg, ref_states.astype(np.float32), max_tokens=24, return_margins=True)
        if min(margins) >= 0.01:
            chosen = (audio, ref_states, toks)
            break
    assert chosen is not None, "no synthetic chunk with a tie-free oracle decode"
    audio, ref_states, toks = chosen
    got_states = model.mel_encode_batch([audio])[0]
    assert np.abs(got_states - ref_states).max() <= ENC_TOL and _cos(got_states, ref_states) >= ENC_COS
    got = D.greedy_decode(dw, ecfg, got_states, max_tokens=24)
    assert got == toks
    assert len(set(toks[4:])) >= 3            # the sequence is not a degenerate repeat
    model.close()


@pytest.mark.gpu
def test_batch_preprocessor_ragged_htk(tiny):
    """BatchPreprocessor::process_batch (batch.rs:157-176): ragged segments, the preprocessor's own HTK filterbank (batch.rs:143),
    no padding; BatchMelResult bookkeeping and to_padded_tensor (batch.rs:107-127); the reference's own edge cases (empty / short)."""
    from whisper_apr_b200 import AudioBatch, BatchPreprocessor
    model, _, _ = tiny
    batch = AudioBatch()
    lens = [16000, 24000, 399, 0, 400, 5000]
    for i, n in enumerate(lens):
        batch.add_segment(synth.synth_audio(30 + i)[:n])
    res = BatchPreprocessor(model).process_batch(batch)
    assert res.frame_counts == [98, 148, 0, 0, 1, 29] and res.max_frames == 148 and len(res) == 6     # mel.rs:660-668 frame-count KATs
    fb = M.htk_filterbank(80)
    for seg, mel in zip(batch.segments, res.mels):
        ref = M.mel_compute(seg, fb)
        assert mel.shape == ref.shape
        if ref.size:
            assert np.abs(mel - ref).max() <= MEL_TOL
    padded = res.to_padded_tensor()
    assert padded.shape == (6, 80, 148)
    assert np.array_equal(padded, M.to_padded_tensor(res.mels, 80))
    norm = BatchPreprocessor.normalize_batch(batch)
    assert abs(float(np.abs(norm.segments[0]).max()) - 1.0) < 1e-6 and norm.segments[3].size == 0
    # a 128-mel preprocessor on the same model handle (tables are per n_mels)
    res128 = BatchPreprocessor(model, n_mels=128).process_batch(batch)
    assert np.abs(res128.mels[1] - M.mel_compute(batch.segments[1], M.htk_filterbank(128))).max() <= MEL_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("S", [1500, 700])
def test_attention_growing_scores_rescale_path(S):
    """Scores that grow along the keys force the running maximum up block after block, so the lazy rescale of O in tensor memory
    (only when the max grows by > 2^8) and the alpha-scaled row sum are exercised many times per row; a second head has its
    largest scores FIRST (no rescale after block 0, tiny later probabilities), a third is flat."""
    rng = np.random.default_rng(S)
    H, d = 3, 192
    qkv = rng.standard_normal((1, S, 3 * d)).astype(np.float32)
    ramp = np.linspace(0.0, 1.0, S, dtype=np.float32)[:, None]
    qkv[0, :, 0:64] = 1.0 + 0.1 * qkv[0, :, 0:64]                       # head 0 queries ~ all-ones
    qkv[0, :, d:d + 64] = (ramp * 9.0) + 0.05 * qkv[0, :, d:d + 64]       # keys: dot product grows to ~ 64 * 9 / 8 = 72 score units
    qkv[0, :, 64:128] = 1.0 + 0.1 * qkv[0, :, 64:128]
    qkv[0, :, d + 64:d + 128] = ((1.0 - ramp) * 9.0) + 0.05 * qkv[0, :, d + 64:d + 128]
    out = np.empty((1, S, d), np.float32)
    _lib.check(_lib.lib().wb_debug_attention(0, _p(qkv), 1, S, d, H, _p(out)))
    r = _bf16_round(qkv).astype(np.float64)
    for h in range(H):
        q, k, v = (r[0, :, o + h * 64: o + (h + 1) * 64] for o in (0, d, 2 * d))
        ref = E.naive_attention(q, k, v)
        got = out[0, :, h * 64:(h + 1) * 64]
        assert np.isfinite(got).all()
        assert np.abs(got - ref).max() <= 2.5e-2 and _cos(got, ref) > 0.9999


def test_n_audio_ctx_smaller_than_1500_is_refused_not_overrun(mel0, fb80):
    """A header with n_audio_ctx < 1500: the fused 30 s paths return the reference's own error (encoder.rs:456-461) instead of
    writing 1500 positions into buffers sized for the header's context; shorter mels still encode."""
    cfg = synth.ModelConfig("ctx1000", 0, 80, 1000, 384, 6, 2, 51865, 448, 384, 6, 2)
    tensors = [(n, a if n != "encoder.positional_embedding" else a[:1000]) for n, a in synth.random_encoder_tensors(synth.CONFIGS["tiny"], 0)
               if "layers.2" not in n and "layers.3" not in n]
    from whisper_apr_b200.apr_writer import write_apr
    model = WhisperApr.load_from_apr(write_apr(cfg, tensors, 0, fb80))
    assert model.config.n_audio_ctx == 1000
    with pytest.raises(WhisperError) as e:
        model.mel_encode_batch([synth.synth_audio(0)])
    assert e.value.kind == "Model" and "sequence length 1500 exceeds max 1000" in str(e.value)
    with pytest.raises(WhisperError) as e:
        model.encode(mel0[:2002])                                               # 1001 positions
    assert "exceeds max 1000" in str(e.value)
    got = model.encode(mel0[:2000])
    ocfg = E.ModelConfig("ctx1000", 0, 80, 1000, 384, 6, 2)
    ref = E.forward_mel(mel0[:2000], dict(tensors), ocfg, attention=E.naive_attention)
    assert got.shape == ref.shape == (1000, 384)
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS
    model.close()


def test_per_channel_int8_requantisation(mel0):
    """quantize_f32_to_i8_per_channel / dequantize_i8_to_f32_per_channel (src/model/quantized.rs:1769-1813) on the device: one scale per
    output channel of every linear weight, int8 rows resident in HBM, scale applied per output column in the GEMM epilogue."""
    cfg = synth.CONFIGS["tiny"]
    data, tensors = synth.random_model_apr(cfg, seed=0)
    model = WhisperApr.load_from_apr(data)
    model.requantize_int8_per_channel()
    w = dict(tensors)
    for name in list(w):
        if name.endswith("_proj.weight") or name.endswith("fc1.weight") or name.endswith("fc2.weight"):
            # the device quantises the bf16-rounded weights it holds
            a = _bf16_round(w[name]).astype(np.float32)
            absmax = np.abs(a).max(axis=1, keepdims=True)
            scale = np.where(absmax < 1e-10, np.float32(1.0), absmax / np.float32(127.0)).astype(np.float32)
            q = np.clip(np.sign(a / scale) * np.floor(np.abs(a / scale) + np.float32(0.5)), -128, 127)
            w[name] = (q * scale).astype(np.float32)
    got = model.encode(mel0[:1000])
    ref = E.forward_mel(mel0[:1000], w, E.CONFIGS["tiny"], attention=E.naive_attention)
    assert np.abs(got - ref).max() <= ENC_TOL and _cos(got, ref) >= ENC_COS
    plain = E.forward_mel(mel0[:1000], dict(tensors), E.CONFIGS["tiny"], attention=E.naive_attention)
    assert np.abs(got - plain).max() > np.abs(got - ref).max()                  # it really runs the quantised weights
    with pytest.raises(WhisperError):
        model.requantize_int8_per_channel()                                     # already quantised
    model.close()

