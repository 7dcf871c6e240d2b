"""GPU bring-up: run each kernel family in its own process against the CPU oracle and print error statistics.

Usage (on a B200 box):  python tools/gpu_bringup.py [case ...]
Each case runs in a subprocess so a kernel trap in one does not poison the CUDA context of the others.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def bf16_round(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def stats(name, got, ref):
    got = np.asarray(got, np.float64).ravel()
    ref = np.asarray(ref, np.float64).ravel()
    err = np.abs(got - ref)
    cos = float(got @ ref / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
    print(f"  {name}: max_abs={err.max():.3e} mean_abs={err.mean():.3e} ref_absmax={np.abs(ref).max():.3e} cos={cos:.7f} "
          f"nan={int(np.isnan(got).sum())}", flush=True)
    return err.max(), cos


def case_layernorm():
    from whisper_apr_b200 import _lib
    from oracle import encoder as E
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for rows, d in [(5, 384), (1500, 1280), (33, 512), (7, 200)]:
        x = rng.standard_normal((rows, d)).astype(np.float32) * 3 + 1
        g = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
        b = (0.1 * rng.standard_normal(d)).astype(np.float32)
        out = np.empty_like(x)
        _lib.check(L.wb_debug_layernorm(0, _p(x), _p(g), _p(b), rows, d, _p(out)))
        stats(f"layernorm {rows}x{d}", out, E.layer_norm(x.astype(np.float64), g, b))


def case_gemm():
    from whisper_apr_b200 import _lib
    from oracle import encoder as E
    L = _lib.lib()
    rng = np.random.default_rng(1)
    shapes = [(128, 128, 64), (128, 256, 64), (256, 384, 128), (1500, 384, 384), (300, 1152, 384), (1000, 1280, 1280),
              (200, 512, 2048), (130, 256, 240)]
    for (M, N, K) in shapes:
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32)
        Ar, Wr = bf16_round(A).astype(np.float64), bf16_round(W).astype(np.float64)
        base = Ar @ Wr.T
        for epi, name in [(4, "f32"), (0, "bf16"), (1, "gelu_bf16"), (2, "resid"), (3, "gelu_pe")]:
            extra = rng.standard_normal((M, N)).astype(np.float32)
            out = np.empty((M, N), np.float32)
            t0 = time.time()
            st = L.wb_debug_gemm(0, _p(A), _p(W), _p(bias), _p(extra), M, N, K, epi, C.c_float(0.5), _p(out))
            if st != 0:
                print(f"  gemm {M}x{N}x{K} epi={name}: ERROR {L.wb_last_error().decode()}", flush=True)
                return
            ref = 0.5 * base + bias
            if epi == 1:
                ref = E.gelu(ref)
            elif epi == 2:
                ref = extra + ref
            elif epi == 3:
                ref = E.gelu(ref) + extra
            stats(f"gemm {M}x{N}x{K} {name} ({time.time() - t0:.2f}s)", out, ref)


def case_attention():
    from whisper_apr_b200 import _lib
    from oracle import encoder as E
    L = _lib.lib()
    rng = np.random.default_rng(2)
    for (B, S, H) in [(1, 128, 1), (1, 256, 2), (2, 1500, 6), (1, 77, 1), (1, 1499, 2)]:
        d = 64 * H
        qkv = (rng.standard_normal((B, S, 3 * d)) * 1.5).astype(np.float32)
        out = np.empty((B, S, d), np.float32)
        st = L.wb_debug_attention(0, _p(qkv), B, S, d, H, _p(out))
        if st != 0:
            print(f"  attention B{B} S{S} H{H}: ERROR {L.wb_last_error().decode()}", flush=True)
            return
        r = bf16_round(qkv).astype(np.float64)
        ref = np.zeros((B, S, d))
        for b in range(B):
            for h in range(H):
                q = r[b, :, h * 64:(h + 1) * 64]
                k = r[b, :, d + h * 64: d + (h + 1) * 64]
                v = r[b, :, 2 * d + h * 64: 2 * d + (h + 1) * 64]
                ref[b, :, h * 64:(h + 1) * 64] = E.naive_attention(q, k, v)
        stats(f"attention B{B} S{S} H{H}", out, ref)


def _tiny_model(quant=0, name="tiny"):
    from whisper_apr_b200 import WhisperApr, synth
    cfg = synth.CONFIGS[name]
    data, tensors = synth.random_model_apr(cfg, quant=quant, seed=0)
    return WhisperApr.load_from_apr(data), cfg, data


def case_mel():
    from whisper_apr_b200 import synth
    from oracle import mel as M
    model, cfg, _ = _tiny_model()
    fb = synth.load_filterbank(80)
    gold = np.fromfile(os.path.join(ROOT, "tests/golden/ref_a_audio.bin"), "<f4")
    got = model.mel_filters.compute(gold, 160)
    stats("mel_compute golden (24000 samples)", got, M.mel_compute(gold, fb, 160, precision="f64"))
    a = synth.synth_audio(0)
    t0 = time.time()
    got = model.compute_mel(a)
    print(f"  compute_mel wall {time.time() - t0:.3f}s")
    stats("compute_mel synth 30 s", got, M.compute_mel(a, fb))
    got = model.compute_mel(a[:80000])
    stats("compute_mel synth 5 s (padded)", got, M.compute_mel(a[:80000], fb))
    got = model.mel_filters.compute(a[:16000], 100)
    stats("mel_compute hop=100", got, M.mel_compute(a[:16000], fb, 100))
    got = model.mel_filters.compute(a[:50000], 517)
    stats("mel_compute hop=517", got, M.mel_compute(a[:50000], fb, 517))


def case_encoder():
    from whisper_apr_b200 import synth
    from oracle import apr_format as F
    from oracle import encoder as E
    from oracle import mel as M
    model, cfg, data = _tiny_model()
    w = F.AprReader(data).load_all()
    ocfg = E.CONFIGS["tiny"]
    fb = synth.load_filterbank(80)
    mel = M.compute_mel(synth.synth_audio(0), fb)
    x0 = E.conv_frontend(mel, w, ocfg) + E.positional_embedding(w, ocfg)[:1500]
    got = model.debug_encode(mel, n_layers=0, ln_post=False)
    stats("conv stem + pos-emb", got, x0)
    x = x0
    for i in range(ocfg.n_audio_layer):
        x = E.encoder_block(x, w, i, ocfg, attention=E.naive_attention)
        got = model.debug_encode(mel, n_layers=i + 1, ln_post=False)
        stats(f"after layer {i + 1}", got, x)
    ref = E.layer_norm(x, w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"])
    t0 = time.time()
    got = model.encode(mel)
    print(f"  encode wall {time.time() - t0:.3f}s")
    stats("encoder output (tiny, f32 apr)", got, ref)
    # short input: 1000 frames -> 500 positions
    got = model.encode(mel[:1001])
    ref_s = E.forward_mel(mel[:1001], w, ocfg, attention=E.naive_attention)
    stats("encoder output 1001 frames", got, ref_s)
    # fused batch entry
    audio = [synth.synth_audio(i) for i in range(3)]
    out = model.mel_encode_batch(audio)
    for i in range(3):
        r = E.forward_mel(M.compute_mel(audio[i], fb), w, ocfg, attention=E.naive_attention)
        stats(f"mel_encode_batch item {i}", out[i], r)


def case_quant():
    from whisper_apr_b200 import synth
    from oracle import apr_format as F
    from oracle import encoder as E
    from oracle import mel as M
    fb = synth.load_filterbank(80)
    mel = M.compute_mel(synth.synth_audio(1), fb)
    for quant, nm in [(2, "int8"), (3, "int4")]:
        model, cfg, data = _tiny_model(quant)
        w = F.AprReader(data).load_all()
        ref = E.forward_mel(mel, w, E.CONFIGS["tiny"], attention=E.naive_attention)
        stats(f"encoder output tiny {nm} apr", model.encode(mel), ref)


def case_attn_perf():
    """A full-size attention launch (8 chunks x 20 heads x 1500) for ncu: `ncu -k regex:attention ... --run attn_perf`."""
    from whisper_apr_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    B, S, H = 8, 1500, 20
    d = 64 * H
    qkv = rng.standard_normal((B, S, 3 * d)).astype(np.float32)
    out = np.empty((B, S, d), np.float32)
    for _ in range(3):
        t0 = time.time()
        _lib.check(L.wb_debug_attention(0, _p(qkv), B, S, d, H, _p(out)))
        print(f"  attn_perf call {time.time() - t0:.3f}s", flush=True)


CASES = {"attn_perf": case_attn_perf, "layernorm": case_layernorm, "gemm": case_gemm, "attention": case_attention, "mel": case_mel,
         "encoder": case_encoder, "quant": case_quant}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        CASES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        print(f"=== {n}", flush=True)
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], timeout=900)
        print(f"=== {n}: exit {r.returncode} in {time.time() - t0:.1f}s", flush=True)
