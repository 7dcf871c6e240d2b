"""Per-product operand-format error budget: CPU emulation (numpy f32 + explicit fp16 / bf16 rounding at the sites where the CUDA path rounds) of the
large-v3 shape at full depth with one product group at a time switched from fp16 to bf16 (its weights AND the activations it multiplies).
Test-side tooling (imports the oracle).  Output recorded in profiles/r02ar_attn_bf16_ab.txt.  Usage: python tools/format_error_budget.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import encoder as E, mel as M
from whisper_apr_b200 import synth
def r(x): return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()
def h(x): return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.float16).to(torch.float32).numpy()
def run(mel, w, cfg, bf):   # bf: set of groups in bf16 (weights AND the activations multiplied with them); everything else fp16
    d, H = cfg.d, cfg.n_audio_head
    g = lambda n, shape, dflt=0.0: E._get(w, n, shape, dflt)
    def W(n, grp): return (r if grp in bf else h)(np.asarray(w[n], np.float32))
    def A(x, grp): return (r if grp in bf else h)(x)
    x = h(np.asarray(mel, np.float32))
    x = h(E.gelu(E.conv1d(x, h(np.asarray(w["encoder.conv1.weight"], np.float32)), g("encoder.conv1.bias", (d,)), 1)))
    x = E.gelu(E.conv1d(x, h(np.asarray(w["encoder.conv2.weight"], np.float32)), g("encoder.conv2.bias", (d,)), 2))
    x = x + E.positional_embedding(w, cfg)[: x.shape[0]]
    for i in range(cfg.n_audio_layer):
        p = f"encoder.layers.{i}"
        n = A(E.layer_norm(x, g(f"{p}.self_attn_layer_norm.weight", (d,), 1.0), g(f"{p}.self_attn_layer_norm.bias", (d,))), "qkv")
        q = A(E.linear(n, W(f"{p}.self_attn.q_proj.weight", "qkv"), g(f"{p}.self_attn.q_proj.bias", (d,))), "attn")
        k = A(E.linear(n, W(f"{p}.self_attn.k_proj.weight", "qkv"), g(f"{p}.self_attn.k_proj.bias", (d,))), "attn")
        v = A(E.linear(n, W(f"{p}.self_attn.v_proj.weight", "qkv"), g(f"{p}.self_attn.v_proj.bias", (d,))), "attn")
        att = np.empty_like(q)
        for hd in range(H):
            sl = slice(hd * 64, (hd + 1) * 64)
            s = (q[:, sl] @ k[:, sl].T) * np.float32(0.125)
            pm = np.exp(s - s.max(axis=1, keepdims=True))
            l = pm.sum(axis=1, keepdims=True)
            att[:, sl] = (A(pm, "attn") @ v[:, sl]) / l
        att = A(att, "out")
        x = x + E.linear(att, W(f"{p}.self_attn.out_proj.weight", "out"), g(f"{p}.self_attn.out_proj.bias", (d,)))
        n = A(E.layer_norm(x, g(f"{p}.final_layer_norm.weight", (d,), 1.0), g(f"{p}.final_layer_norm.bias", (d,))), "fc1")
        hid = A(E.gelu(E.linear(n, W(f"{p}.fc1.weight", "fc1"), g(f"{p}.fc1.bias", (4 * d,)))), "fc2")
        x = x + E.linear(hid, W(f"{p}.fc2.weight", "fc2"), g(f"{p}.fc2.bias", (d,)))
    return E.layer_norm(x, g("encoder.layer_norm.weight", (d,), 1.0), g("encoder.layer_norm.bias", (d,)))
cfg = E.CONFIGS["large-v3"]
w = dict(synth.random_encoder_tensors(synth.CONFIGS["large-v3"], 0))
mel = M.compute_mel(synth.synth_audio(100), synth.load_filterbank(cfg.n_mels))
# f32 reference
import importlib
ref = None
def f32run():
    global r, h
    r0, h0 = r, h
    r = h = lambda x: np.ascontiguousarray(x, np.float32)
    out = run(mel, w, cfg, set())
    r, h = r0, h0
    return out
t0=time.time(); ref = f32run(); print(f"f32 pass {time.time()-t0:.0f}s", flush=True)
def rep(label, bf):
    out = run(mel, w, cfg, bf); e = np.abs(out - ref)
    print(f"{label:34s} max-abs {e.max():.4e}  rms {np.sqrt((e**2).mean()):.4e}", flush=True)
rep("all fp16", set())
rep("attention bf16 (q,k,v,p)", {"attn"})
rep("attn + qkv GEMM bf16", {"attn", "qkv"})
rep("attn + qkv + fc2 bf16", {"attn", "qkv", "fc2"})
rep("attn + qkv + fc1 + fc2 bf16", {"attn", "qkv", "fc1", "fc2"})
rep("fc1 + fc2 bf16", {"fc1", "fc2"})
rep("all bf16", {"attn", "qkv", "out", "fc1", "fc2"})
