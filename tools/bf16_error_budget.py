"""Where does the bf16 path's error against the f32 oracle come from?  CPU emulation (numpy f32 + explicit bf16 roundings at the
points where the CUDA path rounds) of the large-v3 shape at full depth, switching one rounding site off at a time.
Test-side tooling (imports the oracle); usage: python tools/bf16_error_budget.py [model] [audio_seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import encoder as E, mel as M
from whisper_apr_b200 import synth

def r(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()

def run(mel, w, cfg, sites):
    R = lambda name, x: r(x) if name in sites else x
    d, H = cfg.d, cfg.n_audio_head
    W = lambda n: R("weights", np.asarray(w[n], np.float32))
    g = lambda n, shape, dflt=0.0: E._get(w, n, shape, dflt)
    x = R("mel", np.asarray(mel, np.float32))
    x = R("c1", E.gelu(E.conv1d(x, W("encoder.conv1.weight"), g("encoder.conv1.bias", (d,)), 1)))
    x = E.gelu(E.conv1d(x, W("encoder.conv2.weight"), g("encoder.conv2.bias", (d,)), 2))
    x = x + E.positional_embedding(w, cfg)[: x.shape[0]]
    for i in range(cfg.n_audio_layer):
        p = f"encoder.layers.{i}"
        n = R("ln", E.layer_norm(x, g(f"{p}.self_attn_layer_norm.weight", (d,), 1.0), g(f"{p}.self_attn_layer_norm.bias", (d,))))
        q = R("qkv", E.linear(n, W(f"{p}.self_attn.q_proj.weight"), g(f"{p}.self_attn.q_proj.bias", (d,))))
        k = R("qkv", E.linear(n, W(f"{p}.self_attn.k_proj.weight"), g(f"{p}.self_attn.k_proj.bias", (d,))))
        v = R("qkv", E.linear(n, W(f"{p}.self_attn.v_proj.weight"), g(f"{p}.self_attn.v_proj.bias", (d,))))
        att = np.empty_like(q)
        for h in range(H):
            sl = slice(h * 64, (h + 1) * 64)
            s = (q[:, sl] @ k[:, sl].T) * np.float32(0.125)
            pm = np.exp(s - s.max(axis=1, keepdims=True))
            l = pm.sum(axis=1, keepdims=True)
            att[:, sl] = (R("p", pm) @ v[:, sl]) / l
        att = R("att", att)
        x = x + E.linear(att, W(f"{p}.self_attn.out_proj.weight"), g(f"{p}.self_attn.out_proj.bias", (d,)))
        n = R("ln", E.layer_norm(x, g(f"{p}.final_layer_norm.weight", (d,), 1.0), g(f"{p}.final_layer_norm.bias", (d,))))
        hid = R("hid", E.gelu(E.linear(n, W(f"{p}.fc1.weight"), g(f"{p}.fc1.bias", (4 * d,)))))
        x = x + E.linear(hid, W(f"{p}.fc2.weight"), g(f"{p}.fc2.bias", (d,)))
    return E.layer_norm(x, g("encoder.layer_norm.weight", (d,), 1.0), g("encoder.layer_norm.bias", (d,)))

name = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cfg = E.CONFIGS[name]
w = dict(synth.random_encoder_tensors(synth.CONFIGS[name], 0))
mel = M.compute_mel(synth.synth_audio(seed), synth.load_filterbank(cfg.n_mels))
ALL = {"weights", "mel", "c1", "ln", "qkv", "p", "att", "hid"}
t0 = time.time(); ref = run(mel, w, cfg, set()); print(f"f32 reference pass {time.time() - t0:.0f}s", flush=True)
def rep(label, sites):
    out = run(mel, w, cfg, sites)
    e = np.abs(out - ref)
    print(f"{label:28s} max-abs {e.max():.4e}  rms {np.sqrt((e ** 2).mean()):.4e}  p99.99 {np.quantile(e, 0.9999):.4e}", flush=True)
rep("all sites bf16", ALL)
for s in sorted(ALL):
    rep(f"all but {s}", ALL - {s})
for s in sorted(ALL):
    rep(f"only {s}", {s})
