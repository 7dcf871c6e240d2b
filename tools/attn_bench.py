"""Attention kernel alone on a B200: parity against the float64 oracle on small cases, then time per launch at the
large-v3 bench shape.  Usage: python tools/attn_bench.py [B S d heads iters]   (env switches: WB_ATTN_OLD, WB_ATTN_POLY, WB_ATTN_NOTOKEN)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_apr_b200 import _lib
from oracle import encoder as E

L = _lib.lib()
def _p(a): return a.ctypes.data_as(C.c_void_p)
rng = np.random.default_rng(3)
for (B, S, d, H) in [(1, 128, 64, 1), (2, 300, 128, 2), (1, 1500, 384, 6)]:
    qkv = rng.standard_normal((B, S, 3 * d)).astype(np.float32)
    qkv[..., :d] *= 3.0
    out = np.empty((B, S, d), np.float32)
    _lib.check(L.wb_debug_attention(0, _p(qkv), B, S, d, H, _p(out)))
    import torch
    q16 = torch.from_numpy(qkv).to(torch.float16 if L.wb_operand_format() == b'fp16' else torch.bfloat16).to(torch.float64).numpy()
    ref = np.empty((B, S, d))
    for b in range(B):
        for h in range(H):
            q, k, v = (q16[b, :, i * d + h * 64:i * d + (h + 1) * 64] for i in range(3))
            s = q @ k.T / 8.0
            p = np.exp(s - s.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
            ref[b, :, h * 64:(h + 1) * 64] = p @ v
    err = np.abs(out - ref).max()
    print(f"parity B={B} S={S} d={d}: max_abs={err:.3e} ref_absmax={np.abs(ref).max():.2f} nan={int(np.isnan(out).sum())}", flush=True)
a = [int(x) for x in sys.argv[1:]] + [32, 1500, 1280, 20, 20][len(sys.argv) - 1:]
ms = C.c_float(0)
_lib.check(L.wb_debug_attention_bench(0, a[0], a[1], a[2], a[3], a[4], C.byref(ms)))
flops = 4.0 * a[0] * a[3] * a[1] * a[1] * 64
print(f"attention B={a[0]} S={a[1]} d={a[2]} heads={a[3]}: {ms.value:.4f} ms/launch  {flops / ms.value / 1e9:.1f} TFLOP/s", flush=True)
