"""GEMM kernel alone on a B200 at the large-v3 bench shapes (32 chunks x 1500 rows): ms per launch and TFLOP/s per shape.
Usage: python tools/gemm_bench.py [iters]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_apr_b200 import _lib
L = _lib.lib()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B, S, d = 32, 1500, 1280
tot = 0.0
for name, N, K, epi, per_layer in [("qkv", 3 * d, d, 0, 1), ("out_proj+resid", d, d, 2, 1), ("fc1+gelu", 4 * d, d, 1, 1), ("fc2+resid", d, 4 * d, 2, 1)]:
    ms = C.c_float(0)
    _lib.check(L.wb_debug_gemm_bench(0, B, S, N, K, epi, iters, C.byref(ms)))
    fl = 2.0 * B * S * N * K
    tot += ms.value * per_layer
    print(f"{name:16s} M={B*S} N={N} K={K}: {ms.value:.4f} ms  {fl / ms.value / 1e9:.1f} TFLOP/s", flush=True)
print(f"per layer {tot:.4f} ms -> x32 = {32 * tot:.2f} ms")
