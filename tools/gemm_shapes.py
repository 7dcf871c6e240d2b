"""Per-shape roofline of the layer GEMMs for one model width: time per launch alone, against BOTH bounds -- the tensor pipe (measured
cuBLAS bf16 burst rate, MEASURED_PEAKS.json) and HBM (operands once + output; the residual epilogue reads and writes the f32 stream).
Usage: python tools/gemm_shapes.py [model ...]      e.g.  base large-v3"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_apr_b200 import _lib, synth
L = _lib.lib()
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
TF, BW = pk["bf16_tflops"], pk["hbm_gbs"]
TFS = pk.get("bf16_tflops_sustained", TF)
for name in (sys.argv[1:] or ["base", "large-v3"]):
    cfg = synth.CONFIGS[name]
    d = cfg.n_audio_state
    B = 64 if d <= 512 else 32
    M = B * 1500
    print(f"== {name}: d = {d}, M = {M} rows ({B} chunks); bounds: tensor {TF:.0f} TFLOP/s (burst), HBM {BW:.0f} GB/s")
    for label, N, K, epi in [("qkv", 3 * d, d, 0), ("out_proj+resid", d, d, 2), ("fc1+gelu", 4 * d, d, 1), ("fc2+resid", d, 4 * d, 2)]:
        ms = C.c_float(0)
        _lib.check(L.wb_debug_gemm_bench(0, 1, M, N, K, epi, 30, C.byref(ms)))
        flops = 2.0 * M * N * K
        out_bytes = M * N * (8 if epi == 2 else 2)           # residual: f32 read + write; otherwise a 16-bit store
        byts = M * K * 2 + N * K * 2 + out_bytes
        t_tensor, t_hbm = flops / (TF * 1e12) * 1e3, byts / (BW * 1e9) * 1e3
        bound = "tensor" if t_tensor >= t_hbm else "HBM"
        print(f"  {label:15s} N={N:5d} K={K:5d}: {ms.value * 1e3:7.1f} us   tensor floor {t_tensor * 1e3:6.1f} us, HBM floor {t_hbm * 1e3:6.1f} us -> {bound}-bound, "
              f"{max(t_tensor, t_hbm) / ms.value:.2f} of its roofline ({flops / ms.value / 1e9:.0f} TFLOP/s = {flops / ms.value / 1e9 / TFS:.2f} of sustained, {byts / ms.value / 1e6:.0f} GB/s)", flush=True)
