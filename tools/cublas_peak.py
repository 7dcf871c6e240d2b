"""cuBLAS (torch.matmul) sustained rate for bf16 AND fp16 operands on this box, the way MEASURED_PEAKS.json measures its bf16 figure
(8192^3, back to back for ~4 s under the power cap): the like-for-like denominator for a GEMM whose operands are fp16.
Usage: python tools/cublas_peak.py > gpurun_out/cublas_peak.json"""
import json, subprocess, threading, time
import torch

N = 8192
out = {"n": N}
for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16), ("bf16_again", torch.bfloat16), ("fp16_again", torch.float16)):
    a = (torch.randn(N, N, device="cuda") * 0.05).to(dt)
    b = (torch.randn(N, N, device="cuda") * 0.05).to(dt)
    for _ in range(5):
        a @ b
    torch.cuda.synchronize()
    clocks = []
    stop = False

    def sample():
        while not stop:
            try:
                r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout
                clocks.append([float(x) for x in r.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)
    th = threading.Thread(target=sample, daemon=True); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(50):
            a @ b
        iters += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop = True; th.join(timeout=1)
    ms = e0.elapsed_time(e1)
    cl = sorted(c[0] for c in clocks[2:]) or [0]
    out[name] = {"tflops_sustained": 2 * N ** 3 * iters / (ms * 1e-3) / 1e12, "sm_mhz_median": cl[len(cl) // 2], "power_w_max": max((c[1] for c in clocks), default=0)}
print(json.dumps(out))
