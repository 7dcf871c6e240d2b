"""Mel kernels alone on a B200: 32 chunks x 30 s, 80 and 128 mels; per-kernel time through wb_profile (CUDA events on the launching
stream) and parity of chunk 0 against the float64 oracle.  Usage: python tools/mel_bench.py [iters]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_apr_b200 import WhisperApr, synth
from oracle import mel as M

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = 32
audio = np.stack([synth.synth_audio(i) if i < 4 else np.roll(synth.synth_audio(i % 4), 777 * i) for i in range(B)])
d_audio = torch.from_numpy(audio).cuda()
for m in (80, 128):
    cfg = synth.ModelConfig(f"t{m}", 0, m, 1500, 384, 6, 1, 51865, 448, 384, 6, 1)
    data, _ = synth.random_model_apr(cfg, seed=1)
    model = WhisperApr.load_from_apr(data)
    model.set_max_batch(B)
    d_mel = torch.empty((B, 3000, m), dtype=torch.float32, device="cuda")
    for _ in range(2):
        model.compute_mel_batch_dev(d_audio.data_ptr(), B, d_mel.data_ptr())
    model.sync()
    model.profile_enable(True)
    for _ in range(iters):
        model.compute_mel_batch_dev(d_audio.data_ptr(), B, d_mel.data_ptr())
    prof = model.profile_read()
    model.profile_enable(False)
    got = d_mel[0].cpu().numpy()
    err = np.abs(got - M.compute_mel(audio[0], synth.load_filterbank(m))).max()
    stft, fin = prof["mel_stft"]["ms"] / iters, prof["mel_finalize"]["ms"] / iters
    nbytes = B * (4 * 480000 + 4 * 3000 * m)
    print(f"n_mels {m}: mel_stft {stft * 1e3:.1f} us, finalize {fin * 1e3:.1f} us; f32 in+out {nbytes / (stft + fin) / 1e6:.0f} GB/s; max-abs vs oracle {err:.2e}", flush=True)
    model.close()
