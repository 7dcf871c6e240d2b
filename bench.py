#!/usr/bin/env python
"""bench.py -- audio-seconds per second for mel + encoder (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            our arm   (torchrun launches it for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's CPU algorithm (oracle port) on host cores

A "step" is one pass of the hot path (log-mel of 30 s chunks + Whisper encoder forward) over one batch of synthetic
chunks per GPU.  Workload at every N: whisper-large-v3 shape (128 mel, 32 layers, d = 1280), bf16 tensor-core math with
fp32 accumulation / residual stream (16-bit operands: fp16 by default), 32 chunks per GPU per step (BASELINE.json configs[3]: 256 chunks over 8 GPUs), random-init
weights serialised through the `.apr` v1 writer, synthetic audio.  Weak scaling: per-GPU work is fixed, chunks shard by
batch, no data-path collective.

`value`  : whole-job throughput with inputs resident in HBM (wb_mel_encode_batch_dev), CUDA events, max over ranks.
`e2e`    : the same metric through the reference-facing C-ABI call with HOST buffers (wb_mel_encode_batch): pinned host audio in,
           pinned host encoder states out, copies inside the timed region.
`roofline`: dominant kernel (the tcgen05 GEMM): algorithmic FLOPs / CUDA-event time of its launches inside the step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-sec/sec, mel+encoder, whisper-large-v3 shape"
UNIT = "audio-s/s"
CHUNK_SECONDS = 30.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_id: str):
        self.rows, self.proc, self.gpu_id = [], None, gpu_id

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu_id, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        loaded = [r for r in self.rows if len(r) > 7 and r[7].replace(".", "", 1).isdigit() and float(r[7]) >= 50.0]
        for r in (loaded or self.rows):           # samples taken while the GPU was busy; all samples if none qualifies
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(cfg_name: str, threads: int):
    """One bounded sample of the reference's CPU algorithm (oracle/whisper_ref.c, the C restatement of the Rust loops) on one chunk of
    the bench workload.  Nothing inside a stage is sampled any more: the mel, the WHOLE conv stem (all 3000 frames), ONE COMPLETE
    encoder layer (LayerNorms, attention over all 1500 positions with `threads` threads over the heads as the reference's rayon
    feature does, and the scalar FFN over all 1500 rows) and the final LayerNorm are timed as they run; the chunk time is
    mel + stem + L x layer + ln_post -- the only extrapolation is "L identical layers", stated in the sample string.  whisper-tiny
    (4 layers) is timed end to end with no multiplication at all.  Returns (audio_s_per_s, description, seconds spent)."""
    from oracle import cref
    from oracle import encoder as E
    from whisper_apr_b200 import synth
    cfg = E.CONFIGS[cfg_name]
    d, L, H, m = cfg.n_audio_state, cfg.n_audio_layer, cfg.n_audio_head, cfg.n_mels
    t_all = time.perf_counter()
    audio = synth.synth_audio(0)
    fb = synth.load_filterbank(m)
    rng = np.random.default_rng(0)
    t0 = time.perf_counter(); mel = cref.compute_mel(audio, fb); t_mel = time.perf_counter() - t0
    if L <= 4:                                   # configs[0]: the whole chunk, every layer, as it runs
        w = dict(synth.random_encoder_tensors(synth.CONFIGS[cfg_name], 0))
        t0 = time.perf_counter(); cref.forward_mel(mel, w, cfg, threads=threads); t_enc = time.perf_counter() - t0
        t_chunk = t_mel + t_enc
        desc = (f"1 complete chunk of {cfg_name}, nothing sampled or extrapolated: mel {t_mel:.2f}s + conv stem, {L} layers and ln_post {t_enc:.2f}s "
                f"({threads} thread(s) over attention heads)")
        return CHUNK_SECONDS / t_chunk, desc, time.perf_counter() - t_all
    w = {"encoder.conv1.weight": (rng.standard_normal((d, m, 3)) / np.sqrt(3 * m)).astype(np.float32), "encoder.conv1.bias": np.zeros(d, np.float32),
         "encoder.conv2.weight": (rng.standard_normal((d, d, 3)) / np.sqrt(3 * d)).astype(np.float32), "encoder.conv2.bias": np.zeros(d, np.float32)}
    t0 = time.perf_counter(); x = cref.conv_stem(mel, w, cfg); t_stem = time.perf_counter() - t0
    lw = {}
    p = "encoder.layers.0"
    for k in ("q", "k", "v", "out"):
        lw[f"{p}.self_attn.{k}_proj.weight"] = (rng.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    lw[f"{p}.fc1.weight"] = (rng.standard_normal((4 * d, d)) / np.sqrt(d)).astype(np.float32)
    lw[f"{p}.fc2.weight"] = (rng.standard_normal((d, 4 * d)) / np.sqrt(4 * d)).astype(np.float32)
    lw[f"{p}.fc1.bias"] = np.zeros(4 * d, np.float32); lw[f"{p}.fc2.bias"] = np.zeros(d, np.float32)
    for n in ("self_attn_layer_norm", "final_layer_norm"):
        lw[f"{p}.{n}.weight"] = np.ones(d, np.float32); lw[f"{p}.{n}.bias"] = np.zeros(d, np.float32)
    W = cref.LayerWeights(lw, 0, d)
    t0 = time.perf_counter(); x = cref.encoder_layer(x, W, H, threads=threads); t_layer = time.perf_counter() - t0
    t0 = time.perf_counter(); cref.layernorm(x, W.ln1g, W.ln1b); t_ln = time.perf_counter() - t0
    t_chunk = t_mel + t_stem + L * t_layer + t_ln
    desc = (f"1 chunk of {cfg_name}: mel {t_mel:.2f}s, whole conv stem (3000 frames) {t_stem:.1f}s, ONE complete encoder layer (1500 positions: "
            f"attention with {threads} thread(s) over heads, scalar FFN on all 1500 rows) {t_layer:.1f}s, ln_post {t_ln:.3f}s -- all measured; "
            f"chunk = mel + stem + {L} x layer + ln_post = {t_chunk:.0f}s")
    return CHUNK_SECONDS / t_chunk, desc, time.perf_counter() - t_all


def make_config(args, world, extra=None):
    """The same keys in both arms (the driver compares the two config dicts)."""
    B = args.chunks
    cfg = {"workload": workload_name(args), "chunks_per_gpu_per_step": B, "global_chunks_per_step": world * B,
           "apr_payload": args.quant, "out_dtype": args.out_dtype,
           "l2": f"inputs rotate over 3 distinct batches ({3 * B * 1.92:.0f} MB) and each step streams ~2 GB of activations, both > 126 MB L2",
           "parallelism": f"dp{world} (chunks sharded by batch, replicated weights, no data-path collective)"}
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import cref
    threads = cref.max_threads()
    cref.lib()
    vals, desc, spent = [], "", 0.0
    budget_s = 200.0                                  # the whole run ends within a few minutes whatever --steps says
    for i in range(args.warmup + args.steps):
        v, desc, sec = cpu_sample(args.model, threads)
        spent += sec
        if i >= min(args.warmup, 1) or spent > budget_s:      # at most one untimed warm-up sample on the CPU
            vals.append(v)
        if spent > budget_s or len(vals) >= args.steps:
            break
    val = float(np.mean(vals))
    chunks = args.chunks
    out = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": min(args.warmup, 1),
           "ms_per_step": 1e3 * chunks * CHUNK_SECONDS / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": make_config(args, world),
           "note": "reference = the reference's own CPU algorithm (C restatement oracle/whisper_ref.c; the Rust crate cannot be built offline), "
                   "one process on the host cores; a step here is one measured chunk sample, ms_per_step is scaled to the arm's chunks per step; "
                   "throughput does not grow with --gpus",
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_name(args):
    return {"large-v3": "whisper-large-v3 shape (128 mel, 32 enc layers, d=1280), 16-bit tensor-core operands, 30 s chunks sharded by batch",
            }.get(args.model, f"whisper-{args.model} shape, 16-bit tensor-core operands, 30 s chunks sharded by batch")



def bench_gather(args, model, lib, dist, torch, rank, local_rank, world, B, S, d, dev_audio, ROT, timed, step_dev):
    """`value_with_gather`: the same timed step with d_out = the gather rank's buffer (+ this rank's chunk offset)."""
    import whisper_apr_b200
    chk = whisper_apr_b200._lib.check
    per_rank = B * S * d * 2
    info = {"to_rank": 0, "dtype": "bf16", "bytes_per_rank": per_rank}
    try:
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        base = C.c_void_p()
        if rank == 0:
            hb = (C.c_uint8 * 64)()
            chk(lib.wb_ipc_alloc(local_rank, world * per_rank, C.byref(base), hb))
            handle.copy_(torch.tensor(list(hb), dtype=torch.uint8))
        dist.broadcast(handle, 0)
        if rank != 0:
            hb = (C.c_uint8 * 64)(*handle.cpu().tolist())
            chk(lib.wb_ipc_open(local_rank, hb, C.byref(base)))
        mine = base.value + rank * per_rank

        def step_gather(i):
            model.mel_encode_batch_dev(dev_audio[i % ROT].data_ptr(), B, mine, "bf16")

        for i in range(2 * ROT):                   # every (input, destination) key twice: graphs captured before the timed region
            step_gather(i)
        torch.cuda.synchronize()
        ms_g = timed(step_gather, args.steps)
        # verification: what sits in rank 0's buffer at this rank's offset is what this rank computes locally for the same input
        last = (args.steps - 1) % ROT
        local = torch.empty((B, S, d), dtype=torch.bfloat16, device="cuda")
        model.mel_encode_batch_dev(dev_audio[last].data_ptr(), B, local.data_ptr(), "bf16")
        torch.cuda.synchronize()
        got = np.empty(B * S * d, np.uint16)
        chk(lib.wb_read_device(local_rank, C.c_void_p(mine), got.ctypes.data_as(C.c_void_p), got.nbytes))
        same = bool(np.array_equal(got, local.view(torch.int16).cpu().numpy().view(np.uint16).ravel()))
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        info.update({"how": "peer stores over NVLink by the final LayerNorm kernel into rank 0's buffer (CUDA IPC mapping), inside the step",
                     "ms_per_step": ms_g / args.steps, "value_with_gather": world * B * CHUNK_SECONDS * args.steps / (ms_g / 1e3),
                     "verified_bit_identical": bool(flag.item())})
        dist.barrier()
        if rank != 0:
            lib.wb_ipc_close(local_rank, base)
        else:
            torch.cuda.synchronize()
            lib.wb_ipc_free(local_rank, base)
    except Exception as e:                          # IPC unavailable in this container: NCCL all-gather of bf16 states, reported as such
        from whisper_apr_b200 import sharding
        out16 = torch.empty((B, S, d), dtype=torch.bfloat16, device="cuda")
        for _ in range(2):
            sharding.gather_states(out16, B * world)
        ms_g = timed(lambda i: sharding.gather_states(out16, B * world), 5)
        info.update({"how": f"fallback: nccl all_gather_into_tensor of bf16 states after the step (peer-store path failed: {e})",
                     "ms_per_step_gather_only": ms_g / 5})
    return info


def run_streaming(args, model, lib, torch, dist, stream, rank, world, cfg, load_s):
    """BASELINE configs[4]: streams cut into 5 s chunks with 0.5 s overlap (audio::split_into_chunks), every chunk zero-padded to 30 s by
    compute_mel -- here as VIEWS of the uploaded streams.  value: device-resident (wb_mel_encode_views_dev); e2e: wb_stream_encode_views
    with host streams in and host states out.  Throughput counts STREAM seconds (new audio), not padded chunk seconds."""
    import whisper_apr_b200
    from whisper_apr_b200 import synth
    chk = whisper_apr_b200._lib.check
    CH, OV, L = 80000, 8000, 152000
    n_str, d, S = args.streams, cfg.n_audio_state, 1500
    ROT = 3
    host_streams = [torch.empty((n_str, L), dtype=torch.float32).pin_memory() for _ in range(ROT)]
    for r in range(ROT):
        for i in range(n_str):
            base = torch.from_numpy(synth.synth_audio((rank * ROT + r) * 4 + (i % 4))[:L])
            host_streams[r][i] = base if i < 4 else torch.roll(base, 997 * i) * (0.5 + 0.005 * i)
    starts = [0, CH - OV]
    lens_c = [CH, L - (CH - OV)]
    n_chunks = n_str * len(starts)
    seg_off = torch.tensor([i * L + st for i in range(n_str) for st in starts], dtype=torch.int64, device="cuda")
    n_valid = torch.tensor([ln for _ in range(n_str) for ln in lens_c], dtype=torch.int32, device="cuda")
    dev_streams = [h.cuda(non_blocking=True) for h in host_streams]
    out_t = torch.float32 if args.out_dtype == "f32" else torch.bfloat16
    dev_out = torch.empty((n_chunks, S, d), dtype=out_t, device="cuda")
    host_out = torch.empty((n_chunks, S, d), dtype=out_t).pin_memory()
    code = 0 if args.out_dtype == "f32" else 1
    model.set_max_batch(args.chunks)
    torch.cuda.synchronize()

    def step_dev(i):
        chk(lib.wb_mel_encode_views_dev(model._h, C.c_void_p(dev_streams[i % ROT].data_ptr()), C.c_void_p(seg_off.data_ptr()), C.c_void_p(n_valid.data_ptr()),
                                        n_chunks, C.c_void_p(dev_out.data_ptr()), code))

    ptrs = [((C.c_void_p * n_str)(*[host_streams[r].data_ptr() + i * L * 4 for i in range(n_str)]), (C.c_size_t * n_str)(*[L] * n_str)) for r in range(ROT)]
    counts = (C.c_size_t * n_str)()
    total = C.c_size_t(0)

    def step_e2e(i):
        p, ln = ptrs[i % ROT]
        chk(lib.wb_stream_encode_views(model._h, p, ln, n_str, CH, OV, C.c_void_p(host_out.data_ptr()), code, n_chunks, counts, C.byref(total)))

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(steps):
            fn(i)
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(2 * ROT + max(args.warmup, 3)):
        step_dev(i)
    launches0 = lib.wb_launch_count()
    ms = timed(step_dev, args.steps)
    launches = lib.wb_launch_count() - launches0
    stream_seconds = n_str * L / 16000.0
    value = world * stream_seconds * args.steps / (ms / 1e3)
    e2e = None
    if not args.no_e2e:
        for i in range(3):
            step_e2e(i)
        ms_e = timed(step_e2e, args.steps)
        assert total.value == n_chunks
        esz = 4 if args.out_dtype == "f32" else 2
        e2e = {"value": world * stream_seconds * args.steps / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": n_str * L * 4,
               "d2h_bytes_per_step": n_chunks * S * d * esz, "ms_per_step": ms_e / args.steps,
               "note": "the streams are uploaded once (no 30 s padding, no duplicated overlap): 6.3x fewer H2D bytes than padded chunks"}
    if rank == 0:
        print(json.dumps({"metric": f"audio-sec/sec (stream seconds), mel+encoder, whisper-{args.model} {args.quant} streaming 5 s / 0.5 s overlap",
                          "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": lib.wb_operand_format().decode(), "data": "synthetic",
                          "config": {"workload": f"BASELINE configs[4]: whisper-{args.model} {args.quant} .apr payload, {n_str} streams per GPU x 9.5 s, 5 s chunks / 0.5 s overlap "
                                                 f"as views ({n_chunks} chunks per GPU per step, each a full 1500-position encoder pass as the reference runs it)",
                                     "micro_batch": args.chunks, "out_dtype": args.out_dtype},
                          "e2e": e2e, "gpu_launches": int(launches), "setup": {"load_s": round(load_s, 2)}}), flush=True)

def run_transcribe(args, model, lib, torch, dist, stream, rank, world, cfg, load_s):
    """SURVEY section 8(f1), measured: audio -> token ids through wb_transcribe_tokens_batch (mel + encoder + cross K/V + batched greedy
    loop on the device, encoder states never leave HBM).  value: generated tokens per second over the whole call; the decode share is
    the call minus an encoder-only pass of the same batch; its roofline is the HBM stream one token step has to read (f32 decoder
    weights + the op16 cross K/V of every chunk + the self-attention cache)."""
    import whisper_apr_b200
    from whisper_apr_b200 import synth
    chk = whisper_apr_b200._lib.check
    B = min(args.chunks, 32)
    d, S, Ld, V = cfg.n_text_state, 1500, cfg.n_text_layer, cfg.n_vocab
    T = args.max_tokens
    init = np.array([50258, 50259, 50359, 50363], np.int32)          # <|startoftranscript|><|en|><|transcribe|><|notimestamps|>
    ROT = 3
    host_audio = [torch.empty((B, synth.N_SAMPLES_30S), dtype=torch.float32).pin_memory() for _ in range(ROT)]
    for r in range(ROT):
        for i in range(B):
            host_audio[r][i] = torch.from_numpy(synth.synth_audio((rank * ROT + r) * 4 + i)) if i < 4 else torch.roll(host_audio[r][i % 4], 1000 * i) * (0.5 + 0.01 * i)
    ptrs = [((C.c_void_p * B)(*[h.data_ptr() + i * synth.N_SAMPLES_30S * 4 for i in range(B)]), (C.c_size_t * B)(*[synth.N_SAMPLES_30S] * B)) for h in host_audio]
    toks = np.empty((B, T), np.int32)
    lens = np.empty(B, np.int32)
    host_out = torch.empty((B, S, cfg.n_audio_state), dtype=torch.float32).pin_memory()
    model.set_max_batch(B)

    def step_transcribe(i):
        p, ln = ptrs[i % ROT]
        chk(lib.wb_transcribe_tokens_batch(model._h, p, ln, B, init.ctypes.data_as(C.c_void_p), init.size, T, 1, toks.ctypes.data_as(C.c_void_p),
                                           lens.ctypes.data_as(C.c_void_p)))
        return int((lens - init.size).sum()), int(lens.max()) - 1

    def step_encode(i):
        p, ln = ptrs[i % ROT]
        chk(lib.wb_mel_encode_batch(model._h, p, ln, B, C.c_void_p(host_out.data_ptr()), 0))
        return 0, 0

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        n_tok = n_steps = 0
        for i in range(steps):
            a, b = fn(i)
            n_tok += a
            n_steps += b
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n_tok, n_steps

    for i in range(max(args.warmup, 3)):
        step_transcribe(i)
        step_encode(i)
    launches0 = lib.wb_launch_count()
    ms, n_tok, n_steps = timed(step_transcribe, args.steps)
    launches = lib.wb_launch_count() - launches0
    ms_enc, _, _ = timed(step_encode, args.steps)
    ms_dec = max(ms - ms_enc, 1e-3)
    peaks = load_peaks()
    # one token step reads: every f32 decoder matrix once (14 d^2 per layer + the V x d vocabulary projection), the op16 cross K/V of all
    # B chunks (B * S * 2d per layer) and the f32 self-attention cache written so far (ignored: < 1 % at these lengths)
    step_bytes = 4.0 * (Ld * 14 * d * d + V * d) + 2.0 * Ld * B * S * 2 * d
    achieved = step_bytes * n_steps / (ms_dec * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({"metric": f"generated tokens/sec, audio -> token ids (mel + encoder + greedy decode), whisper-{args.model}",
                          "value": world * n_tok / (ms / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32 token path, " + lib.wb_operand_format().decode() + " encoder and cross K/V", "data": "synthetic",
                          "config": {"workload": f"SURVEY 8(f1): {B} chunks x 30 s per GPU per step through wb_transcribe_tokens_batch, greedy, max {T} tokens "
                                                 f"(random-init weights: sequences run until EOT or {T})", "chunks_per_gpu_per_step": B, "max_tokens": T},
                          "e2e": {"value": world * n_tok / (ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": B * synth.N_SAMPLES_30S * 4,
                                  "d2h_bytes_per_step": B * T * 4 + B * 4, "note": "the measured call takes HOST audio and returns HOST token ids"},
                          "decode": {"tokens_per_s": world * n_tok / (ms_dec / 1e3), "token_steps": n_steps, "us_per_token_step": ms_dec * 1e3 / max(n_steps, 1),
                                     "ms_encode_per_step": ms_enc / args.steps, "ms_decode_per_step": ms_dec / args.steps},
                          "roofline": {"kernel": "decoder token step (dec_linear / dec_cross_attn weight + K/V stream)", "bound": "hbm", "achieved": achieved,
                                       "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                                       "algorithmic_bytes_per_launch": step_bytes, "peak_source": peaks["source"]},
                          "gpu_launches": int(launches), "setup": {"load_s": round(load_s, 2)}}), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--chunks", type=int, default=32, help="30 s chunks per GPU per step")
    ap.add_argument("--quant", default="f32", choices=["f32", "int8", "int4"], help=".apr payload type (compute is bf16)")
    ap.add_argument("--out-dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--max-tokens", type=int, default=64, help="--workload transcribe: GreedyDecoder max_tokens (initial tokens included)")
    ap.add_argument("--workload", default="chunks", choices=["chunks", "streaming", "transcribe"],
                    help="chunks: 30 s chunks (the headline, BASELINE configs[1..3]); streaming: configs[4] -- per GPU `--streams` streams of 9.5 s "
                         "cut into 5 s chunks with 0.5 s overlap, read in place as views (wb_stream_encode_views)")
    ap.add_argument("--streams", type=int, default=64)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3                      # timing rule: >= 3 warm-up steps

    import torch
    import torch.distributed as dist
    import whisper_apr_b200
    from whisper_apr_b200 import WhisperApr, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = synth.CONFIGS[args.model]
    B, d, S, m, L = args.chunks, cfg.n_audio_state, 1500, cfg.n_mels, cfg.n_audio_layer
    quant = {"f32": 0, "int8": 2, "int4": 3}[args.quant]
    t0 = time.time()
    data, _ = synth.random_model_apr(cfg, quant=quant, seed=0, with_decoder=args.workload == "transcribe")
    t1 = time.time()
    model = WhisperApr.load_from_apr(data, device=local_rank)
    load_s = time.time() - t1                # wb_model_from_apr alone: .apr bytes in host memory -> weights resident in HBM
    apr_mb = len(data) / 1e6
    del data
    model.set_max_batch(B)
    stream = torch.cuda.Stream()             # a real (non-legacy) stream: the kernels, copies and timing events all go here
    torch.cuda.set_stream(stream)
    model.set_stream(stream.cuda_stream)
    lib = whisper_apr_b200.lib()
    setup_s = time.time() - t0
    if args.workload in ("streaming", "transcribe"):
        (run_streaming if args.workload == "streaming" else run_transcribe)(args, model, lib, torch, dist, stream, rank, world, cfg, load_s)
        model.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # inputs: ROT distinct batches so consecutive steps never re-read a hot L2 line
    ROT = 3
    host_audio = [torch.empty((B, synth.N_SAMPLES_30S), dtype=torch.float32).pin_memory() for _ in range(ROT)]
    for r in range(ROT):
        for i in range(B):
            base = synth.synth_audio((rank * ROT + r) * B + i) if i < 4 else None
            # 4 fully synthesised chunks per buffer, the rest are amplitude-scaled rolls of them (cheap, still distinct data)
            host_audio[r][i] = torch.from_numpy(base) if base is not None else torch.roll(host_audio[r][i % 4], 1000 * i) * (0.5 + 0.01 * i)
    dev_audio = [h.cuda(non_blocking=True) for h in host_audio]
    out_t = torch.float32 if args.out_dtype == "f32" else torch.bfloat16
    dev_out = torch.empty((B, S, d), dtype=out_t, device="cuda")
    host_out = [torch.empty((B, S, d), dtype=out_t).pin_memory() for _ in range(2)]
    torch.cuda.synchronize()

    def step_dev(i):
        model.mel_encode_batch_dev(dev_audio[i % ROT].data_ptr(), B, dev_out.data_ptr(), args.out_dtype)

    ptr_arrays = []
    for r in range(ROT):
        base = host_audio[r].data_ptr()
        ptr_arrays.append(((C.c_void_p * B)(*[base + i * synth.N_SAMPLES_30S * 4 for i in range(B)]), (C.c_size_t * B)(*[synth.N_SAMPLES_30S] * B)))
    code = 0 if args.out_dtype == "f32" else 1

    def step_e2e(i):
        # the public host-buffer call, enqueue form: this step's H2D, compute and D2H overlap the neighbouring steps' through the
        # library's two staging slots (a third call blocks until the oldest result has reached host memory); wb_sync ends the region
        ptrs, lens = ptr_arrays[i % ROT]
        whisper_apr_b200._lib.check(lib.wb_mel_encode_batch_async(model._h, ptrs, lens, B, C.c_void_p(host_out[i % 2].data_ptr()), code))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, drain=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(steps):
            fn(i)
        if drain is not None:
            drain()                     # host-blocking: every result of the region is in host memory before the end event
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # setup, not warm-up: the library captures a step into a CUDA graph the second time it sees a (pointers, batch) key, so every
    # rotating input buffer is shown to it twice before anything is timed (a capture costs ~10 ms of host time once per key)
    for i in range(2 * ROT):
        step_dev(i)
    torch.cuda.synchronize()
    for i in range(args.warmup):
        step_dev(i)
    gpu_id = str(torch.cuda.get_device_properties(local_rank).uuid)
    sampler = ClockSampler(gpu_id if gpu_id.startswith("GPU-") else "GPU-" + gpu_id)
    if rank == 0:
        sampler.start()
    launches0 = lib.wb_launch_count()
    ms = timed(step_dev, args.steps)
    launches = lib.wb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    value = world * B * CHUNK_SECONDS * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        for i in range(4):              # both staging slots seen twice: their graphs exist before the timed region
            step_e2e(i)
        model.sync()
        for i in range(max(args.warmup, 2)):
            step_e2e(i)
        model.sync()
        ms_e = timed(step_e2e, args.steps, drain=model.sync)
        esz = 4 if args.out_dtype == "f32" else 2
        e2e = {"value": world * B * CHUNK_SECONDS * args.steps / (ms_e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": B * synth.N_SAMPLES_30S * 4, "d2h_bytes_per_step": B * S * d * esz, "ms_per_step": ms_e / args.steps}

    # final NVLink gather of the step's encoder states to rank 0, INSIDE the step: every rank's final LayerNorm stores its
    # bf16 states straight into rank 0's buffer over NVLink (peer stores through a CUDA-IPC mapping; no NCCL call, no extra pass).
    gather = None
    if world > 1:
        gather = bench_gather(args, model, lib, dist, torch, rank, local_rank, world, B, S, d, dev_audio, ROT, timed, step_dev)
    # per-kernel timing (CUDA events on the launching stream) over two more steps of the same region
    peaks = load_peaks()
    model.profile_enable(True)
    PSTEPS = 2
    for i in range(PSTEPS):
        step_dev(i)
    prof = model.profile_read()
    model.profile_enable(False)
    gemm_flops = B * (2 * 3000 * 3 * m * d + 2 * 1500 * 3 * d * d + L * 24 * S * d * d)
    attn_flops = B * L * 4 * S * S * d
    mel_bytes = B * (4 * synth.N_SAMPLES_30S + 2 * 3000 * m)
    ln_bytes = B * S * d * ((2 * L) * 6 + 8)
    g_ms = prof["gemm"]["ms"] / PSTEPS
    a_ms = prof["attention"]["ms"] / PSTEPS
    mel_ms = (prof["mel_stft"]["ms"] + prof["mel_finalize"]["ms"]) / PSTEPS
    ln_ms = prof["layernorm"]["ms"] / PSTEPS
    step_ms = ms / args.steps
    traffic = None                      # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if args.model == "large-v3" and B == 32:
            traffic = tr["dram_bytes_per_launch_mean"]
    except Exception:
        pass
    roofline = {"kernel": "gemm2_tn_kernel (tcgen05, cta_group::2)", "bound": "tensor", "achieved": gemm_flops / (g_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": gemm_flops / (g_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"], "traffic": traffic,
                # operands once + outputs (bf16 activations, f32 residual read-modify-write), mean over the 4 GEMMs of a layer
                "algorithmic_bytes_per_launch": (44 * B * S * d + 24 * d * d) / 4 if traffic else None,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": prof["gemm"]["launches"] // PSTEPS, "ms_per_step": g_ms, "share_of_step": g_ms / step_ms}
    kernels = {
        "attention": {"bound": "tensor", "achieved": attn_flops / (a_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "frac": attn_flops / (a_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                      "ms_per_step": a_ms, "share_of_step": a_ms / step_ms},
        "mel": {"bound": "hbm", "achieved": mel_bytes / (mel_ms * 1e-3) / 1e9, "unit": "GB/s", "frac": mel_bytes / (mel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "ms_per_step": mel_ms, "share_of_step": mel_ms / step_ms, "bytes": "f32 audio in + bf16 mel out"},
        "layernorm": {"bound": "hbm", "achieved": ln_bytes / (ln_ms * 1e-3) / 1e9, "unit": "GB/s", "frac": ln_bytes / (ln_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "ms_per_step": ln_ms, "share_of_step": ln_ms / step_ms},
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import cref
            threads = cref.max_threads()
            v, desc, _ = cpu_sample(args.model, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        except Exception as e:      # the CPU baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {e}"}

    # parity of THIS configuration at its real depth, measured by tests/test_gpu_full_depth.py on a B200 and committed under profiles/
    parity = None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "parity_depth.json")))
        key = {"large-v3": "large-v3 32L B=32", "base": "base 6L B=64", "small": "small 12L int8", "medium": "medium 24L int4"}.get(args.model)
        if key in rec:
            fin = rec[key]["final"]
            first = fin.get("pos0", fin)
            parity = {"config": key, "max_abs": first["max_abs"], "cos": first["cos"], "gate": {"max_abs": 2e-2, "cos": 0.9999},
                      "source": "tests/test_gpu_full_depth.py (GPU vs float32 oracle, full depth), profiles/parity_depth.json"}
    except Exception:
        pass

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": lib.wb_operand_format().decode(), "dtype_note": "16-bit tensor-core operands (IEEE fp16 unless built with -DWB_OPERANDS_BF16), fp32 "
               "accumulation, fp32 residual stream and LayerNorm statistics; same tcgen05 kind::f16 rate and bytes as bf16 (DESIGN.md section 4)",
               "data": "synthetic",
               "config": make_config(args, world),
               "setup": {"load_s": round(load_s, 2), "apr_mb": round(apr_mb), "synth_and_load_s": round(setup_s, 1),
                         "note": "load_s = wb_model_from_apr alone (pinned in-place upload + on-device conversion)"},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
               "cpu_baseline": cpu, "gather": gather, "parity": parity}
        if gather and "value_with_gather" in gather:
            out["value_with_gather"] = gather["value_with_gather"]
        print(json.dumps(out), flush=True)
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
