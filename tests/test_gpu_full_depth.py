"""Full-depth parity on BASELINE.json's own configurations (-m gpu): the CUDA path through the C ABI against the CPU oracle
at the REAL depth and width of every measured model -- not a depth-cut stand-in.

  * whisper-large-v3 shape, 32 layers, bf16 (configs[3], the bench workload): one chunk through the fused batch entry point
    at positions 0, 17 and 31 of a 32-chunk batch (= one bench step), plus the residual stream after every 4th layer;
  * whisper-small, 12 layers, int8 `.apr` payload (configs[2]);
  * whisper-medium, 24 layers, int4 `.apr` payload (configs[4]);
  * whisper-base, 6 layers, 64-chunk batch (configs[1]) at positions 0 and 63.

Gates (north_star): encoder hidden states <= 2e-2 max-abs and cosine >= 0.9999 against the oracle.  The oracle pass is float32
numpy (what the reference's f32 Rust produces up to summation order; ~2.3 TFLOP for large-v3, tens of seconds on the host
cores).  Every measured pair is written to gpurun_out/parity_depth.json (copied to profiles/ and quoted by bench.py).
"""
import json
import os
import time

import numpy as np
import pytest

from oracle import apr_format as F
from oracle import encoder as E
from oracle import mel as M
from whisper_apr_b200 import WhisperApr, synth

pytestmark = pytest.mark.gpu

ENC_TOL, ENC_COS = 2e-2, 0.9999
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORD = os.path.join(ROOT, "gpurun_out", os.environ.get("WB_PARITY_RECORD", "parity_depth.json"))


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))


def _record(key, value):
    try:
        os.makedirs(os.path.dirname(RECORD), exist_ok=True)
        cur = json.load(open(RECORD)) if os.path.exists(RECORD) else {}
        cur[key] = value
        json.dump(cur, open(RECORD, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _oracle_stages(mel, w, cfg, every):
    """Residual stream after the conv stem (+ positional embedding) and after every `every`-th block, then the final states
    (Encoder::forward, encoder.rs:450-478), float32."""
    x = E.conv_frontend(mel, w, cfg, np.float32)
    x = x + E.positional_embedding(w, cfg)[: x.shape[0]].astype(np.float32)
    stages = {0: x.copy()}
    for i in range(cfg.n_audio_layer):
        x = E.encoder_block(x, w, i, cfg, attention=E.naive_attention)
        if (i + 1) % every == 0 or i + 1 == cfg.n_audio_layer:
            stages[i + 1] = x.copy()
    d = cfg.d
    out = E.layer_norm(x, E._get(w, "encoder.layer_norm.weight", (d,), 1.0), E._get(w, "encoder.layer_norm.bias", (d,)))
    return stages, out.astype(np.float32)


def _depth_report(model, mel, stages):
    """max-abs / relative / cosine of the GPU residual stream against the oracle's at every recorded depth."""
    rows = []
    for n_layers, ref in sorted(stages.items()):
        got = model.debug_encode(mel, n_layers=n_layers, ln_post=False)
        err = float(np.abs(got - ref).max())
        rows.append({"layers": n_layers, "max_abs": err, "ref_absmax": float(np.abs(ref).max()), "rel": err / float(np.abs(ref).max()),
                     "cos": _cos(got, ref)})
    return rows


def test_large_v3_full_depth_bf16_batch32():
    cfg, ocfg = synth.CONFIGS["large-v3"], E.CONFIGS["large-v3"]
    t0 = time.time()
    data, tensors = synth.random_model_apr(cfg, seed=0)
    w = dict(tensors)
    model = WhisperApr.load_from_apr(data)
    del data
    t_load = time.time() - t0
    audio = synth.synth_audio(100)
    mel = M.compute_mel(audio, synth.load_filterbank(128))
    t0 = time.time()
    stages, ref = _oracle_stages(mel, w, ocfg, every=4)
    t_oracle = time.time() - t0
    # one bench step: 32 chunks, the checked chunk at positions 0, 17 and 31
    others = [synth.synth_audio(200 + i) for i in range(4)]
    batch = [others[i % 4] for i in range(32)]
    for pos in (0, 17, 31):
        batch[pos] = audio
    model.set_max_batch(32)
    out = model.mel_encode_batch(batch)
    assert out.shape == (32, 1500, 1280) and np.isfinite(out).all()
    res = {}
    for pos in (0, 17, 31):
        err, cos = float(np.abs(out[pos] - ref).max()), _cos(out[pos], ref)
        res[f"pos{pos}"] = {"max_abs": err, "cos": cos}
        assert err <= ENC_TOL and cos >= ENC_COS, (pos, err, cos)
    assert np.array_equal(out[0], out[17]) and np.array_equal(out[0], out[31])      # the position in the batch does not matter
    bf = (model.mel_encode_batch(batch[:2], out_dtype="bf16").astype(np.uint32) << 16).view(np.float32)
    res["bf16_out"] = {"max_abs": float(np.abs(bf[0] - ref).max()), "cos": _cos(bf[0], ref)}
    assert res["bf16_out"]["max_abs"] <= 3e-2 and res["bf16_out"]["cos"] >= ENC_COS      # + one bf16 rounding of O(4) values
    growth = _depth_report(model, mel, stages)
    for r in growth:
        print(f"[large-v3 depth] after {r['layers']:2d} layers: max-abs {r['max_abs']:.3e} (|x| <= {r['ref_absmax']:.1f}, rel {r['rel']:.2e}) cos {r['cos']:.7f}")
        assert r["cos"] >= ENC_COS and r["rel"] <= 2e-2
    # the gate is a max over 1.9 M values of an error whose rms is ~4e-3: two more chunks show the spread
    more = [synth.synth_audio(s) for s in (105, 106)]
    out_more = model.mel_encode_batch(more)
    for s, a, o in zip((105, 106), more, out_more):
        r = E.forward_mel(M.compute_mel(a, synth.load_filterbank(128)), w, ocfg, dtype=np.float32, attention=E.naive_attention)
        e = np.abs(o - r)
        res[f"audio{s}"] = {"max_abs": float(e.max()), "rms": float(np.sqrt((e ** 2).mean())), "cos": _cos(o, r)}
        print(f"[large-v3 audio {s}] max-abs {e.max():.4e} rms {np.sqrt((e ** 2).mean()):.3e} cos {res[f'audio{s}']['cos']:.7f}")
    e0 = np.abs(out[0] - ref)
    res["pos0"]["rms"] = float(np.sqrt((e0 ** 2).mean()))
    _record("large-v3 32L B=32", {"operand_format": __import__("whisper_apr_b200")._lib.lib().wb_operand_format().decode(), "final": res, "depth": growth, "oracle": "numpy float32, naive softmax attention", "oracle_s": round(t_oracle, 1),
                                       "load_s": round(t_load, 1), "gate": {"max_abs": ENC_TOL, "cos": ENC_COS}})
    for s in (105, 106):
        assert res[f"audio{s}"]["max_abs"] <= ENC_TOL and res[f"audio{s}"]["cos"] >= ENC_COS, res
    model.close()


@pytest.mark.parametrize("name,quant,every", [("small", F.Q_INT8, 4), ("medium", F.Q_INT4, 8)])
def test_quantised_models_full_depth(name, quant, every):
    cfg, ocfg = synth.CONFIGS[name], E.CONFIGS[name]
    data, _ = synth.random_model_apr(cfg, quant=quant, seed=0)
    w = F.AprReader(data).load_all()                 # dequantised as the reference reader does (format/mod.rs:632-672, quantized.rs:1949-1969)
    model = WhisperApr.load_from_apr(data)
    assert model.config.quantization == quant and model.config.n_audio_layer == ocfg.n_audio_layer
    audio = synth.synth_audio(101)
    mel = M.compute_mel(audio, synth.load_filterbank(80))
    stages, ref = _oracle_stages(mel, w, ocfg, every=every)
    out = model.mel_encode_batch([synth.synth_audio(102), audio, synth.synth_audio(103)])
    err, cos = float(np.abs(out[1] - ref).max()), _cos(out[1], ref)
    growth = _depth_report(model, mel, stages)
    for r in growth:
        print(f"[{name} depth] after {r['layers']:2d} layers: max-abs {r['max_abs']:.3e} rel {r['rel']:.2e} cos {r['cos']:.7f}")
    _record(f"{name} {ocfg.n_audio_layer}L {'int8' if quant == F.Q_INT8 else 'int4'}", {"final": {"max_abs": err, "cos": cos}, "depth": growth})
    assert err <= ENC_TOL and cos >= ENC_COS, (err, cos)
    for r in growth:
        assert r["cos"] >= ENC_COS and r["rel"] <= 2e-2
    model.close()


def test_base_batch64_positions():
    cfg, ocfg = synth.CONFIGS["base"], E.CONFIGS["base"]
    data, tensors = synth.random_model_apr(cfg, seed=0)
    model = WhisperApr.load_from_apr(data)
    audio = synth.synth_audio(104)
    ref = E.forward_mel(M.compute_mel(audio, synth.load_filterbank(80)), dict(tensors), ocfg, dtype=np.float32, attention=E.naive_attention)
    others = [synth.synth_audio(300 + i) for i in range(3)]
    batch = [others[i % 3] for i in range(64)]
    batch[0] = batch[63] = audio
    model.set_max_batch(64)
    out = model.mel_encode_batch(batch)
    res = {}
    for pos in (0, 63):
        err, cos = float(np.abs(out[pos] - ref).max()), _cos(out[pos], ref)
        res[f"pos{pos}"] = {"max_abs": err, "cos": cos}
        assert err <= ENC_TOL and cos >= ENC_COS
    assert np.array_equal(out[0], out[63])
    _record("base 6L B=64", {"final": res})
    model.close()
