"""The C++ host-side mirror of the reference's interface (include/whisper_b200.hpp) from a compiled caller: built with g++ against the
C ABI only (no Python, no torch in the process), CPU checks without a GPU, the real path on cuda:0 compared bit for bit with the
ctypes mirror."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _build(tmp_path):
    import whisper_apr_b200
    if not os.path.exists(whisper_apr_b200.LIB_PATH):
        whisper_apr_b200.build()
    exe = str(tmp_path / "host_mirror")
    libdir = os.path.dirname(whisper_apr_b200.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"),
           "-o", exe, "-L", libdir, "-l:" + os.path.basename(whisper_apr_b200.LIB_PATH), "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_mirror_builds_and_cpu_checks(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "cpu checks ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_ctypes_mirror(tmp_path):
    from whisper_apr_b200 import WhisperApr, synth
    from oracle import decoder as D
    exe = _build(tmp_path)
    cfg = synth.CONFIGS["tiny"]
    data, _ = synth.random_model_apr(cfg, seed=0, with_decoder=True)
    audio = synth.synth_audio(0)
    (tmp_path / "m.apr").write_bytes(bytes(data))
    audio.astype(np.float32).tofile(tmp_path / "a.f32")
    r = subprocess.run([exe, str(tmp_path / "m.apr"), str(tmp_path / "a.f32"), str(tmp_path / "out")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "gpu run ok" in r.stdout, r.stdout + r.stderr
    model = WhisperApr.load_from_apr(data, device=0)
    mel = np.fromfile(tmp_path / "out.mel", np.float32).reshape(3000, cfg.n_mels)
    assert np.array_equal(mel, model.compute_mel(audio))
    states = np.fromfile(tmp_path / "out.states", np.float32).reshape(2, 1500, cfg.n_audio_state)
    ref = model.mel_encode_batch([audio, audio[: audio.size // 2]])
    assert np.array_equal(states, ref)
    from whisper_apr_b200 import api
    up = np.fromfile(tmp_path / "out.resampled", np.float32)
    assert np.array_equal(up, api.SincResampler(model, 8000, 16000).resample(audio[:8000]))
    views = np.fromfile(tmp_path / "out.views", np.float32).reshape(2, 1500, cfg.n_audio_state)
    ref_views, counts = api.stream_encode_views(model, [audio[:152000]], 80000, 8000)
    assert list(counts) == [2] and np.array_equal(views, ref_views)
    toks = np.fromfile(tmp_path / "out.tokens", np.int32).tolist()
    assert toks == model.transcribe_tokens_batch([audio], D.initial_tokens(), 12)[0]
    model.close()
