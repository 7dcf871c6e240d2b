"""Pins the mel oracle against the reference's own golden vectors and known-answer tests (CPU only).

Golden: test_data/ref_a_audio.bin -> ref_c_mel_numpy.bin (made by tools/extract_ground_truth.py:146-178 with a
SYMMETRIC np.hanning window); KATs from src/audio/mel.rs tests (:633-698 frame counts / errors, :786-799 silence,
:1037-1118 tone localisation, :1120-1149 loudness, :1179-1196 determinism) and tests/ground_truth_tests.rs:501-504,
657-662 (mean/std within 10 %).
"""
import numpy as np
import pytest

from oracle import mel as M


def test_golden_exact_with_generating_window(golden_audio, golden_mel, fb80):
    got = M.mel_compute(golden_audio, fb80, 160, window=np.hanning(400), precision="f64")
    assert got.shape == (148, 80)
    assert np.abs(got - golden_mel).max() <= 5e-7      # loader + framing + filterbank + normalisation pinned


def test_golden_with_reference_periodic_window(golden_audio, golden_mel, fb80):
    got = M.mel_compute(golden_audio, fb80, 160, precision="f64")
    err = np.abs(got - golden_mel).max()
    assert 0.005 < err < 0.02                           # 0.0138: periodic vs symmetric window
    cos = float((got * golden_mel).sum() / np.linalg.norm(got) / np.linalg.norm(golden_mel))
    assert cos > 0.99999
    # the reference's own gate: statistics within 10 % (tests/ground_truth_tests.rs:657-662)
    assert abs(got.mean() - (-0.214805)) < 0.1 * 0.214805
    assert abs(got.std() - 0.447922) < 0.1 * 0.447922


def test_f32_restatement_close_to_f64(golden_audio, fb80):
    a = M.mel_compute(golden_audio, fb80, 160, precision="f64")
    b = M.mel_compute(golden_audio, fb80, 160, precision="f32")
    assert np.abs(a - b).max() < 2e-5                   # mel.rs:906-928 asserts scalar == simd within 1e-4


@pytest.mark.parametrize("n,frames", [(16000, 98), (400, 1), (100, 0), (399, 0), (560, 2), (480000, 2998)])
def test_frame_count_kats(n, frames, fb80):
    assert M.n_frames_for(n) == frames                  # mel.rs:660-668
    if n <= 16000:
        assert M.mel_compute(np.zeros(n, np.float32) + 0.1, fb80, 160).shape[0] == frames


def test_empty_and_hop_zero(fb80):
    assert M.mel_compute(np.zeros(0, np.float32), fb80, 160).shape == (0, 80)        # mel.rs:633-640
    with pytest.raises(ValueError):
        M.mel_compute(np.ones(1000, np.float32), fb80, 0)                              # mel.rs:690-698


def test_window_is_periodic_hann():
    w = M.hann_window_periodic(400)
    assert w[0] == 0.0 and abs(w[200] - 1.0) < 1e-6 and w[399] > 0                     # periodic: w[N-1] != 0
    assert np.abs(w[1:] - w[1:][::-1]).max() < 1e-6                                    # w[n] == w[N-n]


def test_silence_all_negative(fb80):
    out = M.mel_compute(np.zeros(16000, np.float32), fb80, 160)
    assert (out < 0).all()                                                             # mel.rs:786-799
    assert np.allclose(out, (-10 + 4) / 4)


def _tone(f, n=16000):
    t = np.arange(n) / 16000.0
    return np.sin(2 * np.pi * f * t).astype(np.float32)


def test_tone_localisation_htk():
    fb = M.htk_filterbank(80)                                                          # MelFilterbank::new
    lo = M.mel_compute(_tone(440.0), fb, 160).mean(axis=0).argmax()
    hi = M.mel_compute(_tone(4000.0), fb, 160).mean(axis=0).argmax()
    assert 10 <= lo <= 35                                                              # mel.rs:1037-1075
    assert hi >= 40 and hi > lo                                                        # mel.rs:1078-1118


def test_louder_is_larger(fb80):
    q = M.mel_compute(0.1 * _tone(440.0), fb80, 160, precision="f32")
    l = M.mel_compute(0.9 * _tone(440.0), fb80, 160, precision="f32")
    assert q.shape == l.shape                                                          # mel.rs:1120-1149 (pre-normalisation energy
    # is monotone in amplitude; after the max-8 clamp both saturate identically, so compare un-normalised energy)
    e_q = (np.abs(np.fft.rfft(0.1 * _tone(440.0)[:400] * M.hann_window_periodic())) ** 2).sum()
    e_l = (np.abs(np.fft.rfft(0.9 * _tone(440.0)[:400] * M.hann_window_periodic())) ** 2).sum()
    assert e_l > e_q


def test_determinism(golden_audio, fb80):
    a = M.mel_compute(golden_audio, fb80, 160, precision="f32")
    b = M.mel_compute(golden_audio, fb80, 160, precision="f32")
    assert np.array_equal(a, b)                                                        # mel.rs:1179-1196


def test_compute_mel_padding(golden_audio, fb80, fb128):
    out = M.compute_mel(golden_audio, fb80)
    assert out.shape == (3000, 80)
    assert (out[2998:] == -1.0).all()                                                  # lib.rs:431-437
    # silent tail frames are clamped to gmax - 8
    assert np.allclose(out[200:2998], out[200, 0])
    out128 = M.compute_mel(golden_audio, fb128)
    assert out128.shape == (3000, 128)
    long = np.concatenate([golden_audio] * 21)[:500000]
    assert M.compute_mel(long, fb80).shape == (3000, 80)                               # truncation, lib.rs:421-424


def test_split_into_chunks_kats():
    s = np.arange(10, dtype=np.float32)
    c = M.split_into_chunks(s, 4, 1)                                                   # batch.rs:219-240
    assert [len(x) for x in c] == [4, 4, 4] and c[1][0] == 3 and c[2][-1] == 9
    assert M.split_into_chunks(s, 0, 0) == [] and M.split_into_chunks(s[:0], 4, 0) == []
    assert [len(x) for x in M.split_into_chunks(s, 4, 0)] == [4, 4, 2]
    assert len(M.split_into_chunks(s, 3, 5)) == 8                                      # overlap >= chunk -> step 1
    # config 5 of BASELINE.json: 5 s chunks, 0.5 s overlap
    assert [len(x) for x in M.split_into_chunks(np.zeros(160000, np.float32), 80000, 8000)] == [80000, 80000, 16000]


def test_to_padded_tensor():
    a = np.arange(6, dtype=np.float32).reshape(3, 2)
    b = np.arange(4, dtype=np.float32).reshape(2, 2) + 10
    t = M.to_padded_tensor([a, b], 2)                                                  # batch.rs:107-127
    assert t.shape == (2, 2, 3)
    assert np.array_equal(t[0], a.T) and np.array_equal(t[1, :, :2], b.T) and (t[1, :, 2] == 0).all()
