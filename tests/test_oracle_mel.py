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


def test_prelog_pipeline_reproduces_reference_recorded_run():
    """test_data/mel_spectrogram.json opens with the recorded stdout of the reference's own `examples/mel_spectrogram` run:
    MelFilterbank::new(80, 400, 16000) (the HTK fallback bank, mel.rs:58-79,144-212) and MelFilterbank::compute on six
    deterministic signals, with `Frames` and `Total energy` = sum over mel bins of the frame-mean of |log-mel|.  That run
    predates today's log10 / clamp / scale tail (it printed |ln(max(E, 1e-10))|: silence gives 23.0259 = ln 1e10 per bin), so
    it pins everything IN FRONT of the logarithm -- HTK bank construction, periodic Hann, framing, 400-point FFT, power,
    filterbank product -- on numbers the reference itself produced.  The signals are rebuilt in f32 exactly as
    examples/mel_spectrogram.rs:50-103 builds them (the chaotic "white noise" included)."""
    sr = 16000
    fb = M.htk_filterbank(80).astype(np.float64)
    w = M.hann_window_periodic().astype(np.float64)
    i = np.arange(sr, dtype=np.float32)
    t = (i / np.float32(sr)).astype(np.float32)
    two_pi = np.float32(2.0) * np.float32(np.pi)

    def tone(f):
        return (np.sin((two_pi * np.float32(f) * t).astype(np.float32)).astype(np.float32) * np.float32(0.5)).astype(np.float32)

    x = (np.sin((i * np.float32(12345.6789)).astype(np.float32)).astype(np.float32) * np.float32(43758.5453)).astype(np.float32)
    noise = (((x - np.trunc(x)) * np.float32(2.0) - np.float32(1.0)) * np.float32(0.3)).astype(np.float32)
    harm = ((np.sin(two_pi * np.float32(150) * t) + np.sin(two_pi * np.float32(300) * t) * np.float32(0.5) +
             np.sin(two_pi * np.float32(450) * t) * np.float32(0.25) + np.sin(two_pi * np.float32(600) * t) * np.float32(0.125)) * np.float32(0.3)).astype(np.float32)

    def total(audio):
        n = M.n_frames_for(len(audio))
        idx = (np.arange(n) * 160)[:, None] + np.arange(400)[None, :]
        spec = np.fft.rfft(audio[idx].astype(np.float64) * w, axis=1)
        energy = (spec.real ** 2 + spec.imag ** 2) @ fb.T
        return n, float(np.abs(np.log(np.maximum(energy, 1e-10))).mean(0).sum())

    recorded = [("silence", np.zeros(sr, np.float32), 1842.0702, 2e-5), ("200 Hz", tone(200), 1521.4725, 2e-5), ("1000 Hz", tone(1000), 1296.1965, 2e-5),
                ("4000 Hz", tone(4000), 1104.2760, 2e-5), ("white noise", noise, 308.0254, 2e-5), ("speech-like", harm, 1164.4906, 1e-3)]
    for name, audio, ref_total, tol in recorded:
        n, tot = total(audio)
        assert n == 98, name                                                       # "Frames: 98"
        assert abs(tot - ref_total) <= tol * ref_total, (name, tot, ref_total)
    # "Mel scale examples" of the same run (hz_to_mel / mel_to_hz, mel.rs:144-160)
    for hz, mel in [(100.0, 150.49), (500.0, 607.45), (1000.0, 999.99), (2000.0, 1521.36), (4000.0, 2146.06), (8000.0, 2840.02)]:
        assert abs(float(M.hz_to_mel(hz)) - mel) < 0.006 and abs(float(M.mel_to_hz(M.hz_to_mel(hz))) - hz) < 0.01
