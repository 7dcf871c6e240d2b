"""Oracle for the ingest stages (oracle/audio_pre.py) pinned on the reference's own known answers:
src/audio/wav.rs:484-672, 881-948 (parse_wav fixtures and conversions), src/audio/resampler.rs:341-565 (length, DC, Bessel, window KATs),
src/vad.rs:934-1160 (frame features, silence, speech-like tone, process_frame), src/audio/streaming.rs:843-870 (get_chunk).
Also: the host WAV parser of the C ABI (no GPU needed) against the oracle on every fixture and error case."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import audio_pre as A
from whisper_apr_b200 import WhisperError, _lib, api


def test_parse_wav_fixtures_of_the_reference():
    d = A.parse_wav(A.make_wav(np.array([0, 16384, -16384, 32767, -32768]) / 32768.0 * 32768 / 32767, 16000))   # wav.rs:484-495
    raw = struct.pack("<5h", 0, 16384, -16384, 32767, -32768)
    wav = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16) + b"data" + struct.pack("<I", len(raw)) + raw
    d = A.parse_wav(wav)
    assert d.sample_rate == 16000 and d.samples.size == 5
    assert np.allclose(d.samples, [0.0, 0.5, -0.5, 32767 / 32768, -1.0])
    raw = struct.pack("<6h", 16384, -16384, 0, 0, 32767, -32767)                                                 # wav.rs:498-507 (stereo)
    wav = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 2, 44100, 176400, 4, 16) + b"data" + struct.pack("<I", len(raw)) + raw
    d = A.parse_wav(wav)
    assert d.sample_rate == 44100 and d.original_channels == 2 and np.allclose(d.samples, [0.0, 0.0, 0.0])
    raw = bytes([128, 255, 0, 192, 64])                                                                          # wav.rs:510-519 (8 bit)
    wav = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 8000, 8000, 1, 8) + b"data" + struct.pack("<I", len(raw)) + raw
    assert np.allclose(A.parse_wav(wav).samples, [0.0, 127 / 128, -1.0, 0.5, -0.5])
    raw = bytes([0, 0, 0, 0xFF, 0xFF, 0x7F, 0, 0, 0x80])                                                         # wav.rs:522-530 (24 bit)
    wav = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 48000, 144000, 3, 24) + b"data" + struct.pack("<I", len(raw)) + raw
    assert np.allclose(A.parse_wav(wav).samples, [0.0, 8388607 / 8388608, -1.0])
    f = np.array([0.0, 0.5, -0.5, 1.0, -1.0], np.float32)                                                         # wav.rs:533-543 (float)
    assert np.array_equal(A.parse_wav(A.make_wav(f, 16000, float_format=True)).samples, f)
    for bits in (24, 32):                                                                                         # wav.rs:881-948 (EXTENSIBLE)
        d = A.parse_wav(A.make_wav(np.array([0.0, 0.25, -0.25]), 48000, bits=bits, extensible=True))
        assert d.bits_per_sample == bits and np.allclose(d.samples, [0, 0.25, -0.25], atol=1e-6)
    d = A.parse_wav(A.make_wav(f, 16000, float_format=True, extensible=True))
    assert np.array_equal(d.samples, f)
    d = A.parse_wav(A.make_wav(np.array([0.1, -0.1]), 16000, extra_chunk=b"abc"))                                 # odd-sized unknown chunk is skipped, aligned
    assert d.samples.size == 2


@pytest.mark.parametrize("bad,msg", [(b"RIFF" + bytes(10), "too small"), (b"XXXX" + bytes(60), "RIFF"), (b"RIFF" + bytes(4) + b"XXXX" + bytes(60), "WAVE")])
def test_parse_wav_errors(bad, msg):
    with pytest.raises(A.WavError) as e:                                                                          # wav.rs:546-566
        A.parse_wav(bad)
    assert msg in str(e.value)
    with pytest.raises(WhisperError) as e2:
        api.parse_wav_header(bad)
    assert e2.value.kind == "Audio" and msg in str(e2.value)


def test_parse_wav_no_data_and_channels():
    fmt = struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16)
    wav = b"RIFF" + struct.pack("<I", 100) + b"WAVEfmt " + fmt + bytes(40)                                        # wav.rs:568-586
    for fn, exc in ((A.parse_wav, A.WavError), (api.parse_wav_header, WhisperError)):
        with pytest.raises(exc) as e:
            fn(wav)
        assert "no data chunk" in str(e.value)
    six = b"RIFF" + struct.pack("<I", 60) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 6, 16000, 192000, 12, 16) + b"data" + struct.pack("<I", 24) + bytes(24)
    for fn, exc in ((A.parse_wav, A.WavError), (api.parse_wav_header, WhisperError)):
        with pytest.raises(exc) as e:                                                                             # wav.rs:589-599
            fn(six)
        assert "channel count" in str(e.value) and "6" in str(e.value)
    adpcm = b"RIFF" + struct.pack("<I", 60) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 2, 1, 16000, 32000, 2, 4) + b"data" + struct.pack("<I", 24) + bytes(24)
    for fn, exc in ((A.parse_wav, A.WavError), (api.parse_wav_header, WhisperError)):
        with pytest.raises(exc) as e:
            fn(adpcm)
        assert "unsupported format 2 with 4 bits" in str(e.value)


def test_host_wav_header_parser_matches_oracle():
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 1001)
    for kw, kind in ((dict(bits=16), 1), (dict(bits=8), 0), (dict(bits=24), 2), (dict(bits=32), 3), (dict(float_format=True), 4),
                     (dict(bits=16, channels=2), 1), (dict(bits=24, extensible=True), 2), (dict(bits=16, extra_chunk=b"hello"), 1)):
        ch = kw.get("channels", 1)
        wav = A.make_wav(x[: 1000 if ch == 2 else 1001], 22050, **kw)
        info = api.parse_wav_header(wav)
        ref = A.parse_wav(wav)
        assert (info.sample_rate, info.channels, info.bits_per_sample, info.sample_kind) == (22050, ch, ref.bits_per_sample, kind)
        assert info.n_frames == ref.samples.size
        assert wav[info.data_offset - 8: info.data_offset - 4] == b"data"


def test_resampler_known_answers():
    assert abs(float(A.bessel_i0(0.0)) - 1.0) < 1e-10                                                             # resampler.rs:469-472
    assert abs(float(A.bessel_i0(1.0)) - 1.2660658777520084) < 1e-10 and abs(float(A.bessel_i0(2.0)) - 2.2795853023360673) < 1e-10
    assert abs(float(A.bessel_i0(3.5)) - float(A.bessel_i0(-3.5))) < 1e-10
    for src, n in ((44100, 44100), (48000, 48000), (8000, 8000)):                                                 # resampler.rs:358-388
        out = A.resample(np.full(n, 0.5, np.float32), src, 16000)
        assert 15900 <= out.size <= 16100
    out = A.resample(np.full(4410, 0.5, np.float32), 44100, 16000)                                                # DC preserved (:391-411)
    mid = out[out.size // 4: out.size // 4 + out.size // 2]
    assert abs(float(mid.mean()) - 0.5) < 1e-4
    t = np.arange(4800, dtype=np.float32)
    sine = np.sin(np.float32(2 * np.pi * 440.0) * t / np.float32(48000.0)).astype(np.float32)
    out = A.resample(sine, 48000, 16000)                                                                          # :414-439
    assert abs(out.size - 1600) <= 2 and np.abs(out).max() > 0.9
    ref = np.sin(2 * np.pi * 440.0 * np.arange(out.size) / 16000.0)
    assert np.abs(out[100:-100] - ref[100:-100]).max() < 2e-3                                                      # a 440 Hz tone survives 3x decimation
    assert np.array_equal(A.resample(sine, 16000, 16000), sine)                                                   # same rate: copy (:341-347)
    assert A.resample(np.array([0.5], np.float32), 44100, 16000).size == 1                                        # single sample (:573-578)
    with pytest.raises(ValueError):
        A.resample(np.zeros(0, np.float32), 44100, 16000)
    # high-frequency rejection (:589-622): a 20 kHz tone at 48 kHz is above the 8 kHz Nyquist of the output
    hf = np.sin(2 * np.pi * 20000.0 * np.arange(4800) / 48000.0).astype(np.float32)
    assert np.abs(A.resample(hf, 48000, 16000)[50:-50]).max() < 0.3


def test_vad_known_answers():
    assert A.frame_energy(np.zeros(480)) < 0.001 and A.zero_crossing_rate(np.zeros(480)) < 0.01                   # vad.rs:934-954
    i = np.arange(480, dtype=np.float32)
    tone = (np.sin(np.float32(2 * np.pi * 440.0) * i / np.float32(16000.0))).astype(np.float32)
    assert A.frame_energy(tone * np.float32(0.5)) > 0.3                                                            # :941-947
    assert 0.04 < A.zero_crossing_rate(tone) < 0.08                                                                # :957-966
    assert A.zero_crossing_rate(np.where(np.arange(480) % 2 == 0, 0.1, -0.1)) > 0.9                                # :969-976
    assert A.zero_crossing_rate(np.array([0.5])) == 0.0                                                            # :1147-1151
    vad = A.VoiceActivityDetector()
    assert vad.detect(np.zeros(16000, np.float32))[0] == []                                                        # :1007-1013
    assert vad.detect(np.zeros(0, np.float32))[0] == [] and vad.detect(np.zeros(100, np.float32))[0] == []         # :1132-1144
    t = np.arange(8000, dtype=np.float32) / np.float32(16000.0)
    speech = (np.sin(np.float32(2 * np.pi * 440.0) * t) * np.float32(0.3)).astype(np.float32)
    audio = np.concatenate([np.zeros(4800, np.float32), speech, np.zeros(4800, np.float32)])
    segs, events = vad.detect(audio)                                                                               # :1016-1045
    assert len(segs) == 1 and 0.2 <= segs[0][0] <= 0.5 and segs[0][1] > segs[0][0] and segs[0][2] > 0.1
    assert events.count(A.EV_START) == 1 and events.count(A.EV_END) == 1
    v = A.VoiceActivityDetector(A.VadConfig(min_speech_frames=1, min_silence_frames=1))                             # :1048-1069
    assert v.process_frame(np.zeros(480, np.float32)) == A.EV_CONTINUE and v.state == A.SILENCE
    assert v.process_frame(tone * np.float32(0.5)) == A.EV_START and v.state == A.SPEECH


def test_chunk_assembler_get_chunk_rules():
    """streaming.rs:843-870: overlap carried into the next chunk, zero pad of a short (flushed) chunk, reset after taking."""
    a = A.ChunkAssembler(chunk_samples=1000, overlap_samples=100)
    x = np.arange(2500, dtype=np.float32)
    a.push(x[:600])
    assert not a.has_chunk() and a.get_chunk() is None
    a.push(x[600:1500])
    c, valid = a.get_chunk()
    assert valid == 1000 and np.array_equal(c, x[:1000])
    assert a.buf.size == 600 and np.array_equal(a.buf[:100], x[900:1000]) and np.array_equal(a.buf[100:], x[1000:1500])
    a.push(x[1500:2500])                                                  # 100 + 500 + 1000 = 1600 held
    c, valid = a.get_chunk()
    assert valid == 1000 and np.array_equal(c, np.concatenate([x[900:1000], x[1000:1900]]))
    c, valid = a.get_chunk(force=True)                                   # flush: [overlap | 600 fresh] zero padded
    assert valid == 700 and np.array_equal(c[:700], np.concatenate([x[1800:1900], x[1900:2500]])) and (c[700:] == 0).all()
    assert a.get_chunk(force=True) is None                               # nothing fresh: flush returns None


# ---- the reference's own WAV fixtures (demos/test-audio/, used by wav.rs:949-988 and tests/cli_parity_tests.rs:28), byte-identical copies
REF_WAVS = ["test-speech-1.5s", "test-8k", "test-24bit", "test-32f", "test-300ms"]


def _ref_wav(name):
    import os
    return open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_wav_{name}.wav"), "rb").read()


def _data_chunk(b):
    """Independent RIFF walk: (offset, size) of the data chunk."""
    i = 12
    while i + 8 <= len(b):
        cid, sz = b[i:i + 4], struct.unpack("<I", b[i + 4:i + 8])[0]
        if cid == b"data":
            return i + 8, sz
        i += 8 + sz + (sz & 1)
    raise AssertionError("no data chunk")


def test_reference_wav_fixtures_parse():
    """parse_wav on the files the reference's own tests read: a LIST chunk sits between fmt and data (the chunk walk), the 16-bit samples
    are i16 / 32768 (convert_16bit_pcm, wav.rs:236-244), and the 24-bit and 32-bit-float renderings of the same recording decode to the
    SAME samples (convert_24bit_pcm / convert_32bit_float, wav.rs:246-286) -- a cross-format known answer."""
    speech = A.parse_wav(_ref_wav("test-speech-1.5s"))
    assert (speech.sample_rate, speech.original_channels, speech.bits_per_sample, speech.samples.size) == (16000, 1, 16, 24000)
    raw = _ref_wav("test-speech-1.5s")
    off, sz = _data_chunk(raw)
    assert off > 44 and sz == 48000                                                  # not the canonical 44-byte header: LIST in between
    assert np.array_equal(speech.samples, np.frombuffer(raw[off:off + sz], "<i2").astype(np.float32) / np.float32(32768.0))
    w24, w32 = A.parse_wav(_ref_wav("test-24bit")), A.parse_wav(_ref_wav("test-32f"))
    assert (w24.bits_per_sample, w32.bits_per_sample) == (24, 32)
    assert np.array_equal(w24.samples, speech.samples) and np.array_equal(w32.samples, speech.samples)
    w8 = A.parse_wav(_ref_wav("test-8k"))
    assert (w8.sample_rate, w8.samples.size) == (8000, 12000)
    up = A.resample(w8.samples, 8000, 16000)
    assert up.size == 24000 and np.isfinite(up).all()
    # the 8 kHz file is the same recording: brought back to 16 kHz it follows the 16 kHz file (band-limited: not equal, but close)
    assert np.corrcoef(up[200:-200], speech.samples[200:-200])[0, 1] > 0.95
    short = A.parse_wav(_ref_wav("test-300ms"))
    assert (short.sample_rate, short.samples.size) == (16000, 4800)


def test_wav_oracle_pinned_on_reference_golden_trace(golden_audio):
    """test_data/ref_a_audio.bin is the REFERENCE's own decode of demos/test-audio/test-speech-1.5s.wav (reference_summary.json:
    step_a_audio, source = that file; the mel goldens start from it): the oracle's parse_wav of the same file must reproduce it bit for
    bit, and the statistics the reference recorded for it."""
    got = A.parse_wav(_ref_wav("test-speech-1.5s")).samples
    assert got.dtype == np.float32 and np.array_equal(got, np.asarray(golden_audio, np.float32))
    # test_data/reference_summary.json, step_a_audio
    assert float(got.min()) == -0.198333740234375 and float(got.max()) == 0.2979736328125
    assert abs(float(got.std()) - 0.06962854415178299) < 1e-9 and int(np.count_nonzero(got)) == 23980
    assert abs(float(got.astype(np.float64).mean()) - 0.00017776997992768884) < 1e-8
