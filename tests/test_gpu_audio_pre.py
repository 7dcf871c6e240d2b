"""Ingest stages on the GPU (-m gpu) through the C ABI against the oracle (oracle/audio_pre.py):
WAV payload conversion (bit-exact), sinc resampling (f64 on both sides; <= 1e-6), VAD (events and segments exact), chunk VIEWS of
streams == per-chunk calls (bit-exact), device-side chunk assembly == the oracle's ChunkAssembler (bit-exact), BASELINE configs[4]'s
streaming shape at test size."""
import numpy as np
import pytest

from oracle import audio_pre as A
from oracle import apr_format as F
from oracle import mel as M
from whisper_apr_b200 import WhisperApr, WhisperError, api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny():
    data, _ = synth.random_model_apr(synth.CONFIGS["tiny"], seed=0)
    model = WhisperApr.load_from_apr(data)
    yield model
    model.close()


@pytest.mark.parametrize("kw", [dict(bits=16), dict(bits=8), dict(bits=24), dict(bits=32), dict(float_format=True), dict(bits=16, channels=2),
                                dict(bits=24, channels=2, extensible=True), dict(float_format=True, channels=2), dict(bits=16, extra_chunk=b"abc")])
def test_wav_decode_bit_exact(tiny, kw):
    rng = np.random.default_rng(len(str(kw)))
    x = rng.uniform(-1, 1, 20001 * kw.get("channels", 1) - kw.get("channels", 1) + 1)[: 20000 * kw.get("channels", 1)]
    wav = A.make_wav(x, 44100, **kw)
    ref = A.parse_wav(wav)
    got = api.parse_wav(tiny, wav)
    assert (got.sample_rate, got.original_channels, got.bits_per_sample) == (ref.sample_rate, ref.original_channels, ref.bits_per_sample)
    assert got.samples.shape == ref.samples.shape and np.array_equal(got.samples, ref.samples)


def test_wav_truncated_payload_and_odd_tail(tiny):
    wav = A.make_wav(np.linspace(-1, 1, 101), 16000, bits=16)
    cut = wav[:-3]                                     # data chunk claims more than the file holds; a trailing half sample is dropped
    ref = A.parse_wav(cut)
    got = api.parse_wav(tiny, cut)
    assert got.samples.size == ref.samples.size == 99 and np.array_equal(got.samples, ref.samples)


@pytest.mark.parametrize("src,dst,n", [(48000, 16000, 4800), (44100, 16000, 44100), (8000, 16000, 8000), (16000, 48000, 1600), (22050, 16000, 7), (44100, 16000, 1)])
def test_resample_matches_oracle(tiny, src, dst, n):
    rng = np.random.default_rng(src + n)
    t = np.arange(n) / src
    x = (0.4 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3000 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    got = api.SincResampler(tiny, src, dst).resample(x)
    ref = A.resample(x, src, dst)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-6


def test_resample_rules(tiny):
    x = np.linspace(-1, 1, 100).astype(np.float32)
    assert np.array_equal(api.SincResampler(tiny, 16000, 16000).resample(x), x)                      # resampler.rs:341-347
    with pytest.raises(WhisperError) as e:
        api.SincResampler(tiny, 44100, 16000).resample(np.zeros(0, np.float32))                      # :350-355
    assert e.value.kind == "Audio" and "empty audio" in str(e.value)
    with pytest.raises(WhisperError):
        api.SincResampler(tiny, 0, 16000)                                                            # :310-321
    with pytest.raises(WhisperError):
        api.SincResampler(tiny, 44100, 16000, kernel_half_len=0)                                     # :324-327
    got = api.SincResampler(tiny, 48000, 16000, kernel_half_len=32, kaiser_beta=8.0).resample(np.tile(x, 30))
    assert np.abs(got - A.resample(np.tile(x, 30), 48000, 16000, 32, 8.0)).max() <= 1e-6


def test_ingest_wav_to_16k_and_mel(tiny, fb80):
    """WAV (44.1 kHz stereo 16-bit) -> mono -> 16 kHz on the device, then the mel of the result equals the mel of the oracle's ingest."""
    rng = np.random.default_rng(3)
    n = 44100
    t = np.arange(n) / 44100
    left = 0.3 * np.sin(2 * np.pi * 300 * t) + 0.02 * rng.standard_normal(n)
    right = 0.3 * np.sin(2 * np.pi * 600 * t)
    wav = A.make_wav(np.stack([left, right], 1).ravel(), 44100, bits=16, channels=2)
    got, info = api.ingest_wav_16k(tiny, wav)
    ref = A.resample(A.parse_wav(wav).samples, 44100, 16000)
    assert info.sample_rate == 44100 and info.channels == 2 and got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-6
    assert np.abs(tiny.mel_filters.compute(got) - M.mel_compute(ref, fb80)).max() <= 1e-4


def _speechy(seed, n):
    rng = np.random.default_rng(seed)
    x = np.zeros(n, np.float32)
    pos = 0
    while pos < n:
        quiet = int(rng.integers(2000, 9000))
        loud = int(rng.integers(3000, 12000))
        pos += quiet
        t = np.arange(max(0, min(loud, n - pos)), dtype=np.float32) / np.float32(16000.0)
        x[pos: pos + t.size] = (np.sin(np.float32(2 * np.pi * rng.uniform(300, 900)) * t) * np.float32(rng.uniform(0.1, 0.5))).astype(np.float32)
        pos += loud
    return x + (0.0005 * rng.standard_normal(n)).astype(np.float32)


def test_vad_batch_matches_oracle_exactly(tiny):
    streams = [_speechy(1, 48000), np.zeros(16000, np.float32), _speechy(2, 30300), np.zeros(0, np.float32), _speechy(3, 100), _speechy(4, 16000 + 240)]
    segs, events = api.vad_detect_batch(tiny, streams)
    for s, sg, ev in zip(streams, segs, events):
        rs, rev = A.VoiceActivityDetector().detect(s)
        assert ev == rev
        assert len(sg) == len(rs)
        for a, b in zip(sg, rs):
            assert a == pytest.approx(b, rel=0, abs=0)
    assert segs[1] == [] and segs[3] == [] and len(segs[0]) >= 2
    # other configurations (low latency / high accuracy presets, vad.rs:76-95)
    for cfg in (dict(frame_size=160, min_speech_frames=5, min_silence_frames=15), dict(frame_size=800, min_speech_frames=2, min_silence_frames=6)):
        segs, events = api.vad_detect_batch(tiny, streams[:3], cfg)
        for s, sg, ev in zip(streams[:3], segs, events):
            rs, rev = A.VoiceActivityDetector(A.VadConfig(**cfg)).detect(s)
            assert ev == rev and [tuple(x) for x in sg] == [tuple(np.float32(v) for v in r) for r in rs]


def test_vad_many_streams(tiny):
    streams = [_speechy(100 + i, 16000 + 37 * i) for i in range(300)]
    segs, events = api.vad_detect_batch(tiny, streams)
    for i in (0, 17, 299):
        rs, rev = A.VoiceActivityDetector().detect(streams[i])
        assert events[i] == rev and len(segs[i]) == len(rs)


def test_stream_views_equal_per_chunk_calls(tiny, fb80):
    """split_into_chunks as zero-copy views: same bits as cutting on the host and calling mel_encode_batch on the chunks."""
    streams = [np.concatenate([synth.synth_audio(70), synth.synth_audio(71)])[:200000], synth.synth_audio(72)[:80000], synth.synth_audio(73)[:5000],
               np.zeros(0, np.float32), synth.synth_audio(74)[:152001]]
    out, counts = api.stream_encode_views(tiny, streams, 80000, 8000)
    ref_chunks = [c for s in streams for c in api.split_into_chunks(s, 80000, 8000)]
    assert counts == [3, 1, 1, 0, 3] and out.shape[0] == len(ref_chunks) == 8
    want = tiny.mel_encode_batch(ref_chunks)
    assert np.array_equal(out, want)
    tiny.set_max_batch(3)                                # micro-batching over the chunk table
    out2, _ = api.stream_encode_views(tiny, streams, 80000, 8000)
    tiny.set_max_batch(32)
    assert np.array_equal(out2, want)
    # an odd chunk start (views at 4-byte, not 8-byte, alignment) takes the scalar load path: same bits
    out3, c3 = api.stream_encode_views(tiny, streams[:1], 79999, 8000)
    assert np.array_equal(out3, tiny.mel_encode_batch(api.split_into_chunks(streams[0], 79999, 8000)))
    with pytest.raises(WhisperError):
        api.stream_encode_views(tiny, streams[:1], 500000, 0)


def test_stream_set_chunk_assembly_matches_oracle(tiny):
    """get_chunk / flush for several streams at once (streaming.rs:843-905): the assembled chunks are the oracle's, bit for bit,
    and their encoder states are what the plain batch entry point gives for the same chunks."""
    CH, OV = 80000, 8000
    n = 5
    ss = api.StreamSet(tiny, n, CH, OV)
    refs = [A.ChunkAssembler(CH, OV) for _ in range(n)]
    rng = np.random.default_rng(11)
    audio = [synth.synth_audio(80 + i) for i in range(n)]
    pos = [0] * n
    seen = 0
    for rnd in range(6):
        ids, pieces = [], []
        for i in range(n):
            k = int(rng.integers(10000, 45000)) if i != 3 else 0          # stream 3 never receives audio
            if k:
                ids.append(i); pieces.append(audio[i][pos[i]: pos[i] + k]); refs[i].push(pieces[-1]); pos[i] += k
        ss.push(ids, pieces)
        assert ss.ready() == [i for i in range(n) if refs[i].has_chunk()]
        states, sid, valid = ss.encode()
        want = [(i,) + refs[i].get_chunk() for i in range(n) if refs[i].has_chunk()]
        assert sid == [w[0] for w in want] and valid == [w[2] for w in want]
        if want:
            got_chunks = ss.debug_chunks(len(want))
            for g, w in zip(got_chunks, want):
                assert np.array_equal(g, w[1])
            assert np.array_equal(states, tiny.mel_encode_batch([w[1] for w in want]))
            seen += len(want)
    states, sid, valid = ss.encode(flush=True)                            # end of stream: the remainders, zero padded
    want = [(i,) + r for i in range(n) for r in [refs[i].get_chunk(force=True)] if r is not None]
    assert sid == [w[0] for w in want] and valid == [w[2] for w in want] and 3 not in sid
    for g, w in zip(ss.debug_chunks(len(want)), want):
        assert np.array_equal(g, w[1])
    assert ss.encode(flush=True)[1] == []                                 # nothing fresh is left
    assert seen >= 4
    with pytest.raises(WhisperError) as e:
        ss.push([0], [np.zeros(3 * CH, np.float32)])
    assert "take chunks first" in str(e.value)
    ss.close()


def test_streaming_config4_shape_int4_views(fb80):
    """BASELINE configs[4] at test size: int4 `.apr` payload, 5 s chunks with 500 ms overlap cut from many streams, as views."""
    from oracle import encoder as E
    cfg = synth.CONFIGS["tiny"]
    data, _ = synth.random_model_apr(cfg, quant=F.Q_INT4, seed=0)
    w = F.AprReader(data).load_all()
    model = WhisperApr.load_from_apr(data)
    streams = [synth.synth_audio(90 + i)[: 80000 + 72000 * (i % 3)] for i in range(12)]
    out, counts = api.stream_encode_views(model, streams, 80000, 8000)
    assert counts == [1 + (i % 3) for i in range(12)] and out.shape[0] == sum(counts)
    chunks = [c for s in streams for c in api.split_into_chunks(s, 80000, 8000)]
    for k in (0, 5, out.shape[0] - 1):
        ref = E.forward_mel(M.compute_mel(chunks[k], fb80), w, E.CONFIGS["tiny"], attention=E.naive_attention)
        assert np.abs(out[k] - ref).max() <= 2e-2
    model.close()


# ---- the reference's own WAV fixtures (demos/test-audio/; wav.rs:949-988, tests/cli_parity_tests.rs:28): real speech through every stage
def _ref_wav(name):
    import os
    return open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_wav_{name}.wav"), "rb").read()


@pytest.mark.parametrize("name", ["test-speech-1.5s", "test-8k", "test-24bit", "test-32f", "test-300ms"])
def test_reference_wav_fixtures_decode_bit_exact(tiny, name):
    wav = _ref_wav(name)
    ref = A.parse_wav(wav)
    got = api.parse_wav(tiny, wav)
    assert (got.sample_rate, got.original_channels, got.bits_per_sample) == (ref.sample_rate, ref.original_channels, ref.bits_per_sample)
    assert np.array_equal(got.samples, ref.samples)


def test_device_wav_decode_reproduces_reference_golden_trace(tiny, golden_audio):
    """The reference's own decode of test-speech-1.5s.wav (test_data/ref_a_audio.bin) from the device's WAV path, bit for bit."""
    got = api.parse_wav(tiny, _ref_wav("test-speech-1.5s"))
    assert np.array_equal(got.samples, np.asarray(golden_audio, np.float32))


def test_reference_speech_through_ingest_mel_and_encoder(tiny):
    """Real speech (the reference's 1.5 s sample and its 8 kHz rendering): WAV -> 16 kHz -> log-mel -> encoder against the oracle at the
    path's gates: resampler 1e-6, mel 1e-4, encoder 2e-2 / 0.9999."""
    from oracle import encoder as E
    cfg = synth.CONFIGS["tiny"]
    _, tensors = synth.random_model_apr(cfg, seed=0)
    fb = synth.load_filterbank(cfg.n_mels)
    got16, _ = api.ingest_wav_16k(tiny, _ref_wav("test-8k"))
    ref16 = A.resample(A.parse_wav(_ref_wav("test-8k")).samples, 8000, 16000)
    assert got16.shape == ref16.shape and np.abs(got16 - ref16).max() <= 1e-6
    speech = A.parse_wav(_ref_wav("test-speech-1.5s")).samples
    mel = tiny.compute_mel(speech)
    ref_mel = M.compute_mel(speech, fb)
    assert mel.shape == ref_mel.shape == (3000, cfg.n_mels)
    assert np.abs(mel - ref_mel).max() <= 1e-4
    out = tiny.mel_encode_batch([speech])[0]
    ref = E.forward_mel(ref_mel, dict(tensors), E.CONFIGS["tiny"], attention=E.naive_attention)
    err = float(np.abs(out - ref).max())
    cos = float((out.astype(np.float64) * ref).sum() / (np.linalg.norm(out.astype(np.float64)) * np.linalg.norm(ref)))
    assert err <= 2e-2 and cos >= 0.9999, (err, cos)
