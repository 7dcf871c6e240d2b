"""Host-side mirror of the reference's public API for the mel + encoder path.

Names, argument meaning and error behaviour follow the Rust items they stand in for
(paths relative to the reference checkout); all compute goes through the C ABI of
libwhisper_b200.so -- nothing here computes on the CPU.

  WhisperApr.load_from_apr / config / compute_mel / encode   src/lib.rs:673-754, 330-333, 407-449
  WhisperApr.mel_filters.compute                              src/audio/mel.rs:233-310
  WhisperApr.encoder.forward_mel / forward_batch{,_padded}    src/model/encoder.rs:566-660
  WhisperApr.mel_encode_batch                                 src/lib.rs:1162-1170 (transcribe_batch_optimized steps 1-2)
  split_into_chunks / BatchMelResult.to_padded_tensor         src/audio/batch.rs:219-240, 107-127
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import WB_BF16, WB_F32, WbConfig, WhisperError, check

N_SAMPLES_30S = 480_000
N_FRAMES_30S = 3000
HOP_LENGTH = 160
N_FFT = 400


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _check_out(out, shape, dtype) -> None:
    """A caller-supplied result buffer is written through its raw pointer by the C ABI: refuse anything that is not exactly
    the array the library will fill."""
    if not isinstance(out, np.ndarray):
        raise TypeError("out must be a numpy array")
    if out.dtype != np.dtype(dtype) or tuple(out.shape) != tuple(shape) or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
        raise ValueError(f"out must be a writable C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}, "
                         f"got {out.dtype.name} {tuple(out.shape)}")


@dataclass
class BatchEncoderOutput:
    """model/encoder.rs:673-720."""
    features: np.ndarray          # [batch_size][max_seq_len][d_model], zero padded
    seq_lengths: list
    max_seq_len: int
    batch_size: int
    d_model: int

    def get(self, batch_idx: int):
        if batch_idx >= self.batch_size:
            return None
        return self.features[batch_idx, : self.seq_lengths[batch_idx]].copy()

    def is_empty(self) -> bool:
        return self.batch_size == 0

    def total_tokens(self) -> int:
        return int(sum(self.seq_lengths))


class _MelFilters:
    """MelFilterbank bound to a loaded model (src/audio/mel.rs)."""

    def __init__(self, owner: "WhisperApr"):
        self._o = owner

    @property
    def n_mels(self) -> int:
        return self._o.config.n_mels

    def compute(self, audio, hop_length: int = HOP_LENGTH) -> np.ndarray:
        """MelFilterbank::compute -> [n_frames][n_mels] f32 (frame-major); empty / short audio -> empty."""
        audio = _f32(audio).ravel()
        n = audio.size
        if n == 0:
            return np.zeros((0, self.n_mels), np.float32)
        if hop_length == 0:
            raise WhisperError(_lib.WB_ERR_AUDIO, "hop_length must be positive")
        n_frames = (n - N_FFT) // hop_length + 1 if n >= N_FFT else 0
        out = np.empty((max(n_frames, 0), self.n_mels), np.float32)
        got = C.c_size_t(0)
        check(_lib.lib().wb_mel_compute(self._o._h, _ptr(audio), n, hop_length, _ptr(out), out.size, C.byref(got)))
        return out[: got.value]


class _Encoder:
    """Encoder bound to a loaded model (src/model/encoder.rs)."""

    def __init__(self, owner: "WhisperApr"):
        self._o = owner

    def forward_mel(self, mel) -> np.ndarray:
        """Encoder::forward_mel: mel [n_frames][n_mels] (or flat) -> [S][d] f32."""
        mel = _f32(mel).ravel()
        cfg = self._o.config
        cap = (mel.size // max(cfg.n_mels, 1) // 2 + 2) * cfg.n_audio_state
        out = np.empty(cap, np.float32)
        s = C.c_size_t(0)
        check(_lib.lib().wb_encode(self._o._h, _ptr(mel), mel.size, _ptr(out), out.size, C.byref(s)))
        return out[: s.value * cfg.n_audio_state].reshape(s.value, cfg.n_audio_state).copy()

    def forward_batch_padded(self, batch) -> BatchEncoderOutput:
        """Encoder::forward_batch_padded."""
        cfg = self._o.config
        mels = [_f32(m).ravel() for m in batch]
        B = len(mels)
        if B == 0:
            return BatchEncoderOutput(np.zeros((0, 0, cfg.n_audio_state), np.float32), [], 0, 0, cfg.n_audio_state)
        ptrs = (C.c_void_p * B)(*[m.ctypes.data for m in mels])
        lens = (C.c_size_t * B)(*[m.size for m in mels])
        max_cap = max(m.size // max(cfg.n_mels, 1) // 2 + 2 for m in mels)
        out = np.empty(B * max_cap * cfg.n_audio_state, np.float32)
        seq = (C.c_size_t * B)()
        mx = C.c_size_t(0)
        check(_lib.lib().wb_encode_batch(self._o._h, ptrs, lens, B, _ptr(out), out.size, seq, C.byref(mx)))
        feats = out[: B * mx.value * cfg.n_audio_state].reshape(B, mx.value, cfg.n_audio_state).copy()
        return BatchEncoderOutput(feats, [int(v) for v in seq], int(mx.value), B, cfg.n_audio_state)

    def forward_batch(self, batch):
        """Encoder::forward_batch: list of [S_i][d] arrays."""
        o = self.forward_batch_padded(batch)
        return [o.get(i) for i in range(o.batch_size)]


class WhisperApr:
    """The slice of `WhisperApr` (src/lib.rs:269-449) this library implements."""

    def __init__(self, handle, apr_bytes_keepalive=None):
        self._h = handle
        cfg = WbConfig()
        check(_lib.lib().wb_model_config(self._h, C.byref(cfg)))
        self.config = cfg
        self.mel_filters = _MelFilters(self)
        self.encoder = _Encoder(self)

    # -- construction ------------------------------------------------------------------
    @classmethod
    def load_from_apr(cls, data: bytes, device: int = 0, devices=None) -> "WhisperApr":
        """WhisperApr::load_from_apr.  `devices` (a list of CUDA ordinals) replicates the model over several GPUs of this process:
        the batch entry points then shard their chunks over the list (parallel::configure_thread_pool's successor)."""
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        h = C.c_void_p()
        devs = list(devices) if devices is not None else [device]
        dev_arr = (C.c_int * len(devs))(*devs)
        st = _lib.lib().wb_model_from_apr_devices(C.c_void_p(arr.ctypes.data if arr.size else 0), arr.size, dev_arr, len(devs), C.byref(h))
        check(st)
        return cls(h)

    @property
    def n_devices(self) -> int:
        return int(_lib.lib().wb_model_n_devices(self._h))

    @property
    def devices(self) -> list:
        return [int(_lib.lib().wb_model_device(self._h, i)) for i in range(self.n_devices)]

    def requantize_int8_per_channel(self):
        """quantize_f32_to_i8_per_channel (src/model/quantized.rs:1769-1794) applied on the device to every linear weight."""
        check(_lib.lib().wb_model_requantize(self._h, 1))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().wb_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        check(_lib.lib().wb_model_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_max_batch(self, n: int):
        check(_lib.lib().wb_model_set_max_batch(self._h, n))

    def sync(self):
        check(_lib.lib().wb_sync(self._h))

    # -- reference API -----------------------------------------------------------------
    def compute_mel(self, audio) -> np.ndarray:
        """WhisperApr::compute_mel -> [3000][n_mels] f32."""
        audio = _f32(audio).ravel()
        out = np.empty((N_FRAMES_30S, self.config.n_mels), np.float32)
        check(_lib.lib().wb_compute_mel(self._h, _ptr(audio) if audio.size else None, audio.size, _ptr(out)))
        return out

    def compute_mel_batch(self, audio) -> np.ndarray:
        audio = _f32(audio).reshape(-1, N_SAMPLES_30S)
        out = np.empty((audio.shape[0], N_FRAMES_30S, self.config.n_mels), np.float32)
        check(_lib.lib().wb_compute_mel_batch(self._h, _ptr(audio), audio.shape[0], _ptr(out)))
        return out

    def encode(self, mel) -> np.ndarray:
        """WhisperApr::encode."""
        return self.encoder.forward_mel(mel)

    def mel_encode_batch(self, audio_batch, out_dtype: str = "f32", out: np.ndarray | None = None) -> np.ndarray:
        """transcribe_batch_optimized steps 1-2: list of 1-D f32 arrays -> [B][1500][d]."""
        chunks = [_f32(a).ravel() for a in audio_batch]
        B = len(chunks)
        d = self.config.n_audio_state
        S = (N_FRAMES_30S - 1) // 2 + 1
        want = np.float32 if out_dtype == "f32" else np.uint16              # uint16: raw bf16 bits
        code = WB_F32 if out_dtype == "f32" else WB_BF16
        if out is None:
            out = np.empty((B, S, d), want)
        else:
            _check_out(out, (B, S, d), want)
        if B == 0:
            return out
        ptrs = (C.c_void_p * B)(*[c.ctypes.data for c in chunks])
        lens = (C.c_size_t * B)(*[c.size for c in chunks])
        check(_lib.lib().wb_mel_encode_batch(self._h, ptrs, lens, B, _ptr(out), code))
        return out

    def mel_encode_batch_async(self, audio_batch, out: np.ndarray, out_dtype: str = "f32"):
        """Enqueue form of mel_encode_batch: `audio_batch` (contiguous f32 arrays) and `out` must stay alive until sync()."""
        B = len(audio_batch)
        d = self.config.n_audio_state
        _check_out(out, (B, (N_FRAMES_30S - 1) // 2 + 1, d), np.float32 if out_dtype == "f32" else np.uint16)
        for a in audio_batch:
            if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("mel_encode_batch_async needs C-contiguous float32 arrays (they are read after the call returns)")
        ptrs = (C.c_void_p * B)(*[a.ctypes.data for a in audio_batch])
        lens = (C.c_size_t * B)(*[a.size for a in audio_batch])
        check(_lib.lib().wb_mel_encode_batch_async(self._h, ptrs, lens, B, _ptr(out), WB_F32 if out_dtype == "f32" else WB_BF16))

    def mel_encode_gather(self, audio_batch, gather_index: int = 0, out_dtype: str = "bf16") -> np.ndarray:
        """The sharded call with all encoder states gathered on ONE device (peer stores over NVLink by the final LayerNorm);
        returns them read back from that device: [B][1500][d] f32, or raw bf16 bits as uint16."""
        chunks = [_f32(a).ravel() for a in audio_batch]
        B = len(chunks)
        d = self.config.n_audio_state
        S = (N_FRAMES_30S - 1) // 2 + 1
        out = np.empty((B, S, d), np.float32 if out_dtype == "f32" else np.uint16)
        if B == 0:
            return out
        ptrs = (C.c_void_p * B)(*[c.ctypes.data for c in chunks])
        lens = (C.c_size_t * B)(*[c.size for c in chunks])
        d_states = C.c_void_p()
        check(_lib.lib().wb_mel_encode_gather(self._h, ptrs, lens, B, WB_F32 if out_dtype == "f32" else WB_BF16, gather_index, C.byref(d_states)))
        self.sync()
        check(_lib.lib().wb_read_device(self.devices[gather_index], d_states, _ptr(out), out.nbytes))
        return out

    # -- decoder front half ------------------------------------------------------
    @property
    def has_decoder(self) -> bool:
        return bool(_lib.lib().wb_decoder_available(self._h))

    def decode_greedy(self, states, initial_tokens, max_tokens: int, suppress_timestamps: bool = True):
        """WhisperApr::decode with the greedy strategy for B chunks: states [B][S][d] (or [S][d]) f32 -> list of token lists."""
        st = _f32(states)
        if st.ndim == 2:
            st = st[None]
        B, S, _ = st.shape
        init = np.ascontiguousarray(initial_tokens, np.int32)
        toks = np.empty((B, max_tokens), np.int32)
        lens = np.empty(B, np.int32)
        check(_lib.lib().wb_decode_greedy(self._h, _ptr(st), S, B, _ptr(init), init.size, max_tokens, int(suppress_timestamps), _ptr(toks), _ptr(lens)))
        return [toks[b, : lens[b]].tolist() for b in range(B)]

    def transcribe_tokens_batch(self, audio_batch, initial_tokens, max_tokens: int, suppress_timestamps: bool = True):
        """transcribe_batch_optimized up to token ids: mel + encoder + greedy decode, states never leave HBM."""
        chunks = [_f32(a).ravel() for a in audio_batch]
        B = len(chunks)
        if B == 0:
            return []
        ptrs = (C.c_void_p * B)(*[c.ctypes.data for c in chunks])
        lens_in = (C.c_size_t * B)(*[c.size for c in chunks])
        init = np.ascontiguousarray(initial_tokens, np.int32)
        toks = np.empty((B, max_tokens), np.int32)
        lens = np.empty(B, np.int32)
        check(_lib.lib().wb_transcribe_tokens_batch(self._h, ptrs, lens_in, B, _ptr(init), init.size, max_tokens, int(suppress_timestamps),
                                                     _ptr(toks), _ptr(lens)))
        return [toks[b, : lens[b]].tolist() for b in range(B)]

    def debug_decoder_logits(self, states, tokens) -> np.ndarray:
        st = _f32(states)
        tk = np.ascontiguousarray(tokens, np.int32)
        out = np.empty(self.config.n_vocab, np.float32)
        check(_lib.lib().wb_debug_decoder_logits(self._h, _ptr(st), st.shape[0], _ptr(tk), tk.size, _ptr(out)))
        return out

    def debug_cross_kv(self, states, layer: int):
        st = _f32(states)
        k = np.empty_like(st)
        v = np.empty_like(st)
        check(_lib.lib().wb_debug_cross_kv(self._h, _ptr(st), st.shape[0], layer, _ptr(k), _ptr(v)))
        return k, v

    # -- device-pointer entry points (inputs already in HBM) ------------------------------
    def mel_encode_batch_dev(self, d_audio_ptr: int, B: int, d_out_ptr: int, out_dtype: str = "f32"):
        check(_lib.lib().wb_mel_encode_batch_dev(self._h, C.c_void_p(d_audio_ptr), B, C.c_void_p(d_out_ptr),
                                                  WB_F32 if out_dtype == "f32" else WB_BF16))

    def compute_mel_batch_dev(self, d_audio_ptr: int, B: int, d_mel_ptr: int):
        check(_lib.lib().wb_compute_mel_batch_dev(self._h, C.c_void_p(d_audio_ptr), B, C.c_void_p(d_mel_ptr)))

    def encode_batch_dev(self, d_mel_ptr: int, B: int, d_out_ptr: int, out_dtype: str = "f32"):
        check(_lib.lib().wb_encode_batch_dev(self._h, C.c_void_p(d_mel_ptr), B, C.c_void_p(d_out_ptr),
                                              WB_F32 if out_dtype == "f32" else WB_BF16))

    # -- per-kernel timing (CUDA events on the launching stream) --------------------------
    PROFILE_CATEGORIES = ("mel_stft", "mel_finalize", "gemm", "attention", "layernorm", "other")

    def profile_enable(self, on: bool = True):
        check(_lib.lib().wb_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        n = len(self.PROFILE_CATEGORIES)
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        check(_lib.lib().wb_profile_read(self._h, ms, cnt, n))
        return {k: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, k in enumerate(self.PROFILE_CATEGORIES)}

    # -- test hook ---------------------------------------------------------------------
    def debug_encode(self, mel, n_layers: int = -1, ln_post: bool = True) -> np.ndarray:
        mel = _f32(mel).ravel()
        cfg = self.config
        T = mel.size // cfg.n_mels
        S = (T - 1) // 2 + 1
        out = np.empty((S, cfg.n_audio_state), np.float32)
        check(_lib.lib().wb_debug_encode(self._h, _ptr(mel), mel.size, n_layers, int(ln_post), _ptr(out), out.size))
        return out


class AudioBatch:
    """audio::AudioBatch (src/audio/batch.rs:10-70): an ordered list of segments of any length."""

    def __init__(self, n_mels: int = 80, hop_length: int = HOP_LENGTH):
        self.n_mels, self.hop_length = n_mels, hop_length          # AudioConfig::default (src/audio/mod.rs:30-61)
        self.segments: list[np.ndarray] = []

    def add_segment(self, samples):
        self.segments.append(_f32(samples).ravel().copy())

    def __len__(self):
        return len(self.segments)

    def is_empty(self) -> bool:
        return not self.segments


class BatchMelResult:
    """audio::BatchMelResult (src/audio/batch.rs:72-127)."""

    def __init__(self, mels, frame_counts, max_frames, n_mels):
        self.mels, self.frame_counts, self.max_frames, self.n_mels = mels, frame_counts, max_frames, n_mels

    def __len__(self):
        return len(self.mels)

    def to_padded_tensor(self) -> np.ndarray:
        return to_padded_tensor(self.mels, self.n_mels)


class BatchPreprocessor:
    """audio::BatchPreprocessor (src/audio/batch.rs:129-205): per-segment mel with the preprocessor's own HTK filterbank
    (MelFilterbank::new, batch.rs:143), no 30 s padding.  Runs on the model's device."""

    def __init__(self, model: "WhisperApr", n_mels: int = 80, hop_length: int = HOP_LENGTH):
        self._model, self.n_mels, self.hop_length = model, n_mels, hop_length

    def process_batch(self, batch: AudioBatch) -> BatchMelResult:
        B = len(batch)
        caps = [max(0, (s.size - N_FFT) // self.hop_length + 1 if s.size >= N_FFT else 0) * self.n_mels for s in batch.segments]
        outs = [np.empty(max(c, 1), np.float32) for c in caps]
        counts = (C.c_size_t * max(B, 1))()
        max_frames = C.c_size_t(0)
        if B:
            ptrs = (C.c_void_p * B)(*[s.ctypes.data for s in batch.segments])
            lens = (C.c_size_t * B)(*[s.size for s in batch.segments])
            optrs = (C.c_void_p * B)(*[o.ctypes.data for o in outs])
            ocap = (C.c_size_t * B)(*caps)
            check(_lib.lib().wb_batch_preprocess(self._model._h, ptrs, lens, B, self.n_mels, self.hop_length, optrs, ocap, counts, C.byref(max_frames)))
        mels = [o[: counts[i] * self.n_mels].reshape(counts[i], self.n_mels) for i, o in enumerate(outs)]
        return BatchMelResult(mels, [int(counts[i]) for i in range(B)], int(max_frames.value), self.n_mels)

    @staticmethod
    def normalize_batch(batch: AudioBatch) -> AudioBatch:
        """normalize_audio per segment (batch.rs:179-215): divide by max |x| unless it is below f32 epsilon."""
        out = AudioBatch(batch.n_mels, batch.hop_length)
        for s in batch.segments:
            m = float(np.abs(s).max()) if s.size else 0.0
            out.add_segment(s if m < np.finfo(np.float32).eps else (s / np.float32(m)).astype(np.float32))
        return out


def split_into_chunks(samples, chunk_size: int, overlap: int):
    """audio::split_into_chunks (src/audio/batch.rs:219-240)."""
    samples = _f32(samples).ravel()
    n = _lib.lib().wb_split_into_chunks(samples.size, chunk_size, overlap, None, None, 0)
    if n == 0:
        return []
    starts = (C.c_size_t * n)()
    lens = (C.c_size_t * n)()
    _lib.lib().wb_split_into_chunks(samples.size, chunk_size, overlap, starts, lens, n)
    return [samples[starts[i]: starts[i] + lens[i]].copy() for i in range(n)]


def to_padded_tensor(mels, n_mels: int) -> np.ndarray:
    """BatchMelResult::to_padded_tensor (src/audio/batch.rs:107-127) -> [B][n_mels][max_frames]."""
    ms = [_f32(m).reshape(-1, n_mels) for m in mels]
    B = len(ms)
    mx = max((m.shape[0] for m in ms), default=0)
    out = np.zeros((B, n_mels, mx), np.float32)
    if B == 0 or mx == 0:
        return out
    ptrs = (C.c_void_p * B)(*[m.ctypes.data for m in ms])
    cnt = (C.c_size_t * B)(*[m.shape[0] for m in ms])
    check(_lib.lib().wb_to_padded_tensor(ptrs, cnt, B, n_mels, mx, _ptr(out)))
    return out


def bf16_bits_to_f32(a: np.ndarray) -> np.ndarray:
    """View raw bf16 bit patterns (uint16) as float32 values."""
    return (a.astype(np.uint32) << 16).view(np.float32)


# ------------------------------------------------------------------------------------------------------------------
# Ingest in front of the hot path: WAV, resampling, VAD, chunk views, streaming chunk assembly.
@dataclass
class WavData:
    """audio::wav::WavData (src/audio/wav.rs:60-72)."""
    samples: np.ndarray
    sample_rate: int
    original_channels: int
    bits_per_sample: int


def parse_wav_header(data: bytes) -> "_lib.WbWavInfo":
    """The chunk walk of parse_wav (host only): format, channel count, payload location."""
    buf = np.frombuffer(data, np.uint8)
    info = _lib.WbWavInfo()
    check(_lib.lib().wb_wav_parse(C.c_void_p(buf.ctypes.data if buf.size else 0), buf.size, C.byref(info)))
    return info


def parse_wav(model: WhisperApr, data: bytes) -> WavData:
    """audio::wav::parse_wav (src/audio/wav.rs:99-224): samples converted and down-mixed on the model's device."""
    buf = np.frombuffer(data, np.uint8)
    info = parse_wav_header(data)
    out = np.empty(int(info.n_frames), np.float32)
    check(_lib.lib().wb_wav_decode(model._h, C.c_void_p(buf.ctypes.data), buf.size, _ptr(out), out.size, C.byref(info)))
    return WavData(out, int(info.sample_rate), int(info.channels), int(info.bits_per_sample))


class SincResampler:
    """audio::SincResampler (src/audio/resampler.rs:28-250), resample() on the model's device."""

    def __init__(self, model: WhisperApr, source_rate: int, target_rate: int, kernel_half_len: int = 16, kaiser_beta: float = 6.0):
        if source_rate == 0 or target_rate == 0:
            raise WhisperError(_lib.WB_ERR_AUDIO, "sample rate must be non-zero")
        if kernel_half_len == 0:
            raise WhisperError(_lib.WB_ERR_AUDIO, "kernel half-length must be non-zero")
        self._m, self.source_rate, self.target_rate = model, source_rate, target_rate
        self.kernel_half_len, self.kaiser_beta = kernel_half_len, kaiser_beta

    @property
    def ratio(self) -> float:
        return float(self.target_rate) / float(self.source_rate)

    def resample(self, audio) -> np.ndarray:
        a = _f32(audio).ravel()
        n_out = _lib.lib().wb_resample_len(a.size, self.source_rate, self.target_rate)
        out = np.empty(max(int(n_out), 1), np.float32)
        got = C.c_size_t(0)
        check(_lib.lib().wb_resample_with_params(self._m._h, _ptr(a) if a.size else None, a.size, self.source_rate, self.target_rate,
                                                  self.kernel_half_len, self.kaiser_beta, _ptr(out), out.size, C.byref(got)))
        return out[: got.value]


def ingest_wav_16k(model: WhisperApr, data: bytes):
    """WAV bytes -> mono f32 at 16 kHz, conversion and resampling on the device -> (samples, WavData-like header info)."""
    buf = np.frombuffer(data, np.uint8)
    info = parse_wav_header(data)
    n_out = _lib.lib().wb_resample_len(int(info.n_frames), int(info.sample_rate), 16000) if info.sample_rate else 0
    out = np.empty(max(int(n_out), 1), np.float32)
    got = C.c_size_t(0)
    check(_lib.lib().wb_ingest_wav_16k(model._h, C.c_void_p(buf.ctypes.data), buf.size, _ptr(out), out.size, C.byref(got), C.byref(info)))
    return out[: got.value], info


def vad_detect_batch(model: WhisperApr, audio_batch, config: dict | None = None, seg_capacity: int = 64):
    """VoiceActivityDetector::detect (src/vad.rs:554-607) for a batch of streams -> (segments per stream [(start, end, energy)],
    per-frame events per stream)."""
    streams = [_f32(a).ravel() for a in audio_batch]
    B = len(streams)
    cfg = _lib.WbVadConfig()
    _lib.lib().wb_vad_config_default(C.byref(cfg))
    for k, v in (config or {}).items():
        setattr(cfg, k, v)
    if B == 0:
        return [], []
    ptrs = (C.c_void_p * B)(*[s.ctypes.data for s in streams])
    lens = (C.c_size_t * B)(*[s.size for s in streams])
    segs = np.zeros((B, seg_capacity, 3), np.float32)
    nseg = np.zeros(B, np.int32)
    fs = int(cfg.frame_size)
    events = [np.zeros(s.size // fs + 1, np.uint8) for s in streams]
    eptrs = (C.c_void_p * B)(*[e.ctypes.data for e in events])
    nfr = (C.c_size_t * B)()
    check(_lib.lib().wb_vad_detect_batch(model._h, ptrs, lens, B, C.byref(cfg), _ptr(segs), seg_capacity, _ptr(nseg), eptrs, nfr))
    out_segs = [[tuple(float(x) for x in segs[b, i]) for i in range(min(int(nseg[b]), seg_capacity))] for b in range(B)]
    return out_segs, [events[b][: nfr[b]].tolist() for b in range(B)]


def stream_encode_views(model: WhisperApr, streams, chunk_size: int, overlap: int, out_dtype: str = "f32"):
    """split_into_chunks + compute_mel + encoder for every chunk of every stream, chunks read in place as views of the uploaded
    streams -> (states [total][1500][d], chunk counts per stream)."""
    ss = [_f32(s).ravel() for s in streams]
    n = len(ss)
    d = model.config.n_audio_state
    S = (N_FRAMES_30S - 1) // 2 + 1
    counts = (C.c_size_t * max(n, 1))()
    total = C.c_size_t(0)
    cap = sum(_lib.lib().wb_split_into_chunks(s.size, chunk_size, overlap, None, None, 0) for s in ss)
    out = np.empty((cap, S, d), np.float32 if out_dtype == "f32" else np.uint16)
    if n == 0:
        return out, []
    ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in ss])
    lens = (C.c_size_t * n)(*[s.size for s in ss])
    check(_lib.lib().wb_stream_encode_views(model._h, ptrs, lens, n, chunk_size, overlap, _ptr(out), WB_F32 if out_dtype == "f32" else WB_BF16,
                                             cap, counts, C.byref(total)))
    return out[: total.value], [int(counts[i]) for i in range(n)]


class StreamSet:
    """The chunk-assembly half of StreamingProcessor (src/audio/streaming.rs:672-675, 843-905) for many streams, accumulators in HBM."""

    def __init__(self, model: WhisperApr, n_streams: int, chunk_samples: int, overlap_samples: int):
        self._m, self.n_streams, self.chunk_samples, self.overlap_samples = model, n_streams, chunk_samples, overlap_samples
        h = C.c_void_p()
        check(_lib.lib().wb_stream_set_new(model._h, n_streams, chunk_samples, overlap_samples, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().wb_stream_set_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, stream_ids, samples):
        arrs = [_f32(s).ravel() for s in samples]
        n = len(arrs)
        ids = np.ascontiguousarray(stream_ids, np.int32)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        lens = (C.c_size_t * max(n, 1))(*[a.size for a in arrs])
        check(_lib.lib().wb_stream_set_push(self._h, _ptr(ids), ptrs, lens, n))

    def ready(self) -> list:
        ids = np.zeros(self.n_streams, np.int32)
        n = _lib.lib().wb_stream_set_ready(self._h, _ptr(ids), ids.size)
        return ids[:n].tolist()

    def encode(self, flush: bool = False, out_dtype: str = "f32"):
        """get_chunk (or flush) for every ready stream, assembled on the device, then mel + encoder ->
        (states [n][1500][d], stream ids, valid samples per chunk)."""
        d = self._m.config.n_audio_state
        S = (N_FRAMES_30S - 1) // 2 + 1
        cap = self.n_streams
        out = np.empty((cap, S, d), np.float32 if out_dtype == "f32" else np.uint16)
        ids = np.zeros(cap, np.int32)
        valid = (C.c_size_t * cap)()
        n = C.c_int(0)
        check(_lib.lib().wb_stream_set_encode(self._h, int(flush), _ptr(out), WB_F32 if out_dtype == "f32" else WB_BF16, _ptr(ids), valid, cap, C.byref(n)))
        return out[: n.value], ids[: n.value].tolist(), [int(valid[i]) for i in range(n.value)]

    def debug_chunks(self, n: int) -> np.ndarray:
        out = np.empty((n, self.chunk_samples), np.float32)
        check(_lib.lib().wb_debug_stream_set_chunks(self._h, n, _ptr(out)))
        return out
