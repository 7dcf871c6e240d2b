"""Oracle: the reference's audio ingest stages in front of the mel (CPU, numpy).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to the reference checkout):
  src/audio/wav.rs:99-293        parse_wav (RIFF walk, fmt / WAVE_FORMAT_EXTENSIBLE, data), convert_{8,16,24,32}bit_pcm,
                                 convert_32bit_float, convert_to_mono
  src/audio/resampler.rs:66-250  SincResampler::{new, with_params, resample, windowed_sinc, kaiser_window}, bessel_i0 (:260-276)
  src/vad.rs:36-66, 501-700      VadConfig, VoiceActivityDetector::{detect, process_frame, is_speech_frame, frame_energy,
                                 zero_crossing_rate}
  src/audio/streaming.rs:843-870 StreamingProcessor::get_chunk (overlap carry + zero pad to the chunk size)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT, WAVE_FORMAT_EXTENSIBLE = 1, 3, 0xFFFE


class WavError(ValueError):
    pass


@dataclass
class WavData:
    samples: np.ndarray
    sample_rate: int
    original_channels: int
    bits_per_sample: int


def parse_wav(data: bytes) -> WavData:
    """wav.rs:99-224."""
    if len(data) < 44:
        raise WavError("WAV file too small")
    if data[0:4] != b"RIFF":
        raise WavError("missing RIFF header")
    if data[8:12] != b"WAVE":
        raise WavError("missing WAVE format")
    pos = 12
    sample_rate = channels = bits = audio_format = sub_format = 0
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        size = struct.unpack_from("<I", data, pos + 4)[0]
        if cid == b"fmt ":
            if pos + 8 + size > len(data):
                raise WavError("fmt chunk truncated")
            audio_format, channels, sample_rate = struct.unpack_from("<HHI", data, pos + 8)
            bits = struct.unpack_from("<H", data, pos + 22)[0]
            if audio_format == WAVE_FORMAT_EXTENSIBLE and size >= 40:
                off = pos + 8 + 24
                if off + 2 <= len(data):
                    sub_format = struct.unpack_from("<H", data, off)[0]
            pos += 8 + size
        elif cid == b"data":
            start = pos + 8
            end = min(start + size, len(data))
            raw = data[start:end]
            eff = sub_format if audio_format == WAVE_FORMAT_EXTENSIBLE else audio_format
            if eff not in (WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT):
                raise WavError(f"unsupported format {audio_format} with {bits} bits")
            if (eff, bits) == (WAVE_FORMAT_PCM, 16):
                n = len(raw) // 2
                s = np.frombuffer(raw[:2 * n], "<i2").astype(np.float32) / np.float32(32768.0)
            elif (eff, bits) == (WAVE_FORMAT_PCM, 8):
                s = (np.frombuffer(raw, np.uint8).astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
            elif (eff, bits) == (WAVE_FORMAT_PCM, 24):
                n = len(raw) // 3
                b = np.frombuffer(raw[:3 * n], np.uint8).reshape(n, 3).astype(np.int32)
                v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
                v = np.where(v & 0x800000, v - (1 << 24), v)
                s = v.astype(np.float32) / np.float32(8388608.0)
            elif (eff, bits) == (WAVE_FORMAT_PCM, 32):
                n = len(raw) // 4
                s = np.frombuffer(raw[:4 * n], "<i4").astype(np.float32) / np.float32(2147483648.0)
            elif (eff, bits) == (WAVE_FORMAT_IEEE_FLOAT, 32):
                n = len(raw) // 4
                s = np.frombuffer(raw[:4 * n], "<f4").astype(np.float32)
            else:
                raise WavError(f"unsupported format {audio_format} with {bits} bits")
            if channels == 1:
                mono = s
            elif channels == 2:
                n = s.size // 2
                mono = ((s[0:2 * n:2] + s[1:2 * n:2]) / np.float32(2.0)).astype(np.float32)
            else:
                raise WavError(f"unsupported channel count {channels}")
            return WavData(np.ascontiguousarray(mono, np.float32), sample_rate, channels, bits)
        else:
            pos += 8 + size
            if size % 2 != 0:
                pos += 1
    raise WavError("no data chunk")


def make_wav(samples, sample_rate: int, bits: int = 16, channels: int = 1, float_format: bool = False, extensible: bool = False,
             extra_chunk: bytes | None = None) -> bytes:
    """A WAV writer for fixtures (interleaved input for 2 channels)."""
    x = np.asarray(samples, np.float64)
    if float_format:
        raw = x.astype("<f4").tobytes()
        fmt_code, bits = WAVE_FORMAT_IEEE_FLOAT, 32
    elif bits == 16:
        raw = np.clip(np.round(x * 32767.0), -32768, 32767).astype("<i2").tobytes()
        fmt_code = WAVE_FORMAT_PCM
    elif bits == 8:
        raw = np.clip(np.round(x * 127.0 + 128.0), 0, 255).astype(np.uint8).tobytes()
        fmt_code = WAVE_FORMAT_PCM
    elif bits == 24:
        v = np.clip(np.round(x * 8388607.0), -8388608, 8388607).astype(np.int64) & 0xFFFFFF
        raw = b"".join(int(t).to_bytes(3, "little") for t in v)
        fmt_code = WAVE_FORMAT_PCM
    elif bits == 32:
        raw = np.clip(np.round(x * 2147483647.0), -2147483648, 2147483647).astype("<i4").tobytes()
        fmt_code = WAVE_FORMAT_PCM
    else:
        raise ValueError(bits)
    block = channels * bits // 8
    if extensible:
        fmt = struct.pack("<HHIIHH", WAVE_FORMAT_EXTENSIBLE, channels, sample_rate, sample_rate * block, block, bits)
        fmt += struct.pack("<HHI", 22, bits, 0) + struct.pack("<H", fmt_code) + bytes(14)
    else:
        fmt = struct.pack("<HHIIHH", fmt_code, channels, sample_rate, sample_rate * block, block, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if extra_chunk is not None:
        body += b"LIST" + struct.pack("<I", len(extra_chunk)) + extra_chunk + (b"\0" if len(extra_chunk) % 2 else b"")
    body += b"data" + struct.pack("<I", len(raw)) + raw
    return b"RIFF" + struct.pack("<I", len(body)) + body


# ----------------------------------------------------------------------------------------------------------- resampler
def bessel_i0(x: np.ndarray) -> np.ndarray:
    """resampler.rs:260-276: series sum (x^2/4)^k / (k!)^2, k < 50, stop when the term is below 1e-15 of the sum."""
    x = np.asarray(x, np.float64)
    s = np.ones_like(x)
    term = np.ones_like(x)
    q = (x * x) / 4.0
    live = np.ones(x.shape, bool)
    for k in range(1, 50):
        term = np.where(live, term * (q / float(k * k)), term)
        s = np.where(live, s + term, s)
        live &= ~(np.abs(term) < 1e-15 * np.abs(s))
    return s


def resample(audio, source_rate: int, target_rate: int, kernel_half_len: int = 16, kaiser_beta: float = 6.0) -> np.ndarray:
    """SincResampler::resample (resampler.rs:136-206): Kaiser-windowed sinc interpolation, f64 accumulation, weight-normalised.
    (kaiser_window's x.mul_add(-x, 1.0) is restated unfused: a <= 1 ulp f64 difference, invisible after the f32 cast.)"""
    audio = np.asarray(audio, np.float32)
    if source_rate == 0 or target_rate == 0:
        raise ValueError("sample rate must be non-zero")
    if kernel_half_len == 0:
        raise ValueError("kernel half-length must be non-zero")
    if audio.size == 0:
        raise ValueError("cannot resample empty audio")
    if source_rate == target_rate:
        return audio.copy()
    ratio = float(target_rate) / float(source_rate)
    n_out = int(np.ceil(audio.size * ratio))
    if n_out == 0:
        raise ValueError("output length would be zero")
    cutoff = ratio if ratio < 1.0 else 1.0
    in_pos = np.arange(n_out, dtype=np.float64) / ratio
    center = np.floor(in_pos).astype(np.int64)
    frac = in_pos - np.floor(in_pos)
    total = np.zeros(n_out, np.float64)
    wsum = np.zeros(n_out, np.float64)
    a64 = audio.astype(np.float64)
    i0b = bessel_i0(np.float64(kaiser_beta))
    for k in range(-kernel_half_len, kernel_half_len + 1):
        idx = center + k
        ok = (idx >= 0) & (idx < audio.size)
        x = float(k) - frac
        sarg = cutoff * x
        with np.errstate(invalid="ignore", divide="ignore"):
            sinc = np.where(np.abs(sarg) < 1e-10, 1.0, np.sin(np.pi * sarg) / (np.pi * sarg))
        warg = x / float(kernel_half_len)
        win = np.where(np.abs(warg) > 1.0, 0.0, bessel_i0(kaiser_beta * np.sqrt(np.maximum(1.0 - warg * warg, 0.0))) / i0b)
        v = sinc * win
        total = np.where(ok, total + a64[np.clip(idx, 0, audio.size - 1)] * v, total)
        wsum = np.where(ok, wsum + v, wsum)
    out = np.zeros(n_out, np.float32)
    good = np.abs(wsum) > 1e-10
    out[good] = (total[good] / wsum[good]).astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------------------------- VAD
@dataclass
class VadConfig:
    """vad.rs:36-66."""
    sample_rate: int = 16000
    frame_size: int = 480
    energy_threshold: float = 2.0
    zcr_threshold: float = 0.3
    min_speech_frames: int = 3
    min_silence_frames: int = 10
    smoothing: float = 0.95


SILENCE, SPEECH = 0, 1
EV_CONTINUE, EV_START, EV_END = 0, 1, 2


def frame_energy(frame: np.ndarray) -> np.float32:
    """vad.rs:671-674: sequential f32 sum of squares, sqrt(sum / len)."""
    f = np.asarray(frame, np.float32)
    s = np.cumsum(f * f, dtype=np.float32)[-1] if f.size else np.float32(0)
    return np.sqrt(np.float32(s) / np.float32(f.size))


def zero_crossing_rate(frame: np.ndarray) -> np.float32:
    """vad.rs:677-688."""
    f = np.asarray(frame, np.float32)
    if f.size < 2:
        return np.float32(0)
    pos = f >= 0
    return np.float32(np.float32(np.count_nonzero(pos[:-1] != pos[1:])) / np.float32(f.size - 1))


@dataclass
class VoiceActivityDetector:
    config: VadConfig = field(default_factory=VadConfig)
    noise_floor: np.float32 = np.float32(0.001)
    state: int = SILENCE
    speech_frames: int = 0
    silence_frames: int = 0
    current_sample: int = 0

    def process_frame(self, frame) -> int:
        """vad.rs:609-660."""
        c = self.config
        energy, zcr = frame_energy(frame), zero_crossing_rate(frame)
        if self.state == SILENCE:
            sm = np.float32(c.smoothing)
            # f32::mul_add: one rounding
            self.noise_floor = np.float32(np.float64(sm) * np.float64(self.noise_floor) + np.float64((np.float32(1.0) - sm) * energy))
        is_speech = bool(energy > self.noise_floor * np.float32(c.energy_threshold)) and bool(zcr > np.float32(0.05)) and bool(zcr < np.float32(c.zcr_threshold))
        if self.state == SILENCE:
            if is_speech:
                self.speech_frames += 1
                self.silence_frames = 0
                if self.speech_frames >= c.min_speech_frames:
                    self.state = SPEECH
                    return EV_START
                return EV_CONTINUE
            self.speech_frames = 0
            self.state = SILENCE
            return EV_CONTINUE
        if is_speech:
            self.silence_frames = 0
            self.speech_frames += 1
            return EV_CONTINUE
        self.silence_frames += 1
        self.speech_frames = 0
        if self.silence_frames >= c.min_silence_frames:
            self.state = SILENCE
            return EV_END
        return EV_CONTINUE

    def detect(self, audio):
        """vad.rs:554-607 -> (segments [(start, end, energy)], per-frame events)."""
        self.__init__(self.config)
        audio = np.asarray(audio, np.float32)
        fs = self.config.frame_size
        segments, events = [], []
        cur = None
        count = 0
        sr = np.float32(self.config.sample_rate)
        for s in range(0, audio.size, fs):
            frame = audio[s:s + fs]
            if frame.size < fs // 2:
                break
            ev = self.process_frame(frame)
            events.append(ev)
            time = np.float32(np.float32(self.current_sample) / sr)
            if ev == EV_START:
                cur = [time, frame_energy(frame)]
                count = 1
            elif ev == EV_END:
                if cur is not None:
                    segments.append((float(cur[0]), float(time), float(np.float32(cur[1] / np.float32(max(count, 1))))))
                    cur = None
            elif cur is not None:
                cur[1] = np.float32(cur[1] + frame_energy(frame))
                count += 1
            self.current_sample += frame.size
        if cur is not None:
            time = np.float32(np.float32(self.current_sample) / sr)
            segments.append((float(cur[0]), float(time), float(np.float32(cur[1] / np.float32(max(count, 1))))))
        return segments, events


# --------------------------------------------------------------------------------------------------------- streaming
class ChunkAssembler:
    """The chunk-assembly half of StreamingProcessor with VAD gating off (streaming.rs:843-870 get_chunk, :872-905 flush): audio
    accumulates behind the carried overlap; a chunk is ready at chunk_samples; taking it keeps the last overlap_samples as the
    prefix of the next one and zero-pads a short (flushed) chunk to chunk_samples."""

    def __init__(self, chunk_samples: int, overlap_samples: int):
        self.chunk_samples, self.overlap_samples = chunk_samples, overlap_samples
        self.buf = np.zeros(0, np.float32)
        self.overlap = np.zeros(0, np.float32)
        self.fresh = 0                       # samples pushed since the last chunk was taken

    def push(self, samples):
        s = np.asarray(samples, np.float32)
        if self.buf.size == 0 and s.size:
            self.buf = self.overlap.copy()   # streaming.rs:749-751: a new chunk starts with the carried overlap
        self.buf = np.concatenate([self.buf, s])
        self.fresh += s.size

    def has_chunk(self) -> bool:
        return self.buf.size >= self.chunk_samples

    def get_chunk(self, force: bool = False):
        if not (self.has_chunk() or (force and self.fresh > 0)):
            return None
        take = self.buf[: self.chunk_samples] if self.buf.size >= self.chunk_samples else self.buf
        rest = self.buf[take.size:]
        if take.size > self.overlap_samples:
            self.overlap = take[take.size - self.overlap_samples:].copy()
        chunk = np.zeros(self.chunk_samples, np.float32)
        chunk[: take.size] = take
        valid = take.size
        self.buf = np.zeros(0, np.float32)
        self.fresh = 0
        if rest.size:
            self.push(rest)
        return chunk, valid
