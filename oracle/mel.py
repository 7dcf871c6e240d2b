"""Oracle: log-mel spectrogram exactly as whisper.apr computes it (CPU, numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to the reference checkout):
  src/audio/mel.rs:215-219   hann_window  (PERIODIC Hann, f32)
  src/audio/mel.rs:144-212   compute_filterbank (HTK triangles, MelFilterbank::new)
  src/audio/mel.rs:233-310   MelFilterbank::compute
  src/lib.rs:407-443         WhisperApr::compute_mel (pad to 30 s, pad frames with -1.0)
  src/audio/batch.rs:107-127 BatchMelResult::to_padded_tensor
  src/audio/batch.rs:219-240 split_into_chunks

The FFT itself lives in the un-vendored crate rustfft 6.4.1
(FftPlanner::plan_fft_forward(400), call site mel.rs:256-257,279): a forward,
un-normalised DFT with kernel exp(-2*pi*i*k*n/N).  It is restated here with
numpy's pocketfft in float64 ("truth") or float32-rounded inputs/outputs.
"""
from __future__ import annotations

import numpy as np

N_FFT = 400
HOP_LENGTH = 160
N_SAMPLES_30S = 480_000
N_FRAMES_30S = 3000
SAMPLE_RATE = 16_000


def hann_window_periodic(size: int = N_FFT) -> np.ndarray:
    """mel.rs:215-219 -- 0.5*(1-cos(2*pi*n/size)) evaluated in f32."""
    n = np.arange(size, dtype=np.float32)
    two_pi = np.float32(2.0) * np.float32(np.pi)
    ang = (two_pi * n) / np.float32(size)
    return (np.float32(0.5) * (np.float32(1.0) - np.cos(ang, dtype=np.float32))).astype(np.float32)


def hz_to_mel(hz):
    """mel.rs:200-203 (HTK)."""
    return np.float32(2595.0) * np.log10(np.float32(1.0) + np.float32(hz) / np.float32(700.0), dtype=np.float32)


def mel_to_hz(mel):
    """mel.rs:209-212."""
    return np.float32(700.0) * (np.power(np.float32(10.0), np.float32(mel) / np.float32(2595.0), dtype=np.float32) - np.float32(1.0))


def htk_filterbank(n_mels: int, n_fft: int = N_FFT, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """mel.rs:144-197 -- the fallback filterbank of MelFilterbank::new ([n_mels][n_freqs])."""
    n_freqs = n_fft // 2 + 1
    filt = np.zeros((n_mels, n_freqs), dtype=np.float32)
    mel_min = hz_to_mel(0.0)
    mel_max = hz_to_mel(np.float32(sample_rate) / np.float32(2.0))
    bins = []
    for i in range(n_mels + 2):
        m = mel_min + (mel_max - mel_min) * np.float32(i) / np.float32(n_mels + 1)
        f = mel_to_hz(m)
        bins.append(int(np.floor((np.float32(n_fft) + np.float32(1.0)) * f / np.float32(sample_rate))))
    for m in range(n_mels):
        lo, ce, hi = bins[m], bins[m + 1], bins[m + 2]
        for k in range(lo, ce):
            if k < n_freqs and ce > lo:
                filt[m, k] = np.float32(k - lo) / np.float32(ce - lo)
        for k in range(ce, hi):
            if k < n_freqs and hi > ce:
                filt[m, k] = np.float32(hi - k) / np.float32(hi - ce)
    return filt


def n_frames_for(n_samples: int, n_fft: int = N_FFT, hop: int = HOP_LENGTH) -> int:
    """mel.rs:245-249."""
    return (n_samples - n_fft) // hop + 1 if n_samples >= n_fft else 0


def mel_compute(audio, filters, hop_length: int = HOP_LENGTH, *, window=None, precision: str = "f64") -> np.ndarray:
    """MelFilterbank::compute (mel.rs:233-310).

    audio: 1-D f32.  filters: [n_mels][n_freqs] f32.  Returns [n_frames][n_mels]
    f32 (FRAME-major, as the code -- not its doc comment -- stores it, mel.rs:298).
    precision "f64": every sum in float64 (the truth the tolerances are quoted against).
    precision "f32": window product, power spectrum, filterbank accumulation and log10
    in float32, filterbank sum over k ascending as mel.rs:290-295.
    """
    audio = np.asarray(audio, dtype=np.float32).ravel()
    filters = np.asarray(filters, dtype=np.float32)
    n_mels, n_freqs = filters.shape
    n_fft = 2 * (n_freqs - 1)
    if audio.size == 0:
        return np.zeros((0, n_mels), np.float32)
    if hop_length == 0:
        raise ValueError("hop_length must be positive")  # WhisperError::Audio, mel.rs:240-242
    n_frames = n_frames_for(audio.size, n_fft, hop_length)
    if n_frames == 0:
        return np.zeros((0, n_mels), np.float32)
    if window is None:
        window = hann_window_periodic(n_fft)
    window = np.asarray(window)
    idx = (np.arange(n_frames) * hop_length)[:, None] + np.arange(n_fft)[None, :]
    # mel.rs:268-273: samples beyond the end read as 0 (cannot happen with the frame count above)
    frames = audio[np.minimum(idx, audio.size - 1)] * (idx < audio.size)
    if precision == "f64":
        y = frames.astype(np.float64) * window.astype(np.float64)
        spec = np.fft.rfft(y, axis=1)
        power = spec.real ** 2 + spec.imag ** 2
        energy = power @ filters.astype(np.float64).T
        logmel = np.log10(np.maximum(energy, 1e-10))
        g = logmel.max()
        out = (np.maximum(logmel, g - 8.0) + 4.0) / 4.0
        return out.astype(np.float32)
    elif precision == "f32":
        y = (frames.astype(np.float32) * window.astype(np.float32)).astype(np.float32)
        spec = np.fft.rfft(y.astype(np.float64), axis=1)
        re = spec.real.astype(np.float32)
        im = spec.imag.astype(np.float32)
        power = (re * re + im * im).astype(np.float32)
        energy = np.zeros((n_frames, n_mels), np.float32)
        for k in range(n_freqs):  # k ascending, f32 accumulate (mel.rs:290-295)
            col = filters[:, k]
            nz = np.nonzero(col)[0]
            if nz.size:
                energy[:, nz] = energy[:, nz] + power[:, k:k + 1] * col[nz][None, :]
        logmel = np.log10(np.maximum(energy, np.float32(1e-10)), dtype=np.float32)
        g = logmel.max()
        out = np.maximum(logmel, g - np.float32(8.0))
        out = (out + np.float32(4.0)) / np.float32(4.0)
        return out.astype(np.float32)
    raise ValueError(precision)


def compute_mel(audio, filters, *, precision: str = "f64", n_samples: int = N_SAMPLES_30S,
                n_frames: int = N_FRAMES_30S) -> np.ndarray:
    """WhisperApr::compute_mel (lib.rs:407-443): pad/truncate to 30 s, mel, pad frames with -1.0.

    The reference hard-codes N_MELS = 80 (lib.rs:410); here n_mels comes from the
    filterbank so a 128-mel model works (documented generalisation)."""
    audio = np.asarray(audio, dtype=np.float32).ravel()
    padded = np.zeros(n_samples, np.float32)
    k = min(audio.size, n_samples)
    padded[:k] = audio[:k]
    mel = mel_compute(padded, filters, HOP_LENGTH, precision=precision)
    n_mels = filters.shape[0]
    out = np.full((n_frames, n_mels), -1.0, np.float32)
    k = min(mel.shape[0], n_frames)
    out[:k] = mel[:k]
    return out


def to_padded_tensor(mels, n_mels: int) -> np.ndarray:
    """BatchMelResult::to_padded_tensor (batch.rs:107-127): [B][n_mels][max_frames], zero padded."""
    max_frames = max((m.shape[0] for m in mels), default=0)
    out = np.zeros((len(mels), n_mels, max_frames), np.float32)
    for b, m in enumerate(mels):
        out[b, :, : m.shape[0]] = m.T
    return out


def split_into_chunks(samples, chunk_size: int, overlap: int):
    """audio::split_into_chunks (batch.rs:219-240)."""
    samples = np.asarray(samples, dtype=np.float32).ravel()
    if samples.size == 0 or chunk_size == 0:
        return []
    step = max(max(chunk_size - overlap, 0), 1)
    chunks, start = [], 0
    while start < samples.size:
        end = min(start + chunk_size, samples.size)
        chunks.append(samples[start:end].copy())
        start += step
        if end >= samples.size:
            break
    return chunks
