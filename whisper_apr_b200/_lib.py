"""ctypes binding of libwhisper_b200.so (the C ABI declared in include/whisper_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C whisper_apr_b200/csrc``.
There is no fallback: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WB_LIB_PATH") or os.path.join(_HERE, "libwhisper_b200.so")     # WB_LIB_PATH: A/B builds (e.g. the bf16-operand variant)

WB_OK, WB_ERR_AUDIO, WB_ERR_MODEL, WB_ERR_FORMAT, WB_ERR_CUDA = 0, 1, 2, 3, 4
WB_F32, WB_BF16 = 0, 1


class WbConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "model_type", "n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
        "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer", "n_mels", "quantization",
        "has_filterbank", "n_tensors")]


_f32p = C.POINTER(C.c_float)
_szp = C.POINTER(C.c_size_t)
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/whisper_b200.h declares
SIGNATURES = {
    "wb_version": (C.c_char_p, []),
    "wb_operand_format": (C.c_char_p, []),
    "wb_last_error": (C.c_char_p, []),
    "wb_device_count": (C.c_int, []),
    "wb_model_from_apr": (C.c_int, [_vp, C.c_size_t, C.c_int, C.POINTER(_vp)]),
    "wb_model_from_apr_devices": (C.c_int, [_vp, C.c_size_t, C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "wb_model_n_devices": (C.c_int, [_vp]),
    "wb_model_device": (C.c_int, [_vp, C.c_int]),
    "wb_model_requantize": (C.c_int, [_vp, C.c_int]),
    "wb_mel_encode_gather": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "wb_ipc_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_vp), _vp]),
    "wb_ipc_open": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "wb_ipc_close": (C.c_int, [C.c_int, _vp]),
    "wb_ipc_free": (C.c_int, [C.c_int, _vp]),
    "wb_read_device": (C.c_int, [C.c_int, _vp, _vp, C.c_size_t]),
    "wb_decoder_available": (C.c_int, [_vp]),
    "wb_decode_greedy": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "wb_transcribe_tokens_batch": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "wb_debug_decoder_logits": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_int, _vp]),
    "wb_debug_cross_kv": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _vp, _vp]),
    "wb_model_config": (C.c_int, [_vp, C.POINTER(WbConfig)]),
    "wb_model_free": (None, [_vp]),
    "wb_model_set_stream": (C.c_int, [_vp, _vp]),
    "wb_model_set_max_batch": (C.c_int, [_vp, C.c_int]),
    "wb_mel_compute": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, C.c_size_t, _szp]),
    "wb_compute_mel": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "wb_compute_mel_batch": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "wb_encode": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _szp]),
    "wb_encode_batch": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, _vp, C.c_size_t, _szp, _szp]),
    "wb_mel_encode_batch": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, _vp, C.c_int]),
    "wb_batch_preprocess": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, C.c_size_t, C.c_size_t, C.POINTER(_vp), _szp, _szp, _szp]),
    "wb_mel_encode_batch_async": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, _vp, C.c_int]),
    "wb_mel_encode_batch_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int]),
    "wb_compute_mel_batch_dev": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "wb_encode_batch_dev": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int]),
    "wb_sync": (C.c_int, [_vp]),
    "wb_split_into_chunks": (C.c_size_t, [C.c_size_t, C.c_size_t, C.c_size_t, _szp, _szp, C.c_size_t]),
    "wb_to_padded_tensor": (C.c_int, [C.POINTER(_vp), _szp, C.c_int, C.c_size_t, C.c_size_t, _vp]),
    "wb_debug_fft400_power_host": (None, [_vp, _vp]),
    "wb_debug_gemm": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _vp]),
    "wb_debug_attention": (C.c_int, [C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "wb_debug_gemm_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "wb_debug_attention_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "wb_debug_set_ln_follow": (C.c_int, [_vp, C.c_int]),
    "wb_debug_encode": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_int, _vp, C.c_size_t]),
    "wb_launch_count": (C.c_longlong, []),
    "wb_debug_apr_tensor_bytes": (C.c_longlong, [_vp, C.c_size_t, C.c_char_p]),
    "wb_profile_enable": (C.c_int, [_vp, C.c_int]),
    "wb_profile_read": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "wb_debug_layernorm": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
}

class WbWavInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint16), ("bits_per_sample", C.c_uint16), ("sample_kind", C.c_uint16),
                ("reserved", C.c_uint16), ("data_offset", C.c_uint64), ("data_bytes", C.c_uint64), ("n_frames", C.c_uint64)]


class WbVadConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("frame_size", C.c_uint32), ("energy_threshold", C.c_float), ("zcr_threshold", C.c_float),
                ("min_speech_frames", C.c_uint32), ("min_silence_frames", C.c_uint32), ("smoothing", C.c_float)]


SIGNATURES.update({
    "wb_wav_parse": (C.c_int, [_vp, C.c_size_t, C.POINTER(WbWavInfo)]),
    "wb_wav_decode": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.POINTER(WbWavInfo)]),
    "wb_resample_len": (C.c_size_t, [C.c_size_t, C.c_uint32, C.c_uint32]),
    "wb_resample": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, C.c_uint32, _vp, C.c_size_t, _szp]),
    "wb_resample_with_params": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_double, _vp, C.c_size_t, _szp]),
    "wb_ingest_wav_16k": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _szp, C.POINTER(WbWavInfo)]),
    "wb_vad_config_default": (None, [C.POINTER(WbVadConfig)]),
    "wb_vad_detect_batch": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, C.POINTER(WbVadConfig), _vp, C.c_int, _vp, C.POINTER(_vp), _szp]),
    "wb_stream_encode_views": (C.c_int, [_vp, C.POINTER(_vp), _szp, C.c_int, C.c_size_t, C.c_size_t, _vp, C.c_int, C.c_size_t, _szp, _szp]),
    "wb_mel_encode_views_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int]),
    "wb_stream_set_new": (C.c_int, [_vp, C.c_int, C.c_size_t, C.c_size_t, C.POINTER(_vp)]),
    "wb_stream_set_free": (None, [_vp]),
    "wb_stream_set_push": (C.c_int, [_vp, _vp, C.POINTER(_vp), _szp, C.c_int]),
    "wb_stream_set_ready": (C.c_int, [_vp, _vp, C.c_int]),
    "wb_stream_set_encode": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _szp, C.c_int, _vp]),
    "wb_debug_stream_set_chunks": (C.c_int, [_vp, C.c_int, _vp]),
})

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libwhisper_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libwhisper_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C whisper_apr_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class WhisperError(RuntimeError):
    """WhisperError::{Audio, Model, Format} (src/error.rs:6-44) + Cuda."""

    KINDS = {WB_ERR_AUDIO: "Audio", WB_ERR_MODEL: "Model", WB_ERR_FORMAT: "Format", WB_ERR_CUDA: "Cuda"}

    def __init__(self, status: int, message: str):
        self.status = status
        self.kind = self.KINDS.get(status, f"status {status}")
        super().__init__(f"{self.kind} error: {message}")


def check(status: int) -> None:
    if status != WB_OK:
        raise WhisperError(status, (lib().wb_last_error() or b"").decode("utf-8", "replace"))
