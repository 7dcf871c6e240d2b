"""whisper_apr_b200 -- B200-native (sm_100a) mel + encoder hot path of paiml/whisper.apr.

(The task names the package ``whisper.apr_b200``; a dot cannot appear in an importable Python
package name, so the directory is ``whisper_apr_b200``.)

Only what the path needs lives here: ``csrc/`` (hand-written CUDA kernels + the C ABI of
``libwhisper_b200.so``), the ctypes binding, the host-side mirror of the reference's
``WhisperApr`` / ``MelFilterbank`` / ``Encoder`` interface, an ``.apr`` v1 writer and the
synthetic workload generator.
"""
from ._lib import LIB_PATH, WhisperError, build, lib  # noqa: F401
from .api import (AudioBatch, BatchEncoderOutput, BatchMelResult, BatchPreprocessor, WhisperApr, bf16_bits_to_f32,  # noqa: F401
                  split_into_chunks, to_padded_tensor)

__all__ = ["WhisperApr", "WhisperError", "BatchEncoderOutput", "AudioBatch", "BatchPreprocessor", "BatchMelResult",
           "split_into_chunks", "to_padded_tensor", "bf16_bits_to_f32", "build", "lib", "LIB_PATH"]
