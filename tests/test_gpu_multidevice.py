"""Multi-device handle behind the C ABI (-m gpu): one process, several GPUs (wb_model_from_apr_devices), chunk sharding
(chunk i of B -> device floor(i * G / B)), the gather of encoder states by peer stores over NVLink, per-device kernel state.
Tests that need two devices skip on a one-GPU box; the single-device forms of the same entry points always run."""
import numpy as np
import pytest

from oracle import decoder as D
from whisper_apr_b200 import WhisperApr, WhisperError, _lib, bf16_bits_to_f32, synth

pytestmark = pytest.mark.gpu


def _n_gpus():
    return _lib.lib().wb_device_count()


@pytest.fixture(scope="module")
def tiny_apr():
    return synth.random_model_apr(synth.CONFIGS["tiny"], seed=0, with_decoder=True)[0]


def test_gather_on_one_device_equals_host_results(tiny_apr):
    """wb_mel_encode_gather with a single-device handle: the states stay in HBM (library-owned buffer), bit-identical to what the
    host-buffer call returns; micro-batching (max_batch 2 over 5 chunks) does not change them."""
    model = WhisperApr.load_from_apr(tiny_apr)
    assert model.n_devices == 1 and model.devices == [0]
    audio = [synth.synth_audio(50 + i)[: 480000 - 7000 * i] for i in range(5)]
    want = model.mel_encode_batch(audio)
    model.set_max_batch(2)
    got = model.mel_encode_gather(audio, 0, out_dtype="f32")
    assert np.array_equal(got, want)
    got16 = bf16_bits_to_f32(model.mel_encode_gather(audio, 0, out_dtype="bf16"))
    want16 = bf16_bits_to_f32(model.mel_encode_batch(audio, out_dtype="bf16"))
    assert np.array_equal(got16, want16)
    with pytest.raises(WhisperError):
        model.mel_encode_gather(audio, 1)                  # gather index outside the device list
    model.close()


def test_device_list_is_validated(tiny_apr):
    with pytest.raises(WhisperError) as e:
        WhisperApr.load_from_apr(tiny_apr, devices=[0, 0])
    assert "listed twice" in str(e.value)
    with pytest.raises(WhisperError):
        WhisperApr.load_from_apr(tiny_apr, devices=[_n_gpus()])
    with pytest.raises(WhisperError):
        WhisperApr.load_from_apr(tiny_apr, devices=[])


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process_match_single_device_bit_for_bit(tiny_apr):
    """two devices in ONE process.  (a) a second, independent model on device 1 (per-device twiddles, constant table,
    shared-memory opt-ins) gives device 0's results; (b) a two-device handle shards 7 chunks 4 + 3 and returns them in order;
    (c) the gather to either device holds the same bits; (d) the decoder shards the same way."""
    audio = [synth.synth_audio(60 + i)[: 480000 - 9000 * i] for i in range(7)]
    m0 = WhisperApr.load_from_apr(tiny_apr, device=0)
    want = m0.mel_encode_batch(audio)
    want_mel = m0.compute_mel(audio[0])
    want_tok = m0.transcribe_tokens_batch(audio[:4], D.initial_tokens(), 12)
    m1 = WhisperApr.load_from_apr(tiny_apr, device=1)
    assert np.array_equal(m1.compute_mel(audio[0]), want_mel)
    assert np.array_equal(m1.mel_encode_batch(audio), want)
    m1.close()
    both = WhisperApr.load_from_apr(tiny_apr, devices=[0, 1])
    assert both.n_devices == 2 and both.devices == [0, 1]
    both.set_max_batch(2)
    assert np.array_equal(both.mel_encode_batch(audio), want)
    for root in (0, 1):
        assert np.array_equal(both.mel_encode_gather(audio, root, out_dtype="f32"), want)
    g16 = bf16_bits_to_f32(both.mel_encode_gather(audio, 1, out_dtype="bf16"))
    assert np.array_equal(g16, bf16_bits_to_f32(m0.mel_encode_batch(audio, out_dtype="bf16")))
    assert both.transcribe_tokens_batch(audio[:4], D.initial_tokens(), 12) == want_tok
    # fewer chunks than devices: device 1 gets nothing, the result is still complete
    assert np.array_equal(both.mel_encode_batch(audio[:1]), want[:1])
    both.close()
    m0.close()
