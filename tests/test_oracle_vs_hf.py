"""Independent cross-check of the encoder oracle: Hugging Face's WhisperEncoder (PyTorch, CPU, fp64) on the same random weights.

The reference cannot run here and ships no numeric golden for encoder states, so the oracle's encoder restatement
is otherwise pinned only structurally.  HF's encoder is a separate implementation of the same architecture; configured with the tanh
GELU the reference uses (`gelu_new` == encoder.rs:314-318) and fed the reference's tensor names (they ARE the HF names,
src/lib.rs:769-840) it must agree with oracle/encoder.py to float64 round-off.  It says nothing about the mel front end (HF pads and
windows differently), and it is not the reference: DESIGN.md keeps "encoder parity unpinned".
"""
import numpy as np
import pytest

from oracle import encoder as E
from whisper_apr_b200 import synth

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


@pytest.mark.parametrize("name", ["tiny", "base"])
def test_oracle_encoder_matches_hf_whisper_encoder(name):
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    cfg, ecfg = synth.CONFIGS[name], E.CONFIGS[name]
    _, tensors = synth.random_model_apr(cfg, seed=0)
    w = dict(tensors)
    hf_cfg = WhisperConfig(d_model=cfg.n_audio_state, encoder_layers=cfg.n_audio_layer, encoder_attention_heads=cfg.n_audio_head,
                           encoder_ffn_dim=4 * cfg.n_audio_state, num_mel_bins=cfg.n_mels, max_source_positions=1500,
                           activation_function="gelu_new", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                           encoder_layerdrop=0.0, decoder_layers=1, decoder_attention_heads=cfg.n_audio_head, vocab_size=1000)
    hf_cfg._attn_implementation = "eager"
    enc = WhisperEncoder(hf_cfg).double().eval()
    sd = {}
    for k, v in w.items():
        if not k.startswith("encoder."):
            continue
        kk = k[len("encoder."):]
        if kk == "positional_embedding":
            kk = "embed_positions.weight"
        sd[kk] = torch.from_numpy(np.asarray(v, np.float64))
    missing, unexpected = enc.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # the only tensors the synthetic model leaves out are the k_proj biases, which HF does not have either
    assert all(".k_proj.bias" in m for m in missing), missing
    rng = np.random.default_rng(11)
    mel = (0.5 * rng.standard_normal((3000, cfg.n_mels))).astype(np.float32)       # reference layout: [frame][mel]
    F = torch.nn.functional
    with torch.no_grad():
        # HF's forward hard-codes the exact (erf) GELU in the conv stem; the reference uses its tanh GELU there as well
        # (encoder.rs:163-172), so the stem is driven module by module and the blocks / final LayerNorm are HF's own forward code
        x = torch.from_numpy(mel.T[None].astype(np.float64))
        h = F.gelu(enc.conv2(F.gelu(enc.conv1(x), approximate="tanh")), approximate="tanh").permute(0, 2, 1)
        h = h + enc.embed_positions.weight[:1500]
        for layer in enc.layers:
            out = layer(h, None)
            h = out[0] if isinstance(out, tuple) else out
        hf = enc.layer_norm(h)[0].numpy()
        whole = enc(x).last_hidden_state[0].numpy()        # HF end to end (erf GELU in the stem): close, not identical
    ours = E.forward_mel(mel, w, ecfg, attention=E.naive_attention)
    assert hf.shape == ours.shape == (1500, cfg.n_audio_state)
    assert np.abs(hf - ours).max() < 1e-6
    assert np.abs(whole - ours).max() < 1e-3
