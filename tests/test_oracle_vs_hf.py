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


def test_oracle_decoder_matches_hf_whisper_decoder():
    """oracle/decoder.py (forward_one with the KV cache, the decoder the greedy-token gate runs) against HF's WhisperDecoder on the
    same random tensors: logits of the last position of a 6-token prefix, float64."""
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperDecoder
    from oracle import decoder as D
    ecfg = E.ModelConfig("toy", 0, 80, 1500, 128, 2, 2, n_vocab=51865, n_text_ctx=32, n_text_state=128, n_text_head=2, n_text_layer=2)
    w = D.random_decoder_tensors(ecfg, seed=5)
    hf_cfg = WhisperConfig(d_model=128, decoder_layers=2, decoder_attention_heads=2, decoder_ffn_dim=512, encoder_layers=1,
                           encoder_attention_heads=2, vocab_size=51865, max_target_positions=32, activation_function="gelu_new",
                           dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, decoder_layerdrop=0.0, pad_token_id=50256)
    hf_cfg._attn_implementation = "eager"
    dec = WhisperDecoder(hf_cfg).double().eval()
    sd = {k[len("decoder."):]: torch.from_numpy(np.asarray(v, np.float64)) for k, v in w.items()}
    missing, unexpected = dec.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(".k_proj.bias" in m for m in missing), missing
    rng = np.random.default_rng(2)
    enc_states = rng.standard_normal((40, 128))
    toks = [50258, 50259, 50359, 50363, 11, 4242]
    with torch.no_grad():
        h = dec(input_ids=torch.tensor([toks]), encoder_hidden_states=torch.from_numpy(enc_states)[None]).last_hidden_state[0, -1]
        hf_logits = (dec.embed_tokens.weight @ h).numpy()                    # project_to_vocab: tied embedding (decoder.rs:1794-1806)
    od = D.Decoder(w, ecfg, enc_states, dtype=np.float64)
    for t in toks:
        logits = od.forward_one(t)
    assert np.abs(logits - hf_logits).max() < 1e-6        # 2e-8: the reference rounds sqrt(2/pi) to 0.7978846
