"""The prepare-stage oracle (oracle/text_encoders.py) pinned against the third-party implementation
the reference actually calls: transformers' CLIPTextModelWithProjection and T5EncoderModel, on
identical random-init weights (no checkpoints, no tokenizer files: ids are synthetic)."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _clip(hidden_act, with_eos_id):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    torch.manual_seed(0)
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3,
                         num_attention_heads=2, max_position_embeddings=77, hidden_act=hidden_act,
                         projection_dim=96, eos_token_id=999 if with_eos_id else 2, bos_token_id=998,
                         pad_token_id=0)
    m = CLIPTextModelWithProjection(cfg).eval()
    with torch.no_grad():  # default init leaves the biases zero: make every term count
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
    return cfg, m


@pytest.mark.parametrize("hidden_act,with_eos_id", [("quick_gelu", True), ("gelu", True), ("quick_gelu", False)])
def test_clip_oracle_matches_transformers(hidden_act, with_eos_id):
    from oracle import text_encoders as ote
    cfg, m = _clip(hidden_act, with_eos_id)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(3, 900, (3, 77), generator=g)
    ids[:, 0] = 998
    for b, n in enumerate((5, 40, 76)):
        ids[b, n:] = 999  # EOS then EOS padding (CLIP pads with the EOS token)
    with torch.no_grad():
        ref = m(ids, output_hidden_states=True)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    hs, last, pooled, emb = ote.clip_text_forward(sd, cfg, ids)
    assert len(hs) == len(ref.hidden_states)
    for a, b in zip(hs, ref.hidden_states):
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-4)
    assert torch.allclose(last, ref.last_hidden_state, atol=2e-5, rtol=1e-4)
    assert torch.allclose(emb, ref.text_embeds, atol=2e-5, rtol=1e-4)


def test_t5_oracle_matches_transformers():
    from transformers import T5Config, T5EncoderModel
    from oracle import text_encoders as ote
    torch.manual_seed(0)
    cfg = T5Config(vocab_size=1000, d_model=128, d_kv=64, d_ff=256, num_layers=3, num_heads=2,
                   feed_forward_proj="gated-gelu", relative_attention_num_buckets=32,
                   relative_attention_max_distance=128)
    m = T5EncoderModel(cfg).eval()
    ids = torch.randint(1, 900, (2, 200), generator=torch.Generator().manual_seed(2))
    ids[0, 50:] = 0  # padding is attended (no mask), as in diffusers' _get_t5_prompt_embeds
    with torch.no_grad():
        ref = m(ids)[0]
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    out = ote.t5_encoder_forward(sd, cfg, ids)
    assert torch.allclose(out, ref, atol=3e-5, rtol=1e-4)
    # bucket table itself
    from transformers.models.t5.modeling_t5 import T5Attention
    ctx, mem = torch.arange(300)[:, None], torch.arange(300)[None, :]
    want = T5Attention._relative_position_bucket(mem - ctx, bidirectional=True, num_buckets=32, max_distance=128)
    assert torch.equal(ote.t5_relative_buckets(300, 32, 128), want)
