"""Prepare-stage text encoders on the B200 kernels (row f-4) against transformers' own
CLIPTextModelWithProjection / T5EncoderModel (the third-party code the reference's encode_prompt
calls; the CPU suite pins oracle/text_encoders.py to the same models) on identical random-init
weights. Tolerance (bf16 kernels vs fp32): cosine >= 0.999 per sequence and max-abs <= 6 % of the
output's max-abs, as for the denoising models."""
import pytest
import torch

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")


def _close(got, ref, what):
    got, ref = got.float().cpu(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    for b in range(ref.shape[0]):
        cos = torch.nn.functional.cosine_similarity(got[b].flatten(), ref[b].flatten(), dim=0).item()
        err = (got[b] - ref[b]).abs().max().item() / ref[b].abs().max().item()
        assert cos >= 0.999 and err <= 0.06, (what, b, cos, err)


def _q(m):
    """Round a transformers model's parameters to bf16-representable values (both sides share them)."""
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            p.copy_(p.to(torch.bfloat16).float())
    return m.eval()


def _clip_model(hidden, heads, layers, inter, act, proj, seed=0):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    torch.manual_seed(seed)
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                         num_attention_heads=heads, max_position_embeddings=77, hidden_act=act,
                         projection_dim=proj, eos_token_id=999, bos_token_id=998, pad_token_id=0)
    return cfg, _q(CLIPTextModelWithProjection(cfg))


def _clip_ids(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 900, (B, 77), generator=g)
    ids[:, 0] = 998
    for b in range(B):
        ids[b, 4 + 23 * b % 70:] = 999
    return ids


@pytest.mark.parametrize("shape", [(128, 2, 3, 512, "quick_gelu", 96), (128, 2, 2, 512, "gelu", 128),
                                   (1280, 20, 2, 5120, "gelu", 1280), (768, 12, 2, 3072, "quick_gelu", 768)])
def test_clip_text_encoder_matches_transformers(cuda, shape):
    from sduss_b200.text_encoders import B200CLIPTextEncoder
    cfg, m = _clip_model(*shape)
    ids = _clip_ids(3)
    with torch.no_grad():
        ref = m(ids, output_hidden_states=True)
    enc = B200CLIPTextEncoder.from_transformers(m, device=cuda)
    for _ in range(2):  # eager, then the captured CUDA graph
        out = enc(ids)
        torch.cuda.synchronize()
        _close(out["hidden_states_penultimate"], ref.hidden_states[-2], "penultimate")
        _close(out["last_hidden_state"], ref.last_hidden_state, "last")
        _close(out["text_embeds"], ref.text_embeds, "text_embeds")


def _t5_model(d_model, heads, d_ff, layers, seed=0):
    from transformers import T5Config, T5EncoderModel
    torch.manual_seed(seed)
    cfg = T5Config(vocab_size=1000, d_model=d_model, d_kv=64, d_ff=d_ff, num_layers=layers, num_heads=heads,
                   feed_forward_proj="gated-gelu", relative_attention_num_buckets=32,
                   relative_attention_max_distance=128)
    m = T5EncoderModel(cfg)
    with torch.no_grad():  # the default init of the bias table is tiny: make it matter
        m.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight.normal_(0, 1.0)
    return cfg, _q(m)


@pytest.mark.parametrize("shape,S", [((128, 2, 256, 3), 200), ((128, 2, 256, 2), 256), ((4096, 64, 10240, 2), 256)])
def test_t5_encoder_matches_transformers(cuda, shape, S):
    from sduss_b200.text_encoders import B200T5Encoder
    cfg, m = _t5_model(*shape)
    ids = torch.randint(1, 900, (2, S), generator=torch.Generator().manual_seed(2))
    ids[0, S // 4:] = 0  # padding is attended (no mask), as diffusers' _get_t5_prompt_embeds does
    with torch.no_grad():
        ref = m(ids)[0]
    enc = B200T5Encoder.from_transformers(m, device=cuda)
    for _ in range(2):
        out = enc(ids)
        torch.cuda.synchronize()
        _close(out, ref, "t5")


def test_prompt_encoder_assembles_like_encode_prompt(cuda):
    """encode_prompt's assembly (SD3: [clip_l | clip_g | 0-pad ; t5] and pooled [emb_l | emb_g]; SDXL:
    [clip_l | clip_g] and pooled emb_g) against the oracle, which is pinned to transformers."""
    from oracle import text_encoders as ote
    from sduss_b200.text_encoders import B200CLIPTextEncoder, B200PromptEncoder, B200T5Encoder
    cl, ml = _clip_model(128, 2, 2, 512, "quick_gelu", 64, seed=1)
    cg, mg = _clip_model(192, 3, 2, 768, "gelu", 96, seed=2)
    ct, mt = _t5_model(512, 8, 1024, 2, seed=3)
    el, eg = B200CLIPTextEncoder.from_transformers(ml, cuda), B200CLIPTextEncoder.from_transformers(mg, cuda)
    et = B200T5Encoder.from_transformers(mt, cuda)
    ids_l, ids_g = _clip_ids(2, 5), _clip_ids(2, 6)
    ids_t = torch.randint(1, 900, (2, 128), generator=torch.Generator().manual_seed(7))
    sd = lambda m: {k: v.detach() for k, v in m.state_dict().items()}
    ref_e, ref_p = ote.sd3_prompt_embeds((sd(ml), cl), (sd(mg), cg), (sd(mt), ct), ids_l, ids_g, ids_t, joint_dim=512)
    got_e, got_p = B200PromptEncoder("sd3", el, eg, et, joint_dim=512).encode(ids_l, ids_g, ids_t)
    torch.cuda.synchronize()
    assert got_e.shape == (2, 77 + 128, 512) and got_p.shape == (2, 64 + 96)
    assert got_e[:, :77, 128 + 192:].abs().max().item() == 0  # zero padding of the CLIP part
    _close(got_e, ref_e, "sd3 prompt_embeds")
    _close(got_p, ref_p, "sd3 pooled")
    ref_e, ref_p = ote.sdxl_prompt_embeds((sd(ml), cl), (sd(mg), cg), ids_l, ids_g)
    got_e, got_p = B200PromptEncoder("sdxl", el, eg).encode(ids_l, ids_g)
    torch.cuda.synchronize()
    _close(got_e, ref_e, "sdxl prompt_embeds")
    _close(got_p, ref_p, "sdxl pooled")


def test_attention_causal_and_relative_bias_variants(cuda):
    """b200_attn_varlen_ex against fp32 softmax attention: causal mask (CLIP) and additive relative
    position bias (T5) on packed sequences of different lengths; the plain path is unchanged."""
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(0)
    H, lens = 3, [77, 200, 256, 130]
    T, D = sum(lens), H * 64
    qkv = (torch.randn(T, 3 * D, generator=g) * 0.7).cuda().bfloat16()
    out = torch.zeros(T, D, device="cuda", dtype=torch.bfloat16)
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    seqs = [(off[i], lens[i], 0, 0, off[i], lens[i], 0, 0) for i in range(len(lens))]
    plan = ops.build_attn_plan(seqs, cuda, H)
    src = ops.attn_source(q=qkv, q_col=0, k=qkv, k_col=D, v=qkv, v_col=2 * D, out=out)
    L = 256
    bias = torch.randn(H, 2 * L - 1, generator=g).cuda()
    for causal, use_bias, scale in ((True, False, 0.125), (False, True, 1.0), (True, True, 0.125), (False, False, 0.125)):
        out.zero_()
        ops.attn_varlen(src, None, *plan, scale, causal=causal,
                        rel_bias=(bias / scale).contiguous() if use_bias else None, rel_len=L)
        torch.cuda.synchronize()
        for i, n in enumerate(lens):
            x = qkv[off[i]:off[i + 1]].float()
            q, k, v = (x[:, j * D:(j + 1) * D].view(n, H, 64).transpose(0, 1) for j in range(3))
            s = q @ k.transpose(-1, -2) * scale
            pos = torch.arange(n, device="cuda")
            if use_bias:
                s = s + bias[:, (pos[None, :] - pos[:, None]) + L - 1]
            if causal:
                s = s.masked_fill(pos[None, :] > pos[:, None], float("-inf"))
            ref = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(n, D)
            err = (out[off[i]:off[i + 1]].float() - ref).abs().max().item()
            assert err < 0.03, (causal, use_bias, i, err)


class _FakeTokenizer:
    """Deterministic stand-in for the CLIP / T5 tokenizers (their vocabulary files are not in the
    image): words hash to ids, BOS / EOS framing and EOS (CLIP) or 0 (T5) padding as the real ones do."""

    def __init__(self, bos=None, eos=999, pad=999):
        self.bos, self.eos, self.pad = bos, eos, pad

    def __call__(self, prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        rows = []
        for p in prompts:
            ids = ([self.bos] if self.bos is not None else []) + \
                  [3 + (sum(ord(c) * (i + 1) for i, c in enumerate(w)) % 890) for w in p.split()]
            ids = ids[:max_length - 1] + [self.eos]
            rows.append(ids + [self.pad] * (max_length - len(ids)))
        return {"input_ids": torch.tensor(rows)}


def test_sd3_prepare_inference_then_step(cuda):
    """prepare_inference (row f-4) hands the requests over exactly as the reference's does: prompt /
    negative prompt embeddings [1, 77 + S, J], pooled [1, P], latents [1, 16, h/8, w/8], scheduler
    state at step 0 -- and the denoising step runs on them. Embeddings are checked against the
    (transformers-pinned) oracle on the same token ids."""
    from types import SimpleNamespace
    from oracle import sd3_mmdit as o3
    from oracle import text_encoders as ote
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel
    from sduss_b200.text_encoders import B200CLIPTextEncoder, B200PromptEncoder, B200T5Encoder
    cfg = o3.sd3_tiny_config()                      # joint dim 128, pooled dim 64, context 45 = 77?? no:
    cfg.context_len = 77 + 32                       # 77 CLIP tokens + 32 T5 tokens in this test
    cl, ml = _clip_model(64, 1, 2, 256, "quick_gelu", 32, seed=1)   # pooled 32 + 32 = 64 = pooled_projection_dim
    cg, mg = _clip_model(64, 1, 2, 256, "gelu", 32, seed=2)         # hidden 64 + 64 = 128 = joint dim (no padding)
    ct, mt = _t5_model(128, 2, 256, 2, seed=3)
    enc = B200PromptEncoder("sd3", B200CLIPTextEncoder.from_transformers(ml, cuda),
                            B200CLIPTextEncoder.from_transformers(mg, cuda),
                            B200T5Encoder.from_transformers(mt, cuda), joint_dim=cfg.joint_attention_dim)
    sd = {k: v.to(torch.bfloat16).float() for k, v in o3.init_sd3_weights(cfg, 0).items()}
    pipe = B200StableDiffusion3Pipeline(B200SD3Transformer2DModel(sd, cfg, device=cuda),
                                        B200FlowMatchEulerDiscreteScheduler())
    toks = [_FakeTokenizer(bos=998), _FakeTokenizer(bos=998), _FakeTokenizer(bos=None, eos=1, pad=0)]
    pipe.attach_text_encoders(enc, toks)
    prompts = ["a photo of a cat", "an oil painting of the sea at dawn", "two dogs"]
    reqs = {}
    for i, (res, p) in enumerate(zip(("256", "512", "256"), prompts)):
        sp = SimpleNamespace(prompt=p, prompt_2=None, prompt_3=None, negative_prompt="blurry" if i == 1 else "",
                             negative_prompt_2=None, negative_prompt_3=None, num_inference_steps=28,
                             height=int(res), width=int(res), latents=None)
        reqs.setdefault(res, []).append(SimpleNamespace(request_id=i, sampling_params=sp))
    pipe.prepare_inference(reqs, guidance_scale=7.0, generator=torch.Generator(device="cuda").manual_seed(0),
                           max_sequence_length=32)
    torch.cuda.synchronize()
    flat = [r for res in sorted(reqs) for r in reqs[res]]
    sdd = lambda m: {k: v.detach() for k, v in m.state_dict().items()}
    for branch, attr, pattr in (("prompt", "prompt_embeds", "pooled_prompt_embeds"),
                                ("negative_prompt", "negative_prompt_embeds", "negative_pooled_prompt_embeds")):
        texts = [getattr(r.sampling_params, branch) or "" for r in flat]
        ids = [t(texts, max_length=n)["input_ids"] for t, n in zip(toks, (77, 77, 32))]
        ref_e, ref_p = ote.sd3_prompt_embeds((sdd(ml), cl), (sdd(mg), cg), (sdd(mt), ct), *ids,
                                             joint_dim=cfg.joint_attention_dim)
        got_e = torch.cat([getattr(r.sampling_params, attr) for r in flat])
        got_p = torch.cat([getattr(r.prepare_output, pattr) for r in flat])
        assert got_e.shape == (3, 77 + 32, cfg.joint_attention_dim) and got_p.shape == (3, cfg.pooled_projection_dim)
        _close(got_e, ref_e, attr)
        _close(got_p, ref_p, pattr)
    for r in flat:
        s = r.sampling_params.height // 8
        assert r.sampling_params.latents.shape == (1, 16, s, s) and r.sampling_params.latents.is_cuda
        assert r.scheduler_states._step_index == 0 and len(r.scheduler_states.sigmas) == 29
    before = [r.sampling_params.latents.clone() for r in flat]
    pipe.denoising_step(reqs, True, 7.0, True, 256)
    torch.cuda.synchronize()
    for r, x in zip(flat, before):
        assert r.scheduler_states._step_index == 1
        assert torch.isfinite(r.sampling_params.latents.float()).all() and not torch.equal(r.sampling_params.latents, x)
