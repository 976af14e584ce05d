"""Synthetic requests shaped like sduss' RunnerRequest (worker/runner/wrappers.py:19-36) for
tests, smoke() and bench.py: random latents / prompt embeddings (no text encoders, no VAE --
prepare and post stages are out of scope), per-request scheduler state."""
from types import SimpleNamespace
from typing import Dict, List

import torch


def make_sd3_requests(cfg, spec: Dict[str, int], steps: int, scheduler, device, ctx_len=333,
                      seed=0, dtype=torch.bfloat16, pin_host=False, latent_dtype=None) -> Dict[str, List]:
    """spec: resolution -> number of requests. Returns dict resolution -> [request].
    latent_dtype: dtype of the running latents (default: `dtype`); the values are drawn in `dtype`
    first, so an fp32 latent starts exactly representable in the model's input dtype."""
    g = torch.Generator().manual_seed(seed)
    reqs, all_reqs, rid = {}, [], 0
    for res, n in spec.items():
        side = int(res) // 8
        reqs[res] = []
        for _ in range(n):
            def rnd(*shape):
                t = torch.randn(*shape, generator=g).to(dtype)
                if pin_host:
                    return t.pin_memory()
                return t.to(device)
            r = SimpleNamespace(
                request_id=rid,
                sampling_params=SimpleNamespace(
                    num_inference_steps=steps, resolution=int(res),
                    latents=rnd(1, cfg.in_channels, side, side).to(latent_dtype or dtype),
                    prompt_embeds=rnd(1, ctx_len, cfg.joint_attention_dim),
                    negative_prompt_embeds=rnd(1, ctx_len, cfg.joint_attention_dim)),
                prepare_output=SimpleNamespace(
                    pooled_prompt_embeds=rnd(1, cfg.pooled_projection_dim),
                    negative_pooled_prompt_embeds=rnd(1, cfg.pooled_projection_dim)),
                scheduler_states=None)
            rid += 1
            reqs[res].append(r)
            all_reqs.append(r)
    scheduler.batch_set_timesteps(all_reqs, device=device)
    return reqs


def make_sdxl_requests(cfg, spec: Dict[str, int], steps: int, scheduler, device, seed=0,
                       dtype=torch.bfloat16, pin_host=False, latent_dtype=None) -> Dict[str, List]:
    g = torch.Generator().manual_seed(seed)
    reqs, all_reqs, rid = {}, [], 0
    for res, n in spec.items():
        side = int(res) // 8
        reqs[res] = []
        for _ in range(n):
            def rnd(*shape):
                t = torch.randn(*shape, generator=g).to(dtype)
                if pin_host:
                    return t.pin_memory()
                return t.to(device)
            ids = torch.tensor([[1024., 1024., 0., 0., 1024., 1024.]]).to(dtype)  # deviation D3
            ids = ids.pin_memory() if pin_host else ids.to(device)
            r = SimpleNamespace(
                request_id=rid,
                sampling_params=SimpleNamespace(
                    num_inference_steps=steps, resolution=int(res),
                    latents=rnd(1, cfg.in_channels, side, side),
                    prompt_embeds=rnd(1, cfg.context_len, cfg.cross_attention_dim),
                    negative_prompt_embeds=rnd(1, cfg.context_len, cfg.cross_attention_dim)),
                prepare_output=SimpleNamespace(
                    pooled_prompt_embeds=rnd(1, cfg.pooled_dim),
                    negative_pooled_prompt_embeds=rnd(1, cfg.pooled_dim),
                    add_time_ids=ids, negative_add_time_ids=ids.clone()),
                scheduler_states=None)
            rid += 1
            reqs[res].append(r)
            all_reqs.append(r)
    scheduler.batch_set_timesteps(all_reqs, device=device)
    # SDXL pipelines scale the initial noise by init_noise_sigma (prepare stage)
    for r in all_reqs:
        r.sampling_params.latents = (r.sampling_params.latents.float() * scheduler.init_noise_sigma).to(dtype) \
            .to(latent_dtype or dtype)
        if pin_host:
            r.sampling_params.latents = r.sampling_params.latents.pin_memory()
    return reqs


# ---------------------------------------------------------------------------------------
# Random-init weights under diffusers' state-dict names, generated on `device` (there is no
# network for checkpoints). Matrices ~ N(0, 1/fan_in) so activations stay O(1).
# ---------------------------------------------------------------------------------------
def _sincos_2d(embed_dim, grid_size, base_size, device):
    coords = torch.arange(grid_size, dtype=torch.float64, device=device) / (grid_size / base_size)
    gw, gh = torch.meshgrid(coords, coords, indexing="xy")

    def one(pos, dim):
        omega = torch.arange(dim // 2, dtype=torch.float64, device=device) / (dim / 2.0)
        out = pos.reshape(-1)[:, None] * (1.0 / 10000 ** omega)[None]
        return torch.cat([torch.sin(out), torch.cos(out)], dim=1)
    return torch.cat([one(gw, embed_dim // 2), one(gh, embed_dim // 2)], dim=1).float()


def random_sd3_state_dict(cfg, device, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    D = cfg.inner_dim
    hd = cfg.attention_head_dim

    def lin(name, fin, fout, gain=1.0):
        sd[name + ".weight"] = (torch.randn(fout, fin, generator=g, device=device) * (gain / fin ** 0.5)).to(dtype)
        sd[name + ".bias"] = (torch.randn(fout, generator=g, device=device) * 0.02).to(dtype)

    def rms(name):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(hd, generator=g, device=device)).to(dtype)

    p = cfg.patch_size
    sd["pos_embed.proj.weight"] = (torch.randn(D, cfg.in_channels, p, p, generator=g, device=device)
                                   / (cfg.in_channels * p * p) ** 0.5).to(dtype)
    sd["pos_embed.proj.bias"] = (torch.randn(D, generator=g, device=device) * 0.02).to(dtype)
    sd["pos_embed.pos_embed"] = _sincos_2d(D, cfg.pos_embed_max_size, cfg.sample_size // p, device).unsqueeze(0)
    lin("time_text_embed.timestep_embedder.linear_1", 256, D)
    lin("time_text_embed.timestep_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", cfg.pooled_projection_dim, D)
    lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("context_embedder", cfg.joint_attention_dim, cfg.caption_projection_dim)
    for i in range(cfg.num_layers):
        b = f"transformer_blocks.{i}"
        dual = i in cfg.dual_attention_layers
        last = i == cfg.num_layers - 1
        lin(b + ".norm1.linear", D, (9 if dual else 6) * D, 0.5)
        lin(b + ".norm1_context.linear", D, (2 if last else 6) * D, 0.5)
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj"):
            lin(f"{b}.attn.{n}", D, D)
        lin(b + ".attn.to_out.0", D, D)
        if not last:
            lin(b + ".attn.to_add_out", D, D)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            rms(f"{b}.attn.{n}")
        if dual:
            for n in ("to_q", "to_k", "to_v"):
                lin(f"{b}.attn2.{n}", D, D)
            lin(b + ".attn2.to_out.0", D, D)
            rms(b + ".attn2.norm_q")
            rms(b + ".attn2.norm_k")
        lin(b + ".ff.net.0.proj", D, 4 * D)
        lin(b + ".ff.net.2", 4 * D, D)
        if not last:
            lin(b + ".ff_context.net.0.proj", D, 4 * D)
            lin(b + ".ff_context.net.2", 4 * D, D)
    lin("norm_out.linear", D, 2 * D, 0.5)
    lin("proj_out", D, p * p * cfg.out_channels)
    return sd


def random_unet_state_dict(cfg, device, seed=0, dtype=torch.bfloat16):
    """SDXL UNet2DConditionModel state dict (diffusers names), random-init on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    T = cfg.block_out_channels[0] * 4
    add_in = cfg.pooled_dim + 6 * cfg.addition_time_embed_dim

    def lin(name, fin, fout, bias=True, gain=1.0):
        sd[name + ".weight"] = (torch.randn(fout, fin, generator=g, device=device) * (gain / fin ** 0.5)).to(dtype)
        if bias:
            sd[name + ".bias"] = (torch.randn(fout, generator=g, device=device) * 0.02).to(dtype)

    def conv(name, cin, cout, k, gain=1.0):
        sd[name + ".weight"] = (torch.randn(cout, cin, k, k, generator=g, device=device)
                                * (gain / (cin * k * k) ** 0.5)).to(dtype)
        sd[name + ".bias"] = (torch.randn(cout, generator=g, device=device) * 0.02).to(dtype)

    def norm(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g, device=device)).to(dtype)
        sd[name + ".bias"] = (0.05 * torch.randn(c, generator=g, device=device)).to(dtype)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        lin(name + ".time_emb_proj", T, cout)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3, 0.5)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    def transformer(name, c, layers):
        norm(name + ".norm", c)
        lin(name + ".proj_in", c, c)
        for j in range(layers):
            b = f"{name}.transformer_blocks.{j}"
            norm(b + ".norm1", c)
            for n in ("to_q", "to_k", "to_v"):
                lin(f"{b}.attn1.{n}", c, c, bias=False)
            lin(b + ".attn1.to_out.0", c, c, gain=0.5)
            norm(b + ".norm2", c)
            lin(b + ".attn2.to_q", c, c, bias=False)
            lin(b + ".attn2.to_k", cfg.cross_attention_dim, c, bias=False)
            lin(b + ".attn2.to_v", cfg.cross_attention_dim, c, bias=False)
            lin(b + ".attn2.to_out.0", c, c, gain=0.5)
            norm(b + ".norm3", c)
            lin(b + ".ff.net.0.proj", c, 8 * c)
            lin(b + ".ff.net.2", 4 * c, c, gain=0.5)
        lin(name + ".proj_out", c, c, gain=0.5)

    ch = cfg.block_out_channels
    conv("conv_in", cfg.in_channels, ch[0], 3)
    lin("time_embedding.linear_1", ch[0], T)
    lin("time_embedding.linear_2", T, T)
    lin("add_embedding.linear_1", add_in, T)
    lin("add_embedding.linear_2", T, T)
    out_c = ch[0]
    for i, c in enumerate(ch):
        in_c, out_c = out_c, c
        for j in range(cfg.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
            if cfg.down_has_attn[i]:
                transformer(f"down_blocks.{i}.attentions.{j}", out_c, cfg.transformer_layers_per_block[i])
        if i != len(ch) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    resnet("mid_block.resnets.0", ch[-1], ch[-1])
    transformer("mid_block.attentions.0", ch[-1], cfg.transformer_layers_per_block[-1])
    resnet("mid_block.resnets.1", ch[-1], ch[-1])
    rev = list(reversed(ch))
    rev_layers = list(reversed(cfg.transformer_layers_per_block))
    rev_attn = list(reversed(cfg.down_has_attn))
    out_c = rev[0]
    for i, c in enumerate(rev):
        prev_out, out_c = out_c, c
        in_c = rev[min(i + 1, len(ch) - 1)]
        for j in range(cfg.layers_per_block + 1):
            skip_c = in_c if j == cfg.layers_per_block else out_c
            resnet(f"up_blocks.{i}.resnets.{j}", (prev_out if j == 0 else out_c) + skip_c, out_c)
            if rev_attn[i]:
                transformer(f"up_blocks.{i}.attentions.{j}", out_c, rev_layers[i])
        if i != len(ch) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    norm("conv_norm_out", ch[0])
    conv("conv_out", ch[0], cfg.out_channels, 3)
    return sd


def random_vae_state_dict(cfg, device, seed=0, dtype=torch.bfloat16):
    """AutoencoderKL decoder-side state dict (diffusers names: post_quant_conv, decoder.*),
    random-init on `device`. cfg: sduss_b200.vae.VAEDecoderConfig."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def lin(name, fin, fout, gain=1.0):
        sd[name + ".weight"] = (torch.randn(fout, fin, generator=g, device=device) * (gain / fin ** 0.5)).to(dtype)
        sd[name + ".bias"] = (torch.randn(fout, generator=g, device=device) * 0.02).to(dtype)

    def conv(name, cin, cout, k, gain=1.0):
        sd[name + ".weight"] = (torch.randn(cout, cin, k, k, generator=g, device=device)
                                * (gain / (cin * k * k) ** 0.5)).to(dtype)
        sd[name + ".bias"] = (torch.randn(cout, generator=g, device=device) * 0.02).to(dtype)

    def norm(name, c):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g, device=device)).to(dtype)
        sd[name + ".bias"] = (0.05 * torch.randn(c, generator=g, device=device)).to(dtype)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3, 0.5)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    ch = list(reversed(cfg.block_out_channels))
    top, C = ch[0], cfg.latent_channels
    if cfg.use_post_quant_conv:
        conv("post_quant_conv", C, C, 1)
    conv("decoder.conv_in", C, top, 3)
    resnet("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v"):
        lin(f"{a}.{n}", top, top)
    lin(a + ".to_out.0", top, top, 0.5)
    resnet("decoder.mid_block.resnets.1", top, top)
    prev = top
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        prev = c
        if i != len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("decoder.conv_norm_out", ch[-1])
    conv("decoder.conv_out", ch[-1], cfg.out_channels, 3)
    return sd


def vae_decode_flops(cfg, h, w):
    """2 x multiply-adds of one VAE decode of an h x w latent (convs, 1x1 shortcuts, attention)."""
    ch = list(reversed(cfg.block_out_channels))
    top, px = ch[0], h * w
    res = lambda cin, cout, p: 2.0 * p * (9 * cin * cout + 9 * cout * cout + (cin * cout if cin != cout else 0))
    fl = 2.0 * 9 * cfg.latent_channels * top * px + 2 * res(top, top, px) + 2.0 * px * 4 * top * top + 4.0 * px * px * top
    prev = top
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block + 1):
            fl += res(prev if j == 0 else c, c, px)
        prev = c
        if i != len(ch) - 1:
            px *= 4
            fl += 2.0 * 9 * c * c * px
    return fl + 2.0 * 9 * ch[-1] * cfg.out_channels * px


# ---------------------------------------------------------------------------------------
# Prepare stage stand-ins: random-init text encoders under transformers' state-dict names (generated
# on the device: T5-XXL is 4.7 B parameters) and a deterministic tokenizer (the real vocabularies are
# not in the image). Shapes are the public model configs of the encoders SD3.5 / SDXL ship with.
# ---------------------------------------------------------------------------------------
CLIP_L = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
              hidden_act="quick_gelu", projection_dim=768, vocab_size=49408, max_position_embeddings=77,
              layer_norm_eps=1e-5, eos_token_id=49407)
CLIP_G = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
              hidden_act="gelu", projection_dim=1280, vocab_size=49408, max_position_embeddings=77,
              layer_norm_eps=1e-5, eos_token_id=49407)
T5_XXL = dict(d_model=4096, d_kv=64, d_ff=10240, num_layers=24, num_heads=64, vocab_size=32128,
              feed_forward_proj="gated-gelu", relative_attention_num_buckets=32,
              relative_attention_max_distance=128, layer_norm_epsilon=1e-6)


def random_clip_state_dict(cfg, device, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device=device).manual_seed(seed)
    D, F, P = cfg["hidden_size"], cfg["intermediate_size"], cfg["projection_dim"]
    sd, p = {}, "text_model."
    rn = lambda *s, std=0.02: (torch.randn(*s, generator=g, device=device) * std).to(dtype)
    sd[p + "embeddings.token_embedding.weight"] = rn(cfg["vocab_size"], D)
    sd[p + "embeddings.position_embedding.weight"] = rn(cfg["max_position_embeddings"], D, std=0.01)
    for i in range(cfg["num_hidden_layers"]):
        b = f"{p}encoder.layers.{i}."
        for n in ("q", "k", "v", "out"):
            sd[f"{b}self_attn.{n}_proj.weight"] = rn(D, D, std=D ** -0.5)
            sd[f"{b}self_attn.{n}_proj.bias"] = rn(D)
        for n in ("layer_norm1", "layer_norm2"):
            sd[f"{b}{n}.weight"] = (1 + 0.1 * torch.randn(D, generator=g, device=device)).to(dtype)
            sd[f"{b}{n}.bias"] = rn(D)
        sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"] = rn(F, D, std=D ** -0.5), rn(F)
        sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"] = rn(D, F, std=F ** -0.5), rn(D)
    sd[p + "final_layer_norm.weight"] = (1 + 0.1 * torch.randn(D, generator=g, device=device)).to(dtype)
    sd[p + "final_layer_norm.bias"] = rn(D)
    sd["text_projection.weight"] = rn(P, D, std=D ** -0.5)
    return sd


def random_t5_state_dict(cfg, device, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device=device).manual_seed(seed)
    D, I, F = cfg["d_model"], cfg["num_heads"] * cfg["d_kv"], cfg["d_ff"]
    rn = lambda *s, std=0.02: (torch.randn(*s, generator=g, device=device) * std).to(dtype)
    sd = {"encoder.embed_tokens.weight": rn(cfg["vocab_size"], D, std=1.0)}
    for i in range(cfg["num_layers"]):
        b = f"encoder.block.{i}.layer."
        for n in "qkv":
            sd[f"{b}0.SelfAttention.{n}.weight"] = rn(I, D, std=(D * (cfg["d_kv"] if n == "q" else 1)) ** -0.5)
        sd[b + "0.SelfAttention.o.weight"] = rn(D, I, std=I ** -0.5)
        sd[b + "0.layer_norm.weight"] = torch.ones(D, device=device, dtype=dtype)
        sd[b + "1.layer_norm.weight"] = torch.ones(D, device=device, dtype=dtype)
        sd[b + "1.DenseReluDense.wi_0.weight"] = rn(F, D, std=D ** -0.5)
        sd[b + "1.DenseReluDense.wi_1.weight"] = rn(F, D, std=D ** -0.5)
        sd[b + "1.DenseReluDense.wo.weight"] = rn(D, F, std=F ** -0.5)
    sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"] = rn(
        cfg["relative_attention_num_buckets"], cfg["num_heads"], std=0.5)
    sd["encoder.final_layer_norm.weight"] = torch.ones(D, device=device, dtype=dtype)
    return sd


class HashTokenizer:
    """Deterministic stand-in with the transformers tokenizer call signature: words hash to ids,
    BOS / EOS framing, EOS (CLIP) or 0 (T5) padding to max_length."""

    def __init__(self, vocab_size, bos=None, eos=1, pad=0):
        self.vocab_size, self.bos, self.eos, self.pad = vocab_size, bos, eos, pad

    def __call__(self, prompts, padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        span = self.vocab_size - 16
        rows = []
        for p in prompts:
            ids = ([self.bos] if self.bos is not None else []) + \
                  [3 + (sum(ord(c) * (i + 1) for i, c in enumerate(w)) % span) for w in p.split()]
            ids = ids[:max_length - 1] + [self.eos]
            rows.append(ids + [self.pad] * (max_length - len(ids)))
        return {"input_ids": torch.tensor(rows)}


def make_prompt_encoder(kind, device, seed=0):
    """B200PromptEncoder with random-init CLIP-L / CLIP-G (/ T5-XXL) + the matching HashTokenizers."""
    from .text_encoders import B200CLIPTextEncoder, B200PromptEncoder, B200T5Encoder
    cl = B200CLIPTextEncoder(random_clip_state_dict(CLIP_L, device, seed), CLIP_L, device=device)
    cg = B200CLIPTextEncoder(random_clip_state_dict(CLIP_G, device, seed + 1), CLIP_G, device=device)
    toks = [HashTokenizer(CLIP_L["vocab_size"], bos=49406, eos=49407, pad=49407),
            HashTokenizer(CLIP_G["vocab_size"], bos=49406, eos=49407, pad=49407)]
    t5 = None
    if kind == "sd3":
        t5 = B200T5Encoder(random_t5_state_dict(T5_XXL, device, seed + 2), T5_XXL, device=device)
        toks.append(HashTokenizer(T5_XXL["vocab_size"], bos=None, eos=1, pad=0))
    return B200PromptEncoder(kind, cl, cg, t5), toks
