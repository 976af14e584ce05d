"""ORACLE (test infrastructure, not product code): CPU/PyTorch restatement of the text encoders the
reference's PREPARE stage runs (second half):

  ESyMReDStableDiffusion3Pipeline.prepare_inference -> diffusers' encode_prompt
      sduss/model_executor/diffusers/pipelines/stable_diffusion_3/pipeline_stable_diffusion_3_esymred.py:49-230
      CLIP-L + CLIP-G (CLIPTextModelWithProjection) + T5-XXL encoder (T5EncoderModel)
  ESyMReDStableDiffusionXLPipeline.prepare_inference -> encode_prompt
      pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl_esymred.py:56-258
      CLIP-L (CLIPTextModel) + CLIP-G (CLIPTextModelWithProjection)

The encoder arithmetic lives in the third-party package `transformers` (the reference pins
transformers==4.47.1, conda.yml:163; this image has 5.5.0). Unlike diffusers it IS importable here,
so this restatement is PINNED: tests/test_text_encoders_cpu.py checks it against transformers' own
CLIPTextModelWithProjection / T5EncoderModel on identical random-init weights. State-dict names are
transformers' names, so real checkpoints load unchanged.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math
from typing import Dict

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ CLIP text transformer
def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def clip_text_forward(sd: Dict[str, torch.Tensor], cfg, ids: torch.Tensor):
    """CLIPTextTransformer (pre-LN, causal mask, learned positions). cfg: hidden_size,
    num_hidden_layers, num_attention_heads, hidden_act, layer_norm_eps, eos_token_id.
    Returns (hidden_states list incl. embeddings [L+1], last_hidden after final LN, pooled =
    final-LN hidden at the EOS token, text_embeds = text_projection(pooled) or None)."""
    p = "text_model."
    B, S = ids.shape
    D, H = cfg.hidden_size, cfg.num_attention_heads
    dh = D // H
    eps = getattr(cfg, "layer_norm_eps", 1e-5)
    act = {"quick_gelu": quick_gelu, "gelu": F.gelu}[cfg.hidden_act]
    x = sd[p + "embeddings.token_embedding.weight"][ids] + sd[p + "embeddings.position_embedding.weight"][:S]
    mask = torch.full((S, S), float("-inf")).triu(1)
    hs = [x]
    for i in range(cfg.num_hidden_layers):
        b = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), sd[b + "layer_norm1.weight"], sd[b + "layer_norm1.bias"], eps)
        q, k, v = (F.linear(h, sd[f"{b}self_attn.{n}_proj.weight"], sd[f"{b}self_attn.{n}_proj.bias"])
                   .view(B, S, H, dh).transpose(1, 2) for n in "qkv")
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + mask, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, S, D)
        x = x + F.linear(a, sd[b + "self_attn.out_proj.weight"], sd[b + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), sd[b + "layer_norm2.weight"], sd[b + "layer_norm2.bias"], eps)
        h = act(F.linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))
        x = x + F.linear(h, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
        hs.append(x)
    last = F.layer_norm(x, (D,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)
    # pooled: the EOS token (first occurrence of eos_token_id; legacy configs with eos id 2: argmax)
    if getattr(cfg, "eos_token_id", 2) == 2:
        pos = ids.argmax(dim=-1)
    else:
        pos = (ids == cfg.eos_token_id).int().argmax(dim=-1)
    pooled = last[torch.arange(B), pos]
    proj = sd.get("text_projection.weight")
    return hs, last, pooled, (F.linear(pooled, proj) if proj is not None else None)


# ------------------------------------------------------------------ T5 encoder
def t5_relative_buckets(S: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """T5Attention._relative_position_bucket, bidirectional: [S(query), S(key)] bucket ids."""
    ctx = torch.arange(S)[:, None]
    mem = torch.arange(S)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    out = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(is_small, rel, large)


def t5_encoder_forward(sd: Dict[str, torch.Tensor], cfg, ids: torch.Tensor) -> torch.Tensor:
    """T5 v1.1 encoder stack (T5EncoderModel(...)[0]): RMS layer norms without bias, attention
    without 1/sqrt(d) scaling plus the layer-0 relative position bias shared by all layers, gated
    tanh-GELU feed-forward, final layer norm. No attention mask (diffusers' SD3 _get_t5_prompt_embeds
    passes none: padding tokens are attended). cfg: d_model, d_kv, d_ff, num_layers, num_heads,
    relative_attention_num_buckets, relative_attention_max_distance, layer_norm_epsilon."""
    B, S = ids.shape
    H, dk = cfg.num_heads, cfg.d_kv
    eps = cfg.layer_norm_epsilon

    def rms(x, w):
        var = x.float().pow(2).mean(-1, keepdim=True)
        return w * (x * torch.rsqrt(var + eps))

    x = sd["encoder.embed_tokens.weight" if "encoder.embed_tokens.weight" in sd else "shared.weight"][ids]
    buckets = t5_relative_buckets(S, cfg.relative_attention_num_buckets, cfg.relative_attention_max_distance)
    bias = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"][buckets]  # [S, S, H]
    bias = bias.permute(2, 0, 1)[None]
    for i in range(cfg.num_layers):
        b = f"encoder.block.{i}.layer."
        h = rms(x, sd[b + "0.layer_norm.weight"])
        q, k, v = (F.linear(h, sd[f"{b}0.SelfAttention.{n}.weight"]).view(B, S, H, dk).transpose(1, 2)
                   for n in "qkv")
        a = torch.softmax((q @ k.transpose(-1, -2) + bias).float(), dim=-1).to(q.dtype) @ v
        x = x + F.linear(a.transpose(1, 2).reshape(B, S, H * dk), sd[b + "0.SelfAttention.o.weight"])
        h = rms(x, sd[b + "1.layer_norm.weight"])
        g = F.gelu(F.linear(h, sd[b + "1.DenseReluDense.wi_0.weight"]), approximate="tanh")
        h = g * F.linear(h, sd[b + "1.DenseReluDense.wi_1.weight"])
        x = x + F.linear(h, sd[b + "1.DenseReluDense.wo.weight"])
    return rms(x, sd["encoder.final_layer_norm.weight"])


# ------------------------------------------------------------------ encode_prompt assembly
def sd3_prompt_embeds(clip_l, clip_g, t5, ids_l, ids_g, ids_t5, joint_dim=4096):
    """diffusers StableDiffusion3Pipeline.encode_prompt for one branch (positive or negative),
    clip_skip=None: (prompt_embeds [B, 77 + S_t5, joint_dim], pooled [B, D_l + D_g]).
    clip_l / clip_g / t5 = (state_dict, config)."""
    hs_l, _, _, emb_l = clip_text_forward(*clip_l, ids_l)
    hs_g, _, _, emb_g = clip_text_forward(*clip_g, ids_g)
    clip = torch.cat([hs_l[-2], hs_g[-2]], dim=-1)            # penultimate hidden states
    pooled = torch.cat([emb_l, emb_g], dim=-1)                # projected EOS embeddings
    t5e = t5_encoder_forward(*t5, ids_t5)
    clip = F.pad(clip, (0, joint_dim - clip.shape[-1]))
    return torch.cat([clip, t5e], dim=-2), pooled


def sdxl_prompt_embeds(clip_l, clip_g, ids_l, ids_g):
    """diffusers StableDiffusionXLPipeline.encode_prompt for one branch:
    (prompt_embeds [B, 77, D_l + D_g], pooled [B, proj_g])."""
    hs_l, _, _, _ = clip_text_forward(*clip_l, ids_l)
    hs_g, _, _, emb_g = clip_text_forward(*clip_g, ids_g)
    return torch.cat([hs_l[-2], hs_g[-2]], dim=-1), emb_g
