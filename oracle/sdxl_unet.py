"""ORACLE (test infrastructure, not product code): CPU/PyTorch restatement of the SDXL-base
UNet forward that sduss' PatchUNet.forward computes for a dict of mixed-resolution latents
(reference: sduss/model_executor/modules/unet.py:205-530, modules/resnet.py:380-460,
modules/transformer.py:32-290, modules/attention.py:52-232, modules/unet_2d_blocks.py).

Layer arithmetic = diffusers==0.32.1 UNet2DConditionModel (third party, conda.yml:50, not
vendored, not installable here), restated from that release's published semantics
(SURVEY.md Appendix A1-A5); state-dict names are diffusers'. PARITY UNPINNED for the layer
arithmetic (no golden vectors in the reference).

This is the EXACT per-latent math (what vanilla diffusers computes on each image alone). The
reference's patched path deviates from it in two documented ways that this oracle does not
reproduce: D1 (GroupNorm variance = mean of per-patch variances,
kernels/norm_silu_concat.cu:361-386) and D2 (wrong corner halos, :210-239). See DESIGN.md.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from .sd3_mmdit import timestep_embedding


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    down_has_attn: Tuple[bool, ...] = (False, True, True)
    num_heads: Tuple[int, ...] = (5, 10, 20)      # attention_head_dim in diffusers' config
    cross_attention_dim: int = 2048
    context_len: int = 77
    addition_time_embed_dim: int = 256
    pooled_dim: int = 1280
    norm_num_groups: int = 32
    norm_eps: float = 1e-5

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    @property
    def add_in_dim(self):
        return self.pooled_dim + 6 * self.addition_time_embed_dim  # 2816 for SDXL


def sdxl_base_config() -> UNetConfig:
    return UNetConfig()


def sdxl_tiny_config() -> UNetConfig:
    return UNetConfig(block_out_channels=(64, 128, 256), transformer_layers_per_block=(1, 1, 2),
                      num_heads=(1, 2, 4), cross_attention_dim=128, context_len=13,
                      addition_time_embed_dim=32, pooled_dim=64)


# ------------------------------------------------------------------ weights
def init_unet_weights(cfg: UNetConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    T = cfg.time_embed_dim

    def lin(name, fin, fout, bias=True, gain=1.0):
        sd[name + ".weight"] = torch.randn(fout, fin, generator=g) * (gain / math.sqrt(fin))
        if bias:
            sd[name + ".bias"] = torch.randn(fout, generator=g) * 0.02

    def conv(name, cin, cout, k, gain=1.0):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * (gain / math.sqrt(cin * k * k))
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def norm(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        lin(name + ".time_emb_proj", T, cout)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3, gain=0.5)
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
    lin("add_embedding.linear_1", cfg.add_in_dim, T)
    lin("add_embedding.linear_2", T, T)
    # down
    out_c = ch[0]
    for i, c in enumerate(ch):
        in_c, out_c = out_c, c
        for j in range(cfg.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
            if cfg.down_has_attn[i]:
                transformer(f"down_blocks.{i}.attentions.{j}", out_c, cfg.transformer_layers_per_block[i])
        if i != len(ch) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    # mid
    resnet("mid_block.resnets.0", ch[-1], ch[-1])
    transformer("mid_block.attentions.0", ch[-1], cfg.transformer_layers_per_block[-1])
    resnet("mid_block.resnets.1", ch[-1], ch[-1])
    # up
    rev = list(reversed(ch))
    rev_layers = list(reversed(cfg.transformer_layers_per_block))
    rev_attn = list(reversed(cfg.down_has_attn))
    out_c = rev[0]
    for i, c in enumerate(rev):
        prev_out, out_c = out_c, c
        in_c = rev[min(i + 1, len(ch) - 1)]
        for j in range(cfg.layers_per_block + 1):
            skip_c = in_c if j == cfg.layers_per_block else out_c
            res_in = prev_out if j == 0 else out_c
            resnet(f"up_blocks.{i}.resnets.{j}", res_in + skip_c, out_c)
            if rev_attn[i]:
                transformer(f"up_blocks.{i}.attentions.{j}", out_c, rev_layers[i])
        if i != len(ch) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    norm("conv_norm_out", ch[0])
    conv("conv_out", ch[0], cfg.out_channels, 3)
    return {k: v.to(dtype) for k, v in sd.items()}


# ------------------------------------------------------------------ layers (A2-A5)
def _linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _gn(sd, name, x, groups, eps):
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def resnet_block(sd, cfg, name, x, emb):
    h = F.silu(_gn(sd, name + ".norm1", x, cfg.norm_num_groups, cfg.norm_eps))
    h = _conv(sd, name + ".conv1", h)
    h = h + _linear(sd, name + ".time_emb_proj", F.silu(emb))[:, :, None, None]
    h = F.silu(_gn(sd, name + ".norm2", h, cfg.norm_num_groups, cfg.norm_eps))
    h = _conv(sd, name + ".conv2", h)
    if name + ".conv_shortcut.weight" in sd:
        x = _conv(sd, name + ".conv_shortcut", x, padding=0)
    return x + h


def _mha(q, k, v, heads):
    B, Sq, C = q.shape
    d = C // heads
    q = q.view(B, Sq, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)
    return o.transpose(1, 2).reshape(B, Sq, C)


def basic_block(sd, name, x, ctx, heads):
    h = _ln(sd, name + ".norm1", x)
    a = _mha(_linear(sd, name + ".attn1.to_q", h), _linear(sd, name + ".attn1.to_k", h),
             _linear(sd, name + ".attn1.to_v", h), heads)
    x = x + _linear(sd, name + ".attn1.to_out.0", a)
    h = _ln(sd, name + ".norm2", x)
    a = _mha(_linear(sd, name + ".attn2.to_q", h), _linear(sd, name + ".attn2.to_k", ctx),
             _linear(sd, name + ".attn2.to_v", ctx), heads)
    x = x + _linear(sd, name + ".attn2.to_out.0", a)
    h = _ln(sd, name + ".norm3", x)
    hid, gate = _linear(sd, name + ".ff.net.0.proj", h).chunk(2, dim=-1)
    return x + _linear(sd, name + ".ff.net.2", hid * F.gelu(gate))


def transformer_2d(sd, cfg, name, x, ctx, heads, layers):
    B, C, Hh, Ww = x.shape
    res = x
    h = _gn(sd, name + ".norm", x, cfg.norm_num_groups, 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(B, Hh * Ww, C)
    h = _linear(sd, name + ".proj_in", h)
    for j in range(layers):
        h = basic_block(sd, f"{name}.transformer_blocks.{j}", h, ctx, heads)
    h = _linear(sd, name + ".proj_out", h)
    return h.reshape(B, Hh, Ww, C).permute(0, 3, 1, 2) + res


def conditioning(sd, cfg, timestep, text_embeds, time_ids):
    """get_time_embed + time_embedding + get_aug_embed('text_time') (unet.py:314-334)."""
    t = timestep_embedding(timestep, cfg.block_out_channels[0]).to(text_embeds.dtype)
    emb = _linear(sd, "time_embedding.linear_2", F.silu(_linear(sd, "time_embedding.linear_1", t)))
    ids = timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim)
    ids = ids.reshape(text_embeds.shape[0], -1).to(text_embeds.dtype)
    add = torch.cat([text_embeds, ids], dim=-1)
    aug = _linear(sd, "add_embedding.linear_2", F.silu(_linear(sd, "add_embedding.linear_1", add)))
    return emb + aug


def unet_single(sd, cfg: UNetConfig, x, emb, ctx):
    """x: [B,4,h,w] same-resolution latents; emb: [B,T]; ctx: [B,77,2048]."""
    ch = cfg.block_out_channels
    x = _conv(sd, "conv_in", x)
    skips = [x]
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            x = resnet_block(sd, cfg, f"down_blocks.{i}.resnets.{j}", x, emb)
            if cfg.down_has_attn[i]:
                x = transformer_2d(sd, cfg, f"down_blocks.{i}.attentions.{j}", x, ctx,
                                   cfg.num_heads[i], cfg.transformer_layers_per_block[i])
            skips.append(x)
        if i != len(ch) - 1:
            x = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", x, stride=2)
            skips.append(x)
    x = resnet_block(sd, cfg, "mid_block.resnets.0", x, emb)
    x = transformer_2d(sd, cfg, "mid_block.attentions.0", x, ctx, cfg.num_heads[-1],
                       cfg.transformer_layers_per_block[-1])
    x = resnet_block(sd, cfg, "mid_block.resnets.1", x, emb)
    rev_layers = list(reversed(cfg.transformer_layers_per_block))
    rev_attn = list(reversed(cfg.down_has_attn))
    rev_heads = list(reversed(cfg.num_heads))
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, cfg, f"up_blocks.{i}.resnets.{j}", x, emb)
            if rev_attn[i]:
                x = transformer_2d(sd, cfg, f"up_blocks.{i}.attentions.{j}", x, ctx, rev_heads[i],
                                   rev_layers[i])
        if i != len(ch) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(_gn(sd, "conv_norm_out", x, cfg.norm_num_groups, cfg.norm_eps))
    return _conv(sd, "conv_out", x)


@torch.no_grad()
def unet_forward(sd, cfg: UNetConfig, sample: Dict[str, torch.Tensor], timestep: torch.Tensor,
                 encoder_hidden_states: torch.Tensor, text_embeds: torch.Tensor,
                 time_ids: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Same contract as PatchUNet.forward (unet.py:205-225,521-530): dict resolution ->
    [n_r,4,h,w]; conditioning rows ordered resolution by resolution in dict order."""
    emb_all = conditioning(sd, cfg, timestep, text_embeds, time_ids)
    out, base = {}, 0
    for res, lat in sample.items():
        n = lat.shape[0]
        if n == 0:
            continue
        out[res] = unet_single(sd, cfg, lat, emb_all[base:base + n], encoder_hidden_states[base:base + n])
        base += n
    return out


def unet_flops_per_latent(cfg: UNetConfig, res: int) -> float:
    """Algorithmic FLOPs (2*MAC) of one forward of one latent: conv 2*Cin*Cout*k^2*Ho*Wo,
    linear 2*tokens*Cin*Cout, attention 4*Sq*Skv*d*heads; text K/V and embedding linears once
    per latent."""
    ch = cfg.block_out_channels
    T = cfg.time_embed_dim
    fl = 0.0
    side = res // 8

    def conv(cin, cout, k, s):
        return 2.0 * cin * cout * k * k * s * s

    def resnet(cin, cout, s):
        f = conv(cin, cout, 3, s) + conv(cout, cout, 3, s) + 2.0 * T * cout
        if cin != cout:
            f += conv(cin, cout, 1, s)
        return f

    def transformer(c, layers, s, heads):
        S = s * s
        f = 2 * 2.0 * S * c * c
        per = (4 * 2.0 * S * c * c + 4.0 * S * S * 64 * heads            # self
               + 2 * 2.0 * S * c * c + 2 * 2.0 * cfg.context_len * cfg.cross_attention_dim * c
               + 4.0 * S * cfg.context_len * 64 * heads                   # cross
               + 2.0 * S * c * 8 * c + 2.0 * S * 4 * c * c)               # GEGLU ff
        return f + layers * per

    fl += conv(cfg.in_channels, ch[0], 3, side)
    fl += 2.0 * (ch[0] * T + T * T + cfg.add_in_dim * T + T * T)
    out_c, s = ch[0], side
    for i, c in enumerate(ch):
        in_c, out_c = out_c, c
        for j in range(cfg.layers_per_block):
            fl += resnet(in_c if j == 0 else out_c, out_c, s)
            if cfg.down_has_attn[i]:
                fl += transformer(out_c, cfg.transformer_layers_per_block[i], s, cfg.num_heads[i])
        if i != len(ch) - 1:
            s //= 2
            fl += conv(out_c, out_c, 3, s)
    fl += 2 * resnet(ch[-1], ch[-1], s)
    fl += transformer(ch[-1], cfg.transformer_layers_per_block[-1], s, cfg.num_heads[-1])
    rev = list(reversed(ch))
    rev_layers = list(reversed(cfg.transformer_layers_per_block))
    rev_attn = list(reversed(cfg.down_has_attn))
    rev_heads = list(reversed(cfg.num_heads))
    out_c = rev[0]
    for i, c in enumerate(rev):
        prev_out, out_c = out_c, c
        in_c = rev[min(i + 1, len(ch) - 1)]
        for j in range(cfg.layers_per_block + 1):
            skip_c = in_c if j == cfg.layers_per_block else out_c
            fl += resnet((prev_out if j == 0 else out_c) + skip_c, out_c, s)
            if rev_attn[i]:
                fl += transformer(out_c, rev_layers[i], s, rev_heads[i])
        if i != len(ch) - 1:
            s *= 2
            fl += conv(out_c, out_c, 3, s)
    fl += conv(ch[0], cfg.out_channels, 3, side)
    return fl
