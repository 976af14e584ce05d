"""ORACLE (test infrastructure, not product code): CPU/PyTorch fp32 restatement of the VAE decode
that sduss' post_inference stage runs on the final latents of a batch of requests (SURVEY.md §8
row f-4; reference call sites:
sduss/model_executor/diffusers/pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl_esymred.py:406-462
-- latents / scaling_factor, fp32-upcast `vae.decode`, one resolution at a time -- and
pipelines/stable_diffusion_3/pipeline_stable_diffusion_3_esymred.py:391-415
-- latents / scaling_factor + shift_factor, `vae.decode`).

Layer arithmetic = diffusers==0.32.1 `AutoencoderKL.decode` (`post_quant_conv` -> `Decoder`:
conv_in, UNetMidBlock2D [resnet, single-head attention with head_dim = channels, resnet],
UpDecoderBlock2D x4 [3 resnets without time embedding, nearest 2x upsample + conv], GroupNorm(32,
eps 1e-6) + SiLU, conv_out). Third party (conda.yml:50), not vendored, not installable here:
restated from that release's published semantics, state-dict names are diffusers'.
PARITY UNPINNED for the layer arithmetic (the reference has no golden vectors for this stage).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.13025
    shift_factor: Optional[float] = None
    use_post_quant_conv: bool = True


def sdxl_vae_config() -> VAEConfig:
    return VAEConfig()


def sd3_vae_config() -> VAEConfig:
    return VAEConfig(latent_channels=16, scaling_factor=1.5305, shift_factor=0.0609,
                     use_post_quant_conv=False)


def vae_tiny_config(latent_channels=4, shift=None, pq=True) -> VAEConfig:
    return VAEConfig(latent_channels=latent_channels, block_out_channels=(64, 64, 128, 128),
                     layers_per_block=1, scaling_factor=0.5, shift_factor=shift,
                     use_post_quant_conv=pq)


# ------------------------------------------------------------------ weights
def init_vae_decoder_weights(cfg: VAEConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, fin, fout, gain=1.0):
        sd[name + ".weight"] = torch.randn(fout, fin, generator=g) * (gain / math.sqrt(fin))
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
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3, gain=0.5)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    ch = cfg.block_out_channels
    top = ch[-1]
    if cfg.use_post_quant_conv:
        conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    conv("decoder.conv_in", cfg.latent_channels, top, 3)
    resnet("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for n in ("to_q", "to_k", "to_v"):
        lin(f"{a}.{n}", top, top)
    lin(a + ".to_out.0", top, top, gain=0.5)
    resnet("decoder.mid_block.resnets.1", top, top)
    rev = list(reversed(ch))
    prev = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        prev = c
        if i != len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("decoder.conv_norm_out", rev[-1])
    conv("decoder.conv_out", rev[-1], cfg.out_channels, 3)
    return sd


# ------------------------------------------------------------------ layers
def _conv(sd, name, x, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=padding)


def _gn(sd, cfg, name, x):
    return F.group_norm(x, cfg.norm_num_groups, sd[name + ".weight"], sd[name + ".bias"], cfg.norm_eps)


def resnet_block(sd, cfg, name, x):
    """diffusers ResnetBlock2D with temb_channels=None, output_scale_factor=1."""
    h = _conv(sd, name + ".conv1", F.silu(_gn(sd, cfg, name + ".norm1", x)))
    h = _conv(sd, name + ".conv2", F.silu(_gn(sd, cfg, name + ".norm2", h)))
    if name + ".conv_shortcut.weight" in sd:
        x = _conv(sd, name + ".conv_shortcut", x, padding=0)
    return x + h


def attention_block(sd, cfg, name, x):
    """diffusers Attention(heads=1, dim_head=C, group_norm, residual_connection=True, bias=True)."""
    B, C, H, W = x.shape
    t = _gn(sd, cfg, name + ".group_norm", x).view(B, C, H * W).transpose(1, 2)
    q = F.linear(t, sd[name + ".to_q.weight"], sd[name + ".to_q.bias"])
    k = F.linear(t, sd[name + ".to_k.weight"], sd[name + ".to_k.bias"])
    v = F.linear(t, sd[name + ".to_v.weight"], sd[name + ".to_v.bias"])
    p = torch.softmax(q @ k.transpose(1, 2) * (1.0 / math.sqrt(C)), dim=-1)
    o = F.linear(p @ v, sd[name + ".to_out.0.weight"], sd[name + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def decode_single(sd, cfg: VAEConfig, z: torch.Tensor) -> torch.Tensor:
    """z: [B, latent_channels, h, w], already unscaled -> image [B, 3, 8h, 8w] in [-1, 1]."""
    if cfg.use_post_quant_conv:
        z = _conv(sd, "post_quant_conv", z, padding=0)
    x = _conv(sd, "decoder.conv_in", z)
    x = resnet_block(sd, cfg, "decoder.mid_block.resnets.0", x)
    x = attention_block(sd, cfg, "decoder.mid_block.attentions.0", x)
    x = resnet_block(sd, cfg, "decoder.mid_block.resnets.1", x)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = resnet_block(sd, cfg, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i != n - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(_gn(sd, cfg, "decoder.conv_norm_out", x))
    return _conv(sd, "decoder.conv_out", x)


def unscale_latents(cfg: VAEConfig, latents: torch.Tensor) -> torch.Tensor:
    """The two reference pipelines: SDXL `latents / scaling_factor` (xl_esymred.py:441),
    SD3 `latents / scaling_factor + shift_factor` (3_esymred.py:408)."""
    z = latents / cfg.scaling_factor
    if cfg.shift_factor is not None:
        z = z + cfg.shift_factor
    return z


def vae_decode(sd, cfg: VAEConfig, latents: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """latents: {resolution: [n, C, h, w]} scheduler-space latents -> {resolution: [n, 3, 8h, 8w]}.
    Every image is decoded on its own (batching never changes a result)."""
    return {res: torch.cat([decode_single(sd, cfg, unscale_latents(cfg, z[i:i + 1].float()))
                            for i in range(z.shape[0])]) for res, z in latents.items()}


def postprocess(image: torch.Tensor) -> torch.Tensor:
    """VaeImageProcessor.postprocess up to the float image: denormalize to [0, 1]."""
    return (image / 2 + 0.5).clamp(0, 1)


def vae_decode_flops(cfg: VAEConfig, h: int, w: int) -> float:
    """Multiply-add FLOPs (x2) of one decode of an h x w latent."""
    ch = list(reversed(cfg.block_out_channels))
    top = ch[0]
    px = h * w
    fl = 2.0 * 9 * cfg.latent_channels * top * px
    res = lambda cin, cout, p: 2.0 * p * (9 * cin * cout + 9 * cout * cout + (cin * cout if cin != cout else 0))
    fl += 2 * res(top, top, px) + 2.0 * px * 4 * top * top + 4.0 * px * px * top
    prev = top
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block + 1):
            fl += res(prev if j == 0 else c, c, px)
        prev = c
        if i != len(ch) - 1:
            px *= 4
            fl += 2.0 * 9 * c * c * px
    fl += 2.0 * 9 * ch[-1] * cfg.out_channels * px
    return fl
