"""Pins the MODEL ARITHMETIC of the oracle (oracle/sd3_mmdit.py, oracle/sdxl_unet.py,
oracle/vae_decoder.py) against the third-party code the reference actually runs -- diffusers 0.32.1
(`/root/reference/conda.yml:50`) -- the day a machine has it. The build container and the GPU boxes
do not (no network), which is why this is the ONE generator whose fixtures are not committed yet and
why DESIGN.md section 5 says "parity unpinned" for the layer arithmetic.

    pip install diffusers==0.32.1 && python tools/make_golden_diffusers.py

builds diffusers' SD3Transformer2DModel / UNet2DConditionModel at the oracle's TINY configs, loads the
oracle's seeded weights into them (the oracle keeps diffusers' state-dict names for exactly this),
runs seeded inputs through the stock diffusers forward and writes tests/golden/diffusers_{sd3,sdxl}.npz
(inputs + outputs). tests/test_diffusers_golden.py compares the oracle with those outputs and is
skipped while the files are absent."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def gen_sd3():
    from diffusers import SD3Transformer2DModel
    from oracle import sd3_mmdit as o3
    cfg = o3.sd3_tiny_config()
    model = SD3Transformer2DModel(
        sample_size=cfg.sample_size, patch_size=cfg.patch_size, in_channels=cfg.in_channels,
        num_layers=cfg.num_layers, attention_head_dim=cfg.attention_head_dim,
        num_attention_heads=cfg.num_attention_heads, joint_attention_dim=cfg.joint_attention_dim,
        caption_projection_dim=cfg.caption_projection_dim, pooled_projection_dim=cfg.pooled_projection_dim,
        out_channels=cfg.out_channels, pos_embed_max_size=cfg.pos_embed_max_size,
        dual_attention_layers=tuple(cfg.dual_attention_layers), qk_norm="rms_norm").eval()
    sd = o3.init_sd3_weights(cfg, 0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("pos_embed.pos_embed" in k for k in missing), (missing, unexpected)
    g = torch.Generator().manual_seed(1)
    cases = {}
    for res in (256, 512):
        x = torch.randn(2, cfg.in_channels, res // 8, res // 8, generator=g)
        ehs = torch.randn(2, cfg.context_len, cfg.joint_attention_dim, generator=g)
        pooled = torch.randn(2, cfg.pooled_projection_dim, generator=g)
        t = torch.tensor([981.0, 333.0])
        with torch.no_grad():
            y = model(hidden_states=x, encoder_hidden_states=ehs, pooled_projections=pooled, timestep=t,
                      return_dict=False)[0]
        for k, v in (("x", x), ("ehs", ehs), ("pooled", pooled), ("t", t), ("y", y)):
            cases[f"{res}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "diffusers_sd3.npz"), **cases)


def gen_sdxl():
    from diffusers import UNet2DConditionModel
    from oracle import sdxl_unet as ox
    cfg = ox.sdxl_tiny_config()
    n = len(cfg.block_out_channels)
    down = tuple("CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg.down_has_attn)
    up = tuple("CrossAttnUpBlock2D" if a else "UpBlock2D" for a in reversed(cfg.down_has_attn))
    model = UNet2DConditionModel(
        sample_size=32, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
        down_block_types=down, up_block_types=up, mid_block_type="UNetMidBlock2DCrossAttn",
        block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
        transformer_layers_per_block=cfg.transformer_layers_per_block, attention_head_dim=cfg.num_heads,
        cross_attention_dim=cfg.cross_attention_dim, norm_num_groups=cfg.norm_num_groups, norm_eps=cfg.norm_eps,
        use_linear_projection=True, addition_embed_type="text_time",
        addition_time_embed_dim=cfg.addition_time_embed_dim,
        projection_class_embeddings_input_dim=cfg.add_in_dim, flip_sin_to_cos=True, freq_shift=0).eval()
    sd = ox.init_unet_weights(cfg, 0)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1)
    cases = {}
    for res in (256, 512):
        x = torch.randn(2, cfg.in_channels, res // 8, res // 8, generator=g)
        ehs = torch.randn(2, cfg.context_len, cfg.cross_attention_dim, generator=g)
        te = torch.randn(2, cfg.pooled_dim, generator=g)
        ids = torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]] * 2)
        t = torch.tensor([981.0, 333.0])
        with torch.no_grad():
            y = model(x, t, encoder_hidden_states=ehs, added_cond_kwargs={"text_embeds": te, "time_ids": ids},
                      return_dict=False)[0]
        for k, v in (("x", x), ("ehs", ehs), ("te", te), ("ids", ids), ("t", t), ("y", y)):
            cases[f"{res}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "diffusers_sdxl.npz"), **cases)


if __name__ == "__main__":
    import diffusers
    print("diffusers", diffusers.__version__, "(the reference pins 0.32.1)")
    gen_sd3()
    gen_sdxl()
    print("wrote tests/golden/diffusers_{sd3,sdxl}.npz -- commit them and tests/test_diffusers_golden.py starts running")
