"""CPU checks of the model oracles themselves (tiny configs): the properties the parity tests
lean on -- batch invariance, dict-order contract, FLOP / parameter accounting that reproduces
SURVEY.md §8d."""
import torch

from oracle import sd3_mmdit as o3
from oracle import sdxl_unet as ox


def test_sd3_oracle_batch_invariance_and_order():
    cfg = o3.sd3_tiny_config()
    sd = o3.init_sd3_weights(cfg, 0)
    g = torch.Generator().manual_seed(1)
    hs = {"256": torch.randn(2, 16, 32, 32, generator=g), "512": torch.randn(1, 16, 64, 64, generator=g)}
    ehs = torch.randn(3, cfg.context_len, cfg.joint_attention_dim, generator=g)
    pp = torch.randn(3, cfg.pooled_projection_dim, generator=g)
    t = torch.tensor([900.0, 500.0, 100.0])
    out = o3.sd3_forward(sd, cfg, hs, ehs, pp, t)
    assert list(out) == ["256", "512"] and out["256"].shape == hs["256"].shape
    solo = o3.sd3_forward(sd, cfg, {"512": hs["512"]}, ehs[2:], pp[2:], t[2:])
    assert torch.allclose(out["512"], solo["512"], atol=1e-5)
    one = o3.sd3_forward(sd, cfg, {"256": hs["256"][1:2]}, ehs[1:2], pp[1:2], t[1:2])
    assert torch.allclose(out["256"][1:2], one["256"], atol=1e-5)


def test_sdxl_oracle_batch_invariance():
    cfg = ox.sdxl_tiny_config()
    sd = ox.init_unet_weights(cfg, 0)
    g = torch.Generator().manual_seed(1)
    s = {"256": torch.randn(1, 4, 32, 32, generator=g), "512": torch.randn(2, 4, 64, 64, generator=g)}
    e = torch.randn(3, cfg.context_len, cfg.cross_attention_dim, generator=g)
    te = torch.randn(3, cfg.pooled_dim, generator=g)
    ids = torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]] * 3)
    t = torch.tensor([981.0, 961.0, 941.0])
    out = ox.unet_forward(sd, cfg, s, t, e, te, ids)
    solo = ox.unet_forward(sd, cfg, {"512": s["512"][1:2]}, t[2:3], e[2:3], te[2:3], ids[2:3])
    assert torch.allclose(out["512"][1:2], solo["512"], atol=1e-4)


def test_flop_and_parameter_accounting_matches_survey():
    c3, cx = o3.sd35_medium_config(), ox.sdxl_base_config()
    sd3 = [o3.sd3_flops_per_latent(c3, r) / 1e12 for r in (256, 512, 768, 1024)]
    sdxl = [ox.unet_flops_per_latent(cx, r) / 1e12 for r in (256, 512, 768, 1024)]
    for got, want in zip(sd3, (0.911, 2.443, 5.591, 11.250)):
        assert abs(got - want) < 2e-3
    for got, want in zip(sdxl, (0.428, 1.589, 3.641, 6.761)):
        assert abs(got - want) < 2e-3
    assert abs(2 * (sdxl[1] + sdxl[3]) - 16.70) < 0.01      # config 1 step
    assert abs(2 * (sd3[1] + sd3[2] + sd3[3]) - 38.57) < 0.01  # config 2 step


def test_vae_decoder_oracle_batch_independence_and_flops():
    """oracle/vae_decoder.py (row f-4): every image is decoded on its own, shapes follow the 8x
    upsampling, the two reference un-scalings are restated (SDXL: / scaling_factor; SD3: / scaling_factor
    + shift_factor), and the FLOP count agrees with the product-side bench accounting."""
    import torch
    from oracle import vae_decoder as ov
    from sduss_b200.synthetic import vae_decode_flops
    from sduss_b200.vae import VAEDecoderConfig
    cfg = ov.vae_tiny_config(latent_channels=16, shift=0.0609, pq=False)
    sd = ov.init_vae_decoder_weights(cfg, 0)
    g = torch.Generator().manual_seed(0)
    lat = {"64": torch.randn(2, 16, 8, 8, generator=g), "96": torch.randn(1, 16, 16, 8, generator=g)}
    out = ov.vae_decode(sd, cfg, lat)
    assert out["64"].shape == (2, 3, 64, 64) and out["96"].shape == (1, 3, 128, 64)
    solo = ov.vae_decode(sd, cfg, {"64": lat["64"][1:2]})
    assert torch.equal(solo["64"][0], out["64"][1])
    z = torch.ones(1, 16, 2, 2)
    assert torch.allclose(ov.unscale_latents(cfg, z), z / cfg.scaling_factor + 0.0609)
    assert torch.allclose(ov.unscale_latents(ov.sdxl_vae_config(), z[:, :4]), z[:, :4] / 0.13025)
    img = ov.postprocess(out["64"])
    assert img.min() >= 0 and img.max() <= 1
    for ocfg, pcfg in ((ov.sdxl_vae_config(), VAEDecoderConfig()),
                       (ov.sd3_vae_config(), VAEDecoderConfig(latent_channels=16, use_post_quant_conv=False))):
        assert ov.vae_decode_flops(ocfg, 128, 128) == vae_decode_flops(pcfg, 128, 128)
    assert 9e12 < ov.vae_decode_flops(ov.sdxl_vae_config(), 128, 128) < 12e12
