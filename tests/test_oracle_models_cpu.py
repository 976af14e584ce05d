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
