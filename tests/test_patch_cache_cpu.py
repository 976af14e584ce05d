"""Patch cache (row f-3), CPU side: the decision bookkeeping of oracle/patch_cache.py against a
fixture produced by the reference's own CacheManager.get_sd3_mask / get_mask
(tests/golden/cache_mask.npz, tools/make_golden.py), and the cached oracle's invariants."""
import json
import os

import numpy as np
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _replay(name, refresh):
    from oracle import patch_cache as pc
    z = np.load(os.path.join(G, "cache_mask.npz"))
    log = json.loads(bytes(z[name + "_log"]).decode())
    counters, cached = {}, {}
    for k, e in enumerate(log):
        keys = e["keys"]
        x = torch.from_numpy(z[f"{name}_x{k}"])
        feats = np.asarray(e["features"])
        # the feature rows: [block, timestep, mse]; a patch without cached input gets sys.maxsize
        assert (feats[:, 0] == 3).all() and np.allclose(feats[:, 1], 900.0 - 30 * k)
        mine = [float(pc.chunk_mse(x[i].reshape(1, -1), cached[key].reshape(1, -1) if key in cached else None, 1)[0])
                for i, key in enumerate(keys)]
        # D11: the reference pairs the MSE values of the patches it already knows with those patches
        # through set() iteration order (cache_manager.py:104,168), so WHICH known patch gets which
        # value depends on the string hash seed; the values themselves must agree
        known = [i for i, key in enumerate(keys) if key in cached]
        assert np.allclose(sorted(mine[i] for i in known), sorted(feats[i, 2] for i in known), rtol=1e-4, atol=1e-7)
        for i, key in enumerate(keys):
            if key not in cached:
                assert feats[i, 2] == pc.MSE_MISSING == mine[i]
        # bookkeeping given the predictor's answers on the rows the reference built
        predicted = [int(f[2] > 0.5) for f in feats]
        mask, new = pc.mask_bookkeeping([key in cached for key in keys], [counters.get(key, 0) for key in keys],
                                        predicted, refresh)
        assert mask == e["mask"], (name, k)
        counters = dict(zip(keys, new))          # patches that left the batch are forgotten (as in the reference)
        assert counters == e["counters"], (name, k)
        cached = {key: x[i] for i, key in enumerate(keys)}


def test_mask_bookkeeping_matches_reference_get_sd3_mask():
    _replay("sd3", refresh=2)


def test_mask_bookkeeping_matches_reference_get_mask_down_blocks():
    _replay("down", refresh=4)


def test_cached_oracle_is_exact_when_everything_is_flagged_and_skips_clean_patches():
    from oracle import patch_cache as pc
    from oracle import sd3_mmdit as o3
    cfg = o3.sd3_tiny_config()
    sd = o3.init_sd3_weights(cfg, 0)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 16, 64, 64, generator=g)
    ehs = torch.randn(1, cfg.context_len, cfg.joint_attention_dim, generator=g)
    pooled = torch.randn(1, cfg.pooled_projection_dim, generator=g)
    ref = o3.sd3_forward(sd, cfg, {"x": lat}, ehs, pooled, torch.tensor([900.0]))["x"]
    orc = pc.CachedSD3Oracle(sd, cfg, lambda i, t, mse: (mse > 1e-3).int())
    out, masks = orc.forward_latent("a", lat, ehs, pooled, 900.0)
    assert torch.equal(out, ref) and all(all(m) for m in masks)      # first sight: everything computed
    lat2 = lat.clone()
    lat2[:, :, :16] += 0.5 * torch.randn(1, 16, 16, 64, generator=g)    # only the first token chunk changes
    out2, masks2 = orc.forward_latent("a", lat2, ehs, pooled, 900.0)
    assert masks2 == [[True, False, False, False]] * cfg.num_layers
    ref2 = o3.sd3_forward(sd, cfg, {"x": lat2}, ehs, pooled, torch.tensor([900.0]))["x"]
    assert torch.nn.functional.cosine_similarity(out2.flatten(), ref2.flatten(), dim=0) > 0.9999
    _, masks3 = orc.forward_latent("a", lat2, ehs, pooled, 900.0)
    assert not any(any(m) for m in masks3)                           # nothing moved: every patch reused
    _, masks4 = orc.forward_latent("a", lat2, ehs, pooled, 900.0)
    assert masks4 == [[False, True, True, True]] * cfg.num_layers    # skipped twice -> forced refresh
    # forced all-true masks reproduce the exact forward whatever the cache holds
    out5, _ = orc.forward_latent("a", lat2, ehs, pooled, 900.0, forced_masks=[[True] * 4] * cfg.num_layers)
    assert torch.allclose(out5, ref2, atol=1e-5)


def test_cached_sdxl_oracle_is_exact_when_everything_is_flagged_and_follows_the_refresh_rule():
    """SDXL variant (oracle.patch_cache.CachedSDXLOracle): one decision per UNet block,
    refresh = 4 (cache_manager.py:147), an up block's features include the MSE of its skip tensors."""
    from oracle import patch_cache as pc
    from oracle import sdxl_unet as ox
    cfg = ox.sdxl_tiny_config()
    sd = ox.init_unet_weights(cfg, 0)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, 64, 64, generator=g)                    # level 1: 1024 rows = 4 patches, level 2: 1
    ctx = torch.randn(1, cfg.context_len, cfg.cross_attention_dim, generator=g)
    emb = ox.conditioning(sd, cfg, torch.tensor([801.0]), torch.randn(1, cfg.pooled_dim, generator=g),
                          torch.tensor([[512.0, 512, 0, 0, 512, 512]]))
    ref = ox.unet_single(sd, cfg, lat, emb, ctx)
    seen_feats = []

    def rule(index, t, feats):
        seen_feats.append((index, tuple(feats.shape)))
        return (feats > 1e-4).any(dim=1).int()
    orc = pc.CachedSDXLOracle(sd, cfg, rule, refresh=4)
    out, masks = orc.forward_latent("a", lat, emb, ctx, 801.0)
    assert torch.equal(out, ref)                                     # first sight: everything computed
    order = ("down_blocks.0", "down_blocks.1", "down_blocks.2", "mid_block", "up_blocks.0", "up_blocks.1", "up_blocks.2")
    assert tuple(masks) == order
    assert [len(masks[k]) for k in order] == [16, 4, 1, 1, 1, 4, 16] and all(all(m) for m in masks.values())
    # block numbering of modules/unet.py:369-503 and the feature count (up blocks: input + 3 skips)
    assert seen_feats == [(0, (16, 1)), (1, (4, 1)), (2, (1, 1)), (3, (1, 1)), (4, (1, 4)), (5, (4, 4)), (6, (16, 4))]
    pattern = []
    for _ in range(6):                                               # the same input again and again
        out2, masks2 = orc.forward_latent("a", lat, emb, ctx, 801.0)
        pattern.append([any(m) for m in masks2.values()])
        assert torch.allclose(out2, ref, atol=1e-5)                  # reuse of unchanged patches is exact
    assert pattern == [[False] * 7] * 4 + [[True] * 7] + [[False] * 7]   # four skips, then the forced refresh
    lat2 = lat.clone()
    lat2[:, :, :16] += 0.5 * torch.randn(1, 4, 16, 64, generator=g)  # the top quarter of the image changes
    out3, masks3 = orc.forward_latent("a", lat2, emb, ctx, 801.0)
    assert masks3["down_blocks.0"][0] and masks3["down_blocks.1"][0]   # (GroupNorm couples the whole image: the other bands move too)
    ref3 = ox.unet_single(sd, cfg, lat2, emb, ctx)
    assert torch.nn.functional.cosine_similarity(out3.flatten(), ref3.flatten(), dim=0) > 0.99
    full = {k: [True] * len(m) for k, m in masks3.items()}
    out4, _ = orc.forward_latent("a", lat2, emb, ctx, 801.0, forced_masks=full)
    assert torch.allclose(out4, ref3, atol=1e-5)
