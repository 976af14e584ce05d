"""Patch cache, SDXL variant, on the GPU: the up-block decision (features continue
with the MSE of the block's skip tensors, CacheManager.get_mask(is_upsample=True),
cache_manager.py:106-135) against sklearn and the reference-pinned bookkeeping, and the cached UNet
step against oracle/patch_cache.py::CachedSDXLOracle (same masks) and the exact step."""
from dataclasses import asdict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BLOCKS = ("down_blocks.0", "down_blocks.1", "down_blocks.2", "mid_block", "up_blocks.0", "up_blocks.1", "up_blocks.2")


def test_up_block_decision_uses_the_skip_tensors(cuda):
    from sklearn.ensemble import RandomForestClassifier
    from oracle import patch_cache as pc
    from sduss_b200 import ops
    rng = np.random.RandomState(0)
    # forest on [block, timestep, mse(input), mse(skip 0), mse(skip 1)]: recompute when ANY tensor moved
    X = np.stack([rng.randint(0, 7, 6000), rng.uniform(0, 1000, 6000)] +
                 [10 ** rng.uniform(-6, 0, 6000) for _ in range(3)], 1)
    y = (X[:, 2:].max(1) > 2e-3).astype(int)
    rf = RandomForestClassifier(n_estimators=10, max_depth=8, random_state=0).fit(X.astype(np.float32), y)
    forest = ops.DeviceForest.from_sklearn(rf, cuda)
    g = torch.Generator().manual_seed(3)
    S = [256, 1024, 512]
    L, T, n = len(S), sum(S), sum(S) // 256
    patch_latent = torch.tensor(np.repeat(np.arange(L), [s // 256 for s in S]), dtype=torch.int32).cuda()
    t32 = torch.tensor([900.0, 500.0, 120.0]).cuda()
    dims = (128, 64, 192)                         # block input, two skip tensors (other channel counts)
    prev0 = [torch.randn(T, d, generator=g).cuda().bfloat16() for d in dims]
    drift = [(10 ** (torch.rand(n, generator=g) * 4.5 - 4.5))[:, None].repeat_interleave(256, 0) for _ in dims]
    xs = [(p.float().cpu() + s * torch.randn(T, d, generator=g)).cuda().bfloat16()
          for p, s, d in zip(prev0, drift, dims)]
    ws = ops.patch_mask_workspace(n, cuda)
    valid_host, skipped_host = [1, 1, 0], [0, 4, 2, 1, 3, 4, 0]
    valid = torch.tensor(valid_host, dtype=torch.float32).cuda()
    prev = [p.clone() for p in prev0]
    extra = torch.zeros(2, n).cuda()
    mask = torch.full((n,), -1, dtype=torch.int32).cuda()
    skipped = torch.tensor(skipped_host, dtype=torch.int32).cuda()
    for k in (1, 2):                              # MSE-only launches: no decision, the kept copy is refreshed
        ops.patch_mse(xs[k], prev[k], patch_latent, t32, valid, extra[k - 1], ws)
    torch.cuda.synchronize()
    assert (mask == -1).all() and skipped.cpu().tolist() == skipped_host
    assert torch.equal(prev[1], xs[1]) and torch.equal(prev[2], xs[2])
    mse = torch.zeros(n).cuda()
    ops.patch_mask(xs[0], prev[0], patch_latent, t32, valid, skipped, mask, forest, 5, 4, ws, mse=mse, extra_mse=extra)
    torch.cuda.synchronize()
    lat = patch_latent.cpu().numpy()
    ok = np.asarray(valid_host)[lat] > 0
    want = [((x.float() - p.float()) ** 2).view(n, -1).mean(1).cpu().numpy() for x, p in zip(xs, prev0)]
    got = [mse.cpu().numpy(), extra[0].cpu().numpy(), extra[1].cpu().numpy()]
    for w_, g_ in zip(want, got):
        assert np.allclose(g_[ok], w_[ok], rtol=2e-4) and (g_[~ok] == np.float32(pc.MSE_MISSING)).all()
    feats = np.stack([np.full(n, 5), t32.cpu().numpy()[lat]] + got, 1).astype(np.float32)
    pred = rf.predict(feats)
    pred[~ok] = 1
    want_mask, want_cnt = pc.mask_bookkeeping(list(ok), skipped_host, list(pred), 4)
    assert mask.cpu().tolist() == [int(m) for m in want_mask]
    assert skipped.cpu().tolist() == want_cnt
    assert 0 < sum(want_mask) < n
    # the skip tensors matter: with their MSE zeroed the forest sees only the input's
    pred0 = rf.predict(np.concatenate([feats[:, :3], np.zeros((n, 2), np.float32)], 1))
    assert (pred0 != pred)[ok].any()


def _tiny_pipe(cuda):
    from oracle import sdxl_unet as ox
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline
    from sduss_b200.schedulers import B200EulerDiscreteScheduler
    from sduss_b200.unet import B200UNet, UNetConfig
    oc = ox.sdxl_tiny_config()
    d = asdict(oc)
    d.pop("context_len")
    sd = {k: v.to(torch.bfloat16).float() for k, v in ox.init_unet_weights(oc, 0).items()}
    model = B200UNet(sd, UNetConfig(**d), device=cuda)
    sched = B200EulerDiscreteScheduler()
    return oc, sd, model, sched, B200StableDiffusionXLPipeline(model, sched)


def _step(pipe, reqs):
    pipe.denoising_step(reqs, True, 0.0, 5.0, None, {}, None, None, None, True, 256)
    torch.cuda.synchronize()


def test_cached_sdxl_step_matches_oracle_policy_and_exact_forward(cuda):
    """Four steps of a two-request batch (512^2 and 1024^2, CFG on: 4 latents; 16 / 64 patches per latent
    at level 0, 4 / 16 at level 1, 1 / 4 at level 2) with the cache on. Step 0: nothing kept, everything computed -- equal to
    the uncached step bit for bit. Step 1: rule 'never' -> every patch reused. Steps 2, 3: rule 'MSE > the
    median of what down block 0 saw', part of one image disturbed in between -> a mixture. Per step and
    request the applied prediction is compared with CachedSDXLOracle run with the SAME masks."""
    import _parity as P
    from oracle import patch_cache as pc
    from oracle import schedulers as osch
    from oracle import sdxl_unet as ox
    from sduss_b200 import ops
    from sduss_b200.synthetic import make_sdxl_requests
    oc, sd, model, sched, pipe = _tiny_pipe(cuda)
    _, _, model0, _, pipe0 = _tiny_pipe(cuda)
    model.use_graphs = False
    mk = lambda: make_sdxl_requests(oc, {"512": 1, "1024": 1}, 50, sched, cuda, seed=5, latent_dtype=torch.float32)
    reqs, reqs0 = mk(), mk()
    model.enable_patch_cache(ops.DeviceForest.threshold_rule(3e38, cuda), refresh=4)
    orc = pc.CachedSDXLOracle(sd, oc, refresh=4)
    sig, ts, _ = osch.euler_sigmas(50)
    order = [("512", 0), ("512", 1), ("1024", 0), ("1024", 1)]      # plan order: per resolution [uncond, cond]
    rows = {"512": {0: 4096, 1: 1024, 2: 256}, "1024": {0: 16384, 1: 4096, 2: 1024}}
    level = {"down_blocks.0": 0, "down_blocks.1": 1, "down_blocks.2": 2, "mid_block": 2, "up_blocks.0": 2,
             "up_blocks.1": 1, "up_blocks.2": 0}
    seen = []
    for k in range(4):
        before = P.snapshot(reqs)
        _step(pipe, reqs)
        plan = next(iter(model._plans.values()))
        dev_masks = {b: m.cpu().numpy().astype(bool) for b, m in plan.cache.masks().items()}
        assert set(dev_masks) == set(BLOCKS)
        seen.append(np.concatenate([dev_masks[b] for b in BLOCKS]))
        if k == 0:
            assert seen[0].all()
            _step(pipe0, reqs0)
            for res in reqs:
                assert torch.equal(reqs[res][0].sampling_params.latents, reqs0[res][0].sampling_params.latents)
        col = {b: 0 for b in BLOCKS}
        preds = {}
        for res, branch in order:
            r = reqs[res][0]
            x, kk = before[r.request_id]
            sp, po = r.sampling_params, r.prepare_output
            f = lambda t_: t_.float().cpu()
            xin = osch.batch_scale_model_input(x, [sig[kk]]).to(torch.bfloat16).float()
            ctx = f(sp.negative_prompt_embeds if branch == 0 else sp.prompt_embeds)
            te = f(po.negative_pooled_prompt_embeds if branch == 0 else po.pooled_prompt_embeds)
            ids = f(po.negative_add_time_ids if branch == 0 else po.add_time_ids)
            emb = ox.conditioning(sd, oc, torch.tensor([float(ts[kk])]), te, ids)
            forced = {}
            for b in BLOCKS:
                n = rows[res][level[b]] // 256
                forced[b] = list(dev_masks[b][col[b]:col[b] + n])
                col[b] += n
            preds[(res, branch)], _ = orc.forward_latent((r.request_id, branch), xin, emb, ctx, float(ts[kk]),
                                                         forced_masks=forced)
        for res in reqs:
            r = reqs[res][0]
            ref = preds[(res, 0)] + 5.0 * (preds[(res, 1)] - preds[(res, 0)])
            cos, err, scale = P.metrics(P.implied_prediction(before, r, sig), ref)
            assert P.ok(cos, err, scale), (k, res, cos, err / scale)
        if k == 1:
            assert not seen[1].any()                                 # rule 'never': everything reused
            tau = float(plan.cache.blocks["down_blocks.0"].mse.median())   # (the first block always sees the new latents)
            model.enable_patch_cache(ops.DeviceForest.threshold_rule(tau, cuda), refresh=4)
        # disturb the top quarter of the 1024^2 image before the next step
        lat = reqs["1024"][0].sampling_params.latents.clone()
        lat[:, :, :32] += 0.3 * torch.randn(lat[:, :, :32].shape, device=cuda,
                                            generator=torch.Generator(device="cuda").manual_seed(k))
        reqs["1024"][0].sampling_params.latents = lat
    assert 0.0 < seen[2].mean() < 1.0                                # a real mixture


def test_sdxl_cache_refresh_recomposition_and_unaligned_levels(cuda):
    from sduss_b200 import ops
    from sduss_b200.synthetic import make_sdxl_requests
    oc, sd, model, sched, pipe = _tiny_pipe(cuda)
    model.enable_patch_cache(ops.DeviceForest.threshold_rule(1e30, cuda), refresh=4)   # never flags by itself
    a = make_sdxl_requests(oc, {"512": 1}, 50, sched, cuda, seed=1)
    b = make_sdxl_requests(oc, {"768": 1}, 50, sched, cuda, seed=2)
    b["768"][0].request_id = 99

    def masks(n_lat):
        torch.cuda.synchronize()
        pl = [p for p in model._plans.values() if p.L == n_lat and p.cache is not None][0]
        return {k: m.cpu().numpy().copy() for k, m in pl.cache.masks().items()}
    allm = lambda m: all(v.all() for v in m.values())
    nonem = lambda m: not any(v.any() for v in m.values())
    _step(pipe, a)
    assert allm(masks(2)) and set(masks(2)) == set(BLOCKS)   # first sight: everything computed
    for _ in range(4):                                       # graph capture happens on the way
        _step(pipe, a)
        assert nonem(masks(2))
    _step(pipe, a)
    assert allm(masks(2))                                    # four skips in a row -> forced refresh
    _step(pipe, {**a, **b})
    m4 = masks(4)
    # 768^2 at level 2 is 24 x 24 = 576 rows, not whole 256-row patches: those blocks run uncached
    assert set(m4) == {"down_blocks.0", "down_blocks.1", "up_blocks.1", "up_blocks.2"} and allm(m4)
    _step(pipe, a)
    assert allm(masks(2))                                    # back, but a step was missed: recompute
    _step(pipe, a)
    assert nonem(masks(2))
    assert torch.isfinite(a["512"][0].sampling_params.latents.float()).all()


@pytest.mark.parametrize("sizes,cin,cout,stride,scale", [
    ([(32, 32), (64, 64)], 64, 128, 1, 0),       # same level: one / four patches per 16-row strip
    ([(64, 64), (32, 32)], 128, 320, 1, 0),      # 256-wide tiles (CTA pairs above one round)
    ([(32, 32), (64, 64)], 64, 64, 2, 2),        # downsampler: the mask describes the INPUT level
    ([(32, 32), (64, 64)], 64, 128, 1, -2),      # convolution after the 2x upsample: mask one level down
])
def test_conv3x3_skips_strips_of_clean_patches(cuda, sizes, cin, cout, stride, scale):
    """Pixel blocks whose 16 output pixel rows touch only clean 256-row patches (of the level the mask
    describes) keep their previous output; every other block equals the unmasked convolution bit for bit."""
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    g = torch.Generator().manual_seed(0)
    lin = LevelLayout(sizes, cuda)
    osz = [(h // stride, w // stride) for h, w in sizes]
    lout = LevelLayout(osz, cuda)
    x = torch.randn(lin.T, cin, generator=g).cuda().bfloat16()
    wt = (torch.randn(cout, 9 * cin, generator=g) / (3 * cin ** 0.5)).cuda().bfloat16()
    bias = torch.randn(cout, generator=g).cuda().bfloat16()
    maps = ops.conv3x3_encode_maps(x, cin, lin.desc_host, stride)
    full = torch.zeros(lout.T, cout, device=cuda, dtype=torch.bfloat16)
    omaps = ops.conv3x3_encode_maps(full, cout, lout.desc_host, 1)
    ops.conv3x3(maps, lout.tiles, lout.n_tiles, lout.desc, cin, cout, stride, wt, full, out_maps=omaps,
                bias=bias)
    # mask level: rows of the output level x 4^(scale / 2)
    msz = [((h << 1, w << 1) if scale > 0 else (h >> 1, w >> 1) if scale < 0 else (h, w)) for h, w in osz]
    n_bands = sum(h * w for h, w in msz) // 256
    mask_host = (torch.rand(n_bands, generator=g) < 0.4).int()
    mask_host[0] = 0
    for mask_host in (mask_host, torch.zeros(n_bands, dtype=torch.int32)):
        out = torch.full((lout.T, cout), 7.0, device=cuda, dtype=torch.bfloat16)
        om = ops.conv3x3_encode_maps(out, cout, lout.desc_host, 1)
        ops.conv3x3(maps, lout.tiles, lout.n_tiles, lout.desc, cin, cout, stride, wt, out, out_maps=om,
                    bias=bias, row_mask=mask_host.cuda(), row_mask_scale=scale)
        torch.cuda.synchronize()
        want = torch.zeros(lout.T, dtype=torch.bool)
        off_o = off_m = 0
        for (ho, wo), (hm, wm) in zip(osz, msz):
            f = hm / ho
            for y0 in range(0, ho, 16):
                y1 = min(y0 + 16, ho)
                lo, hi = off_m + int(y0 * f) * wm, off_m + int(y1 * f) * wm
                if mask_host[lo // 256:(hi - 1) // 256 + 1].any():
                    want[off_o + y0 * wo:off_o + y1 * wo] = True
            off_o, off_m = off_o + ho * wo, off_m + hm * wm
        o, fl = out.cpu(), full.cpu()
        assert torch.equal(o[want], fl[want]) and (o[~want] == 7.0).all()
        assert mask_host.any() == want.any()
