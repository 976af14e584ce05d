"""Patch cache on the GPU: the on-device decision kernel (MSE + flattened
RandomForest + refresh rule) against sklearn and the reference-pinned bookkeeping, the tile-skip
arguments of the GEMM / LayerNorm / attention kernels, and the cached SD3 forward against
oracle/patch_cache.py (same masks) and against the exact forward (all patches flagged)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_patch_mask_kernel_matches_sklearn_and_reference_bookkeeping(cuda):
    from sklearn.ensemble import RandomForestClassifier
    from oracle import patch_cache as pc
    from sduss_b200 import ops
    rng = np.random.RandomState(0)
    # a forest on [block, timestep, mse]: "recompute" when the MSE is large for its block / timestep
    X = np.stack([rng.randint(0, 24, 4000), rng.uniform(0, 1000, 4000), 10 ** rng.uniform(-6, 0, 4000)], 1)
    y = (X[:, 2] > 1e-3 * (1 + X[:, 0] / 8) * (0.5 + X[:, 1] / 1000)).astype(int)
    rf = RandomForestClassifier(n_estimators=12, max_depth=7, random_state=0).fit(X.astype(np.float32), y)
    forest = ops.DeviceForest.from_sklearn(rf, cuda)
    g = torch.Generator().manual_seed(1)
    S = [256, 1024, 512, 256]                      # four latents: 1 + 4 + 2 + 1 patches
    L, T, D = len(S), sum(S), 192
    n = T // 256
    patch_latent = torch.tensor(np.repeat(np.arange(L), [s // 256 for s in S]), dtype=torch.int32).cuda()
    t32 = torch.tensor([900.0, 500.0, 120.0, 700.0]).cuda()
    prev0 = torch.randn(T, D, generator=g).cuda().bfloat16()
    scale = (10 ** torch.linspace(-3.5, 0, n))[:, None].repeat_interleave(256, 0)   # per-patch drift
    x = (prev0.float().cpu() + scale * torch.randn(T, D, generator=g)).cuda().bfloat16()
    ws = ops.patch_mask_workspace(n, cuda)
    for block, valid_host, skipped_host in ((3, [1, 1, 1, 0], [0, 1, 2, 0, 2, 1, 0, 2]),
                                            (17, [1, 0, 1, 1], [1, 2, 2, 2, 0, 0, 1, 1])):
        prev = prev0.clone()
        valid = torch.tensor(valid_host, dtype=torch.float32).cuda()
        skipped = torch.tensor(skipped_host, dtype=torch.int32).cuda()
        mask = torch.full((n,), -1, dtype=torch.int32).cuda()
        mse = torch.zeros(n).cuda()
        for _ in range(2):   # the workspace re-arms itself: a second launch on fresh state must agree
            prev.copy_(prev0)
            skipped.copy_(torch.tensor(skipped_host, dtype=torch.int32))
            ops.patch_mask(x, prev, patch_latent, t32, valid, skipped, mask, forest, block, 2, ws, mse=mse)
            torch.cuda.synchronize()
            assert torch.equal(prev, x)                                  # the kept copy is refreshed
            lat = patch_latent.cpu().numpy()
            ok = np.asarray(valid_host)[lat] > 0
            want_mse = ((x.float() - prev0.float()) ** 2).view(n, -1).mean(1).cpu().numpy()
            got_mse = mse.cpu().numpy()
            assert np.allclose(got_mse[ok], want_mse[ok], rtol=2e-4)
            assert (got_mse[~ok] == np.float32(pc.MSE_MISSING)).all()
            feats = np.stack([np.full(n, block), t32.cpu().numpy()[lat], got_mse], 1).astype(np.float32)
            pred = rf.predict(feats)
            pred[~ok] = 1                                                # nothing kept: always computed
            want_mask, want_cnt = pc.mask_bookkeeping(list(ok), skipped_host, list(pred), 2)
            assert mask.cpu().tolist() == [int(m) for m in want_mask]
            assert skipped.cpu().tolist() == want_cnt
            assert 0 < sum(want_mask) < n                                # both outcomes occur
    # the one-node rule
    rule = ops.DeviceForest.threshold_rule(1e-3, cuda)
    prev, skipped = prev0.clone(), torch.zeros(n, dtype=torch.int32).cuda()
    ops.patch_mask(x, prev, patch_latent, t32, torch.ones(L).cuda(), skipped, mask, rule, 0, 2, ws, mse=mse)
    torch.cuda.synchronize()
    assert mask.cpu().tolist() == [int(v > 1e-3) for v in mse.cpu().tolist()]


def test_gemm_layernorm_attention_skip_clean_patches(cuda):
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(0)
    M, K, N = 2560, 256, 384                  # 10 patches of 256 rows = 10 CTA pairs x 3 column tiles
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    w = (torch.randn(N, K, generator=g) / 16).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda().bfloat16()
    resid = torch.randn(M, N, generator=g).cuda().bfloat16()
    mask = torch.tensor([1, 0, 0, 1, 1, 0, 1, 0, 0, 1], dtype=torch.int32).cuda()
    rows = mask.bool().repeat_interleave(256)
    for kw in (dict(epi=ops.EPI_BIAS, bias=bias), dict(epi=ops.EPI_GELU_TANH, bias=bias),
               dict(epi=ops.EPI_GATE_RESID, bias=bias, resid=resid)):
        full = ops.gemm(a, w, **kw)
        out = torch.full((M, N), 7.0, device=cuda, dtype=torch.bfloat16)
        ops.gemm(a, w, out, row_mask=mask, **kw)
        torch.cuda.synchronize()
        assert torch.equal(out[rows], full[rows]) and (out[~rows] == 7.0).all(), kw["epi"]
    # a short problem (no CTA pairs), everything clean: nothing is written
    out = torch.full((256, N), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.gemm(a[:256], w, out, bias=bias, row_mask=torch.zeros(1, dtype=torch.int32, device=cuda))
    assert (out == 7.0).all()
    # LayerNorm
    x = torch.randn(M, 1536, generator=g).cuda().bfloat16()
    gam, bet = torch.randn(1536, generator=g).cuda().bfloat16(), torch.randn(1536, generator=g).cuda().bfloat16()
    full = ops.layernorm_mod(x, torch.empty_like(x), eps=1e-6, gamma=gam, beta=bet)
    y = torch.full_like(x, 7.0)
    ops.layernorm_mod(x, y, eps=1e-6, gamma=gam, beta=bet, row_mask=mask)
    assert torch.equal(y[rows], full[rows]) and (y[~rows] == 7.0).all()
    # attention: two sequences (3 + 7 patches) x 3 heads, joint with a 45-token context segment
    H, D = 3, 192
    lens, ctx = [768, 1792], 45
    qkv = (torch.randn(M, 3 * D, generator=g) * 0.7).cuda().bfloat16()
    qkc = (torch.randn(2 * ctx, 3 * D, generator=g) * 0.7).cuda().bfloat16()
    seqs = [(0, 768, 0, ctx, 0, 768, 0, ctx), (768, 1792, ctx, ctx, 768, 1792, ctx, ctx)]
    plan = ops.build_attn_plan(seqs, cuda, H)

    def run(q_mask):
        oa = torch.full((M, D), 7.0, device=cuda, dtype=torch.bfloat16)
        ob = torch.full((2 * ctx, D), 7.0, device=cuda, dtype=torch.bfloat16)
        sa = ops.attn_source(q=qkv, k=qkv, k_col=D, v=qkv, v_col=2 * D, out=oa)
        sb = ops.attn_source(q=qkc, k=qkc, k_col=D, v=qkc, v_col=2 * D, out=ob)
        ops.attn_varlen(sa, sb, *plan, 0.125, q_mask=q_mask)
        torch.cuda.synchronize()
        return oa, ob
    fa, fb = run(None)
    ma, mb = run(mask)
    assert torch.equal(ma[rows], fa[rows]) and (ma[~rows] == 7.0).all()
    assert torch.equal(mb, fb)                 # context queries are always computed
    za, zb = run(torch.zeros(10, dtype=torch.int32, device=cuda))
    assert (za == 7.0).all() and torch.equal(zb, fb)


def _tiny_pipe(cuda):
    from oracle import sd3_mmdit as o3
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel
    cfg = o3.sd3_tiny_config()
    sd = {k: v.to(torch.bfloat16).float() for k, v in o3.init_sd3_weights(cfg, 0).items()}
    model = B200SD3Transformer2DModel(sd, cfg, device=cuda)
    sched = B200FlowMatchEulerDiscreteScheduler()
    return cfg, sd, model, sched, B200StableDiffusion3Pipeline(model, sched)


def test_cached_step_matches_oracle_policy_and_exact_forward(cuda):
    """Four steps of a two-request batch (256^2: 1 patch, 512^2: 4 patches per latent, CFG on) with the
    cache on. Step 0: nothing kept, everything computed -- must equal the uncached step bit for bit.
    Step 1: rule 'never' -> every patch reused. Steps 2, 3: rule 'MSE > median of what block 0 saw at step 1'
    with part of one image disturbed in between -> a mixture. Per step and request the applied
    prediction is compared with oracle/patch_cache.py run with the SAME masks (read back from the
    device). Eager launches here (the rule changes between steps; graph replays of the cached forward
    are covered by the next test)."""
    import _parity as P
    from oracle import patch_cache as pc
    from oracle import schedulers as osch
    from sduss_b200 import ops
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _tiny_pipe(cuda)
    _, _, model0, _, pipe0 = _tiny_pipe(cuda)          # the same weights, cache off
    model.use_graphs = False
    mk = lambda: make_sd3_requests(cfg, {"256": 1, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=7,
                                   latent_dtype=torch.float32)
    reqs, reqs0 = mk(), mk()
    model.enable_patch_cache(ops.DeviceForest.threshold_rule(3e38, cuda), refresh=2)
    orc = pc.CachedSD3Oracle(sd, cfg, predictor=None)
    sig, ts = osch.flow_match_sigmas(28)
    order = [("256", 0), ("256", 1), ("512", 0), ("512", 1)]     # plan order: per resolution [uncond, cond]
    n_p = {"256": 1, "512": 4}
    seen = []
    for k in range(4):
        before = P.snapshot(reqs)
        pipe.denoising_step(reqs, True, 7.0, True, 256)
        torch.cuda.synchronize()
        plan = next(iter(model._plans.values()))
        masks = plan.cache.mask.cpu().numpy().astype(bool)          # [blocks, patches]
        seen.append(masks)
        if k == 0:
            assert masks.all()
            pipe0.denoising_step(reqs0, True, 7.0, True, 256)
            torch.cuda.synchronize()
            for res in reqs:
                assert torch.equal(reqs[res][0].sampling_params.latents, reqs0[res][0].sampling_params.latents)
        col, preds = 0, {}
        for res, branch in order:
            r = reqs[res][0]
            x, kk = before[r.request_id]
            sp, po = r.sampling_params, r.prepare_output
            e = (sp.negative_prompt_embeds if branch == 0 else sp.prompt_embeds).float().cpu()
            pl_ = (po.negative_pooled_prompt_embeds if branch == 0 else po.pooled_prompt_embeds).float().cpu()
            forced = [list(masks[b, col:col + n_p[res]]) for b in range(cfg.num_layers)]
            out, _ = orc.forward_latent((r.request_id, branch), x.to(torch.bfloat16).float(), e, pl_, float(ts[kk]),
                                        forced_masks=forced)
            preds[(res, branch)] = out
            col += n_p[res]
        for res in reqs:
            r = reqs[res][0]
            ref = preds[(res, 0)] + 7.0 * (preds[(res, 1)] - preds[(res, 0)])
            cos, err, scale = P.metrics(P.implied_prediction(before, r, sig), ref)
            assert P.ok(cos, err, scale), (k, res, cos, err / scale)
        if k == 1:
            assert not masks.any()                                   # rule 'never': everything reused
            # (everything was reused, so only block 0 -- the patch embedding of the new latents -- saw a
            # non-zero MSE; its median splits the patches)
            tau = float(plan.cache.mse[0].median())
            model.enable_patch_cache(ops.DeviceForest.threshold_rule(tau, cuda), refresh=2)
        # disturb the top quarter of the 512^2 image (= its first patch) before the next step
        lat = reqs["512"][0].sampling_params.latents.clone()   # (the step's result is an inference tensor)
        lat[:, :, :16] += 0.3 * torch.randn(lat[:, :, :16].shape, device=cuda,
                                            generator=torch.Generator(device="cuda").manual_seed(k))
        reqs["512"][0].sampling_params.latents = lat
    assert 0.0 < seen[2].mean() < 1.0                   # a real mixture of recomputed and reused patches
    assert seen[2][0, 2] and seen[2][0, 6]              # the disturbed patch is recomputed (both CFG branches)


def test_cache_restarts_when_the_batch_composition_changes(cuda):
    from sduss_b200 import ops
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _tiny_pipe(cuda)
    model.enable_patch_cache(ops.DeviceForest.threshold_rule(1e9, cuda), refresh=2)   # never flags by itself
    a = make_sd3_requests(cfg, {"512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=1)
    b = make_sd3_requests(cfg, {"256": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=2)
    b["256"][0].request_id = 99

    def masks():
        torch.cuda.synchronize()
        return {len(pl.S): pl.cache.mask.cpu().numpy().copy() for pl in model._plans.values() if pl.cache is not None}
    pipe.denoising_step(a, True, 7.0, True, 256)
    assert masks()[2].all()                              # first sight: everything computed
    pipe.denoising_step(a, True, 7.0, True, 256)
    assert not masks()[2].any()                          # same slot, next step: everything reused
    pipe.denoising_step({**a, **b}, True, 7.0, True, 256)
    assert masks()[4].all()                              # new composition: the cache restarts
    pipe.denoising_step(a, True, 7.0, True, 256)
    assert masks()[2].all()                              # back, but a step was missed: recompute
    for _ in range(2):
        pipe.denoising_step(a, True, 7.0, True, 256)
        assert not masks()[2].any()
    pipe.denoising_step(a, True, 7.0, True, 256)
    assert masks()[2].all()                              # two skips in a row -> forced refresh (refresh = 2)
    assert torch.isfinite(a["512"][0].sampling_params.latents.float()).all()
