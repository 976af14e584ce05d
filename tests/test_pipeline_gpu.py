"""Drop-in `denoising_step` behaviour on the GPU path: multi-step rollouts vs the oracle,
CFG on/off, changing batch compositions (plan + CUDA-graph cache), request state side effects."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _sd3(seed=0):
    from oracle import sd3_mmdit as o3
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel
    cfg = o3.sd3_tiny_config()
    sd = {k: v.to(torch.bfloat16).float() for k, v in o3.init_sd3_weights(cfg, seed).items()}
    model = B200SD3Transformer2DModel(sd, cfg, device="cuda")
    sched = B200FlowMatchEulerDiscreteScheduler()
    return cfg, sd, model, sched, B200StableDiffusion3Pipeline(model, sched)


def _oracle_sd3_step(sd, cfg, r, x, sig, ts, k, cfg_on, g):
    from oracle import schedulers as osch
    from oracle import sd3_mmdit as o3
    f = lambda t: t.float().cpu()
    if cfg_on:
        ehs = torch.cat([f(r.sampling_params.negative_prompt_embeds), f(r.sampling_params.prompt_embeds)])
        pooled = torch.cat([f(r.prepare_output.negative_pooled_prompt_embeds), f(r.prepare_output.pooled_prompt_embeds)])
        out = o3.sd3_forward(sd, cfg, {"x": torch.cat([x, x])}, ehs, pooled, ts[k:k + 1].repeat(2))["x"]
        eps = osch.cfg_combine(out, g)
    else:
        eps = o3.sd3_forward(sd, cfg, {"x": x}, f(r.sampling_params.prompt_embeds),
                             f(r.prepare_output.pooled_prompt_embeds), ts[k:k + 1])["x"]
    return osch.flow_match_batch_step(eps, x, sig[k:k + 1], sig[k + 1:k + 2])


@pytest.mark.parametrize("cfg_on", [True, False])
def test_sd3_three_step_rollout(cuda, cfg_on):
    from oracle import schedulers as osch
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _sd3()
    reqs = make_sd3_requests(cfg, {"256": 2, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=3)
    flat = [r for rs in reqs.values() for r in rs]
    ref = [r.sampling_params.latents.float().cpu() for r in flat]
    sig, ts = osch.flow_match_sigmas(28)
    for k in range(3):
        pipe.denoising_step(reqs, cfg_on, 7.0, True, 256)
        ref = [_oracle_sd3_step(sd, cfg, r, x, sig, ts, k, cfg_on, 7.0) for r, x in zip(flat, ref)]
    torch.cuda.synchronize()
    for r, x in zip(flat, ref):
        got = r.sampling_params.latents.float().cpu()
        assert got.shape == x.shape and r.sampling_params.latents.dtype == torch.bfloat16
        cos = torch.nn.functional.cosine_similarity(got.flatten(), x.flatten(), dim=0).item()
        assert cos > 0.998, cos  # three bf16 steps vs fp32 oracle
        assert r.scheduler_states._step_index == 3 and r.scheduler_states.timestep_idx == 3


def test_changing_batch_composition_reuses_plans(cuda):
    """Requests join and leave between steps (what sduss' scheduler does): every composition gets
    its own plan / CUDA graph, and a request's trajectory does not depend on its batch mates.
    Repeated with fresh requests and a churned allocator: the steps are asynchronous graph
    replays, so any host staging that is reused across calls shows up here as a mismatch."""
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _sd3()
    for it in range(6):
        junk = [torch.full((1 << 20,), float("nan"), device=cuda, dtype=torch.bfloat16) for _ in range(4)]
        del junk
        a = make_sd3_requests(cfg, {"256": 1, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=5 + it)
        b = make_sd3_requests(cfg, {"256": 1, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=5 + it)
        extra = make_sd3_requests(cfg, {"768": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=60 + it)
        # a: alone for three steps; b: same requests, but a 768 request joins for step 2 only
        for _ in range(3):
            pipe.denoising_step(a, True, 7.0, True, 256)
        pipe.denoising_step(b, True, 7.0, True, 256)
        pipe.denoising_step({**b, **extra}, True, 7.0, True, 256)
        pipe.denoising_step(b, True, 7.0, True, 256)
        torch.cuda.synchronize()
        assert len(model._plans) == 2
        for res in a:
            assert torch.equal(a[res][0].sampling_params.latents, b[res][0].sampling_params.latents), (it, res)
            assert a[res][0].scheduler_states._step_index == 3


def test_sdxl_three_step_rollout_cfg_off(cuda):
    from dataclasses import asdict
    from oracle import schedulers as osch
    from oracle import sdxl_unet as ox
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline
    from sduss_b200.schedulers import B200EulerDiscreteScheduler
    from sduss_b200.synthetic import make_sdxl_requests
    from sduss_b200.unet import B200UNet, UNetConfig
    oc = ox.sdxl_tiny_config()
    d = asdict(oc); d.pop("context_len")
    sd = {k: v.to(torch.bfloat16).float() for k, v in ox.init_unet_weights(oc, 1).items()}
    model = B200UNet(sd, UNetConfig(**d), device="cuda")
    sched = B200EulerDiscreteScheduler()
    pipe = B200StableDiffusionXLPipeline(model, sched)
    reqs = make_sdxl_requests(oc, {"256": 1, "768": 1}, 50, sched, cuda, seed=2)
    flat = [r for rs in reqs.values() for r in rs]
    ref = [r.sampling_params.latents.float().cpu() for r in flat]
    sig, ts, _ = osch.euler_sigmas(50)
    f = lambda t: t.float().cpu()
    for k in range(3):
        pipe.denoising_step(reqs, False, 0.0, 5.0, None, {}, None, None, None, True, 256)
        nxt = []
        for r, x in zip(flat, ref):
            xin = osch.batch_scale_model_input(x.to(torch.bfloat16), [sig[k]]).float()
            out = ox.unet_forward(sd, oc, {"x": xin}, ts[k:k + 1], f(r.sampling_params.prompt_embeds),
                                  f(r.prepare_output.pooled_prompt_embeds), f(r.prepare_output.add_time_ids))["x"]
            nxt.append(osch.euler_batch_step(out, x, [sig[k]], [sig[k + 1]]))
        ref = nxt
    torch.cuda.synchronize()
    for r, x in zip(flat, ref):
        got = f(r.sampling_params.latents)
        cos = torch.nn.functional.cosine_similarity(got.flatten(), x.flatten(), dim=0).item()
        assert cos > 0.998, cos
        assert r.scheduler_states._step_index == 3


def test_unsupported_arguments_raise(cuda):
    cfg, sd, model, sched, pipe = _sd3()
    x = {"256": torch.zeros(1, 16, 32, 32, device=cuda, dtype=torch.bfloat16)}
    e = torch.zeros(1, cfg.context_len, cfg.joint_attention_dim, device=cuda, dtype=torch.bfloat16)
    p = torch.zeros(1, cfg.pooled_projection_dim, device=cuda, dtype=torch.bfloat16)
    t = torch.zeros(1, device=cuda)
    with pytest.raises(AssertionError):
        model(x, e, p, t, skip_layers=[1])
    with pytest.raises(AssertionError):
        model(x, e, p, t, joint_attention_kwargs={"scale": 0.5})


def test_plan_cache_eviction_keeps_results(cuda):
    """A serving run sees more batch compositions than fit in memory: plans (and their CUDA graphs)
    are dropped and rebuilt. The trajectory of a request must not depend on that."""
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _sd3()
    specs = [{"256": 1, "512": 1}, {"256": 2}, {"512": 1, "768": 1}, {"256": 1, "512": 1}, {"256": 2}]

    def rollout():
        reqs = [make_sd3_requests(cfg, s, 28, sched, cuda, ctx_len=cfg.context_len, seed=20 + i)
                for i, s in enumerate(specs)]
        for _ in range(2):
            for r in reqs:
                pipe.denoising_step(r, True, 7.0, True, 256)
        torch.cuda.synchronize()
        return [r[res][0].sampling_params.latents.clone() for r in reqs for res in r]

    ref = rollout()
    assert model._plans.evictions == 0 and len(model._plans) == 3
    model._plans.clear()
    model._plans.budget_bytes = 1          # every new composition evicts all cached plans
    got = rollout()
    assert model._plans.evictions > 0
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
