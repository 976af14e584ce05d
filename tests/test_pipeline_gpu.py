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


@pytest.mark.parametrize("cfg_on", [True, False])
def test_sd3_three_step_rollout(cuda, cfg_on):
    """Three steps of a mixed batch; after EVERY step the prediction the step applied,
    (x' - x) / (sigma' - sigma), is compared per request with the oracle's (CFG-combined)
    prediction at the same latent and timestep: cosine and max-abs (tests/_parity.py)."""
    import _parity as P
    from oracle import schedulers as osch
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _sd3()
    reqs = make_sd3_requests(cfg, {"256": 2, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=3,
                             latent_dtype=torch.float32)
    sig, ts = osch.flow_match_sigmas(28)
    for k in range(3):
        before = P.snapshot(reqs)
        pipe.denoising_step(reqs, cfg_on, 7.0, True, 256)
        torch.cuda.synchronize()
        P.check_step(reqs, before, lambda r: sig,
                     lambda r, x, kk: P.oracle_sd3_prediction(sd, cfg, r, x, ts[kk], cfg_on, 7.0),
                     label=f"sd3 step {k}")
    for rs in reqs.values():
        for r in rs:
            assert r.sampling_params.latents.dtype == torch.float32  # the latent keeps its dtype
            assert r.scheduler_states._step_index == 3 and r.scheduler_states.timestep_idx == 3


def test_parity_check_catches_a_blind_model(cuda, monkeypatch):
    """Negative controls for the step-level parity check: a step whose model output is zeroed, or
    replaced by noise, or whose CFG branches are swapped must FAIL it -- while the updated latents
    themselves still look fine (cosine > 0.999 against the oracle's x'), which is why comparing
    latents is not a parity test."""
    import _parity as P
    from oracle import schedulers as osch
    from sduss_b200 import ops
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _sd3()
    sig, ts = osch.flow_match_sigmas(28)
    real_run = ops.run_plan

    def sabotage(kind):
        def run(model_, plan, prologue=None):
            real_run(model_, plan, prologue)
            if kind == "zero":
                plan.flat_out.zero_()
            elif kind == "noise":
                plan.flat_out.copy_(torch.randn_like(plan.flat_out, dtype=torch.float32))
            elif kind == "swap":  # uncond <-> cond predictions of every resolution
                for res, n, _, _ in plan.comp:
                    o = plan.stage_out[res]
                    o.copy_(torch.cat([o[n // 2:], o[:n // 2]]).clone())
        return run

    for kind in ("zero", "noise", "swap"):
        reqs = make_sd3_requests(cfg, {"256": 1, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=9,
                                 latent_dtype=torch.float32)
        before = P.snapshot(reqs)
        monkeypatch.setattr(ops, "run_plan", sabotage(kind))
        pipe.denoising_step(reqs, True, 7.0, True, 256)
        torch.cuda.synchronize()
        monkeypatch.setattr(ops, "run_plan", real_run)
        rows = []
        P.check_step(reqs, before, lambda r: sig,
                     lambda r, x, kk: P.oracle_sd3_prediction(sd, cfg, r, x, ts[kk], True, 7.0), report=rows)
        assert rows and not any(P.ok(c, e, sc) for _, c, e, sc in rows), (kind, rows)
        # ... and the latent-level comparison the round-1 tests used does not notice:
        for rs in reqs.values():
            for r in rs:
                x, k = before[r.request_id]
                ref = x + (sig[k + 1] - sig[k]) * P.oracle_sd3_prediction(sd, cfg, r, x, ts[k], True, 7.0)
                got = r.sampling_params.latents.float().cpu()
                cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
                assert cos > 0.99, (kind, cos)
    # the unsabotaged step passes on the same requests
    reqs = make_sd3_requests(cfg, {"256": 1, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=9,
                             latent_dtype=torch.float32)
    before = P.snapshot(reqs)
    pipe.denoising_step(reqs, True, 7.0, True, 256)
    torch.cuda.synchronize()
    P.check_step(reqs, before, lambda r: sig,
                 lambda r, x, kk: P.oracle_sd3_prediction(sd, cfg, r, x, ts[kk], True, 7.0))


def test_bf16_latents_follow_the_reference_rounding(cuda):
    """With bf16 request latents (the model dtype) the step is, bit for bit, the reference's
    arithmetic on the B200 model output: bf16 CFG combine, fp32 update, bf16 result."""
    from oracle import schedulers as osch
    from sduss_b200.synthetic import make_sd3_requests
    cfg, sd, model, sched, pipe = _sd3()
    reqs = make_sd3_requests(cfg, {"256": 1, "512": 2}, 28, sched, cuda, ctx_len=cfg.context_len, seed=4)
    flat = [r for res in sorted(reqs, key=int) for r in reqs[res]]
    x0 = [r.sampling_params.latents.clone() for r in flat]
    pipe.denoising_step(reqs, True, 7.0, True, 256)
    torch.cuda.synchronize()
    plan = next(iter(model._plans.values()))
    sig, _ = osch.flow_match_sigmas(28)
    i = 0
    for res in sorted(reqs, key=int):
        out = plan.stage_out[res]
        eps = osch.cfg_combine(out.cpu(), 7.0)          # bf16 tensor ops
        for j, r in enumerate(reqs[res]):
            want = osch.flow_match_batch_step(eps[j:j + 1], x0[i].cpu(), sig[:1], sig[1:2])
            assert r.sampling_params.latents.dtype == torch.bfloat16
            assert torch.equal(r.sampling_params.latents.cpu(), want)
            i += 1


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


@pytest.mark.parametrize("cfg_on", [False, True])
def test_sdxl_three_step_rollout(cuda, cfg_on):
    import _parity as P
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
    reqs = make_sdxl_requests(oc, {"256": 1, "768": 1}, 50, sched, cuda, seed=2, latent_dtype=torch.float32)
    sig, ts, _ = osch.euler_sigmas(50)
    for k in range(3):
        before = P.snapshot(reqs)
        pipe.denoising_step(reqs, cfg_on, 0.0, 5.0, None, {}, None, None, None, True, 256)
        torch.cuda.synchronize()
        P.check_step(reqs, before, lambda r: sig,
                     lambda r, x, kk: P.oracle_sdxl_prediction(sd, oc, r, x, sig[kk], ts[kk], cfg_on, 5.0),
                     label=f"sdxl step {k}")


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
