"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the CPU
oracle would need minutes per step at these sizes):
  * batch invariance: a request stepped inside the mixed-resolution batch ends bit-identical to
    the same request stepped alone (config-2: SD3.5-medium 512^2+768^2+1024^2 with CFG;
    config-1 shape: SDXL-base 512^2+1024^2 with CFG),
  * replay determinism: the same batch stepped twice from the same state gives the same bits,
  * the step is a contraction towards the model's prediction: finite, and the scheduler state
    advances exactly once per call."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _clone_reqs(reqs):
    out = {}
    for res, rs in reqs.items():
        out[res] = []
        for r in rs:
            c = copy.copy(r)
            c.sampling_params = copy.copy(r.sampling_params)
            c.sampling_params.latents = r.sampling_params.latents.clone()
            c.scheduler_states = copy.deepcopy(r.scheduler_states)
            out[res].append(c)
    return out


def _check(pipe, reqs, step):
    mixed, again = _clone_reqs(reqs), _clone_reqs(reqs)
    step(mixed)
    step(again)
    torch.cuda.synchronize()
    for res in reqs:
        for m, a, r0 in zip(mixed[res], again[res], reqs[res]):
            assert torch.isfinite(m.sampling_params.latents.float()).all()
            assert torch.equal(m.sampling_params.latents, a.sampling_params.latents), ("replay", res)
            assert m.scheduler_states._step_index == r0.scheduler_states._step_index + 1
            assert not torch.equal(m.sampling_params.latents, r0.sampling_params.latents)
    for res in reqs:
        solo = _clone_reqs({res: reqs[res]})
        step(solo)
        torch.cuda.synchronize()
        for m, s in zip(mixed[res], solo[res]):
            assert torch.equal(m.sampling_params.latents, s.sampling_params.latents), ("batch invariance", res)


def test_sd35_medium_config2_invariance(cuda):
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel, SD3Config
    from sduss_b200.synthetic import make_sd3_requests, random_sd3_state_dict
    cfg = SD3Config()
    model = B200SD3Transformer2DModel(random_sd3_state_dict(cfg, cuda, seed=0), cfg, device=cuda)
    sched = B200FlowMatchEulerDiscreteScheduler()
    pipe = B200StableDiffusion3Pipeline(model, sched)
    reqs = make_sd3_requests(cfg, {"512": 1, "768": 1, "1024": 1}, 28, sched, cuda, seed=3)
    _check(pipe, reqs, lambda r: pipe.denoising_step(r, True, 7.0, True, 256))
    del model, pipe
    torch.cuda.empty_cache()


def test_sdxl_base_config1_invariance(cuda):
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline
    from sduss_b200.schedulers import B200EulerDiscreteScheduler
    from sduss_b200.synthetic import make_sdxl_requests, random_unet_state_dict
    from sduss_b200.unet import B200UNet, UNetConfig
    cfg = UNetConfig()
    cfg.context_len = 77
    model = B200UNet(random_unet_state_dict(cfg, cuda, seed=0), cfg, device=cuda)
    sched = B200EulerDiscreteScheduler()
    pipe = B200StableDiffusionXLPipeline(model, sched)
    reqs = make_sdxl_requests(cfg, {"512": 1, "1024": 1}, 50, sched, cuda, seed=3)
    _check(pipe, reqs, lambda r: pipe.denoising_step(r, True, 0.0, 5.0, None, {}, None, None, None, True, 256))
    del model, pipe
    torch.cuda.empty_cache()
