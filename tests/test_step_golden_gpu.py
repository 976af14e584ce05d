"""The B200 `denoising_step` against fixtures computed by THE REFERENCE'S OWN `denoising_step`
(tests/golden/step_{sd3,sdxl}.npz, tools/make_golden.py: reference orchestration + reference
scheduler classes around the fp32 oracle model, tiny config, requests at different step indices).

Compared per request and step on the noise prediction each side applied,
(x_k - x_{k-1}) / (sigma_k - sigma_{k-1}), with the tolerance of tests/_parity.py (bf16 kernels vs
fp32): cosine >= 0.999, max-abs <= 10 % of the range; scheduler state side effects exactly."""
import pytest
import torch

import _parity as P

pytestmark = pytest.mark.gpu


def _build(kind, cfg, sd):
    if kind == "sd3":
        from sduss_b200.pipelines import B200StableDiffusion3Pipeline
        from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
        from sduss_b200.sd3_transformer import B200SD3Transformer2DModel
        model = B200SD3Transformer2DModel(sd, cfg, device="cuda")
        return B200StableDiffusion3Pipeline(model, B200FlowMatchEulerDiscreteScheduler())
    from dataclasses import asdict
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline
    from sduss_b200.schedulers import B200EulerDiscreteScheduler
    from sduss_b200.unet import B200UNet, UNetConfig
    d = asdict(cfg)
    d.pop("context_len")
    model = B200UNet(sd, UNetConfig(**d), device="cuda")
    return B200StableDiffusionXLPipeline(model, B200EulerDiscreteScheduler())


@pytest.mark.parametrize("kind", ["sd3", "sdxl"])
@pytest.mark.parametrize("tag", ["cfg", "nocfg"])
def test_b200_step_matches_reference_generated_fixture(cuda, kind, tag):
    cfg, sd = P.fixture_weights(kind)
    pipe = _build(kind, cfg, sd)
    reqs, z, g, sig, ts = P.load_step_fixture(kind, tag, device=cuda)
    cfg_on = tag == "cfg"
    prev = {r.request_id: r.sampling_params.latents.float().cpu() for rs in reqs.values() for r in rs}
    for k in (1, 2):
        if kind == "sd3":
            pipe.denoising_step(reqs, cfg_on, g, True, 256)
        else:
            pipe.denoising_step(reqs, cfg_on, 0.0, g, None, {}, None, None, None, True, 256)
        torch.cuda.synchronize()
        for rs in reqs.values():
            for r in rs:
                i = r.request_id
                idx = [int(v) for v in z[f"{tag}_idx{k}_{i}"]]
                assert [r.scheduler_states._step_index, r.scheduler_states.timestep_idx] == idx
                dt = float(sig[idx[0]]) - float(sig[idx[0] - 1])
                got = r.sampling_params.latents.float().cpu()
                pred = (got - prev[i]) / dt
                ref = (torch.from_numpy(z[f"{tag}_x{k}_{i}"]) - torch.from_numpy(z[f"{tag}_x{k - 1}_{i}"])) / dt
                cos, err, scale = P.metrics(pred, ref)
                assert P.ok(cos, err, scale), (kind, tag, i, k, cos, err / scale)
                prev[i] = got
