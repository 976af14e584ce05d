"""Step-level parity on the NOISE PREDICTION (north_star: "max-abs and cosine on the noise
prediction per step").

A denoising step only moves the latent by (sigma' - sigma) * prediction -- at step 0 of a 28-step
flow-match schedule that is 1.3 % of a unit -- so comparing the updated latents cannot see the
model: a step that does nothing scores cosine 0.9991 against the oracle's x'. These helpers recover
what the step actually applied,

    prediction = (x' - x) / (sigma' - sigma)        (flow match and Euler-epsilon alike),

from fp32 request latents (the B200 step keeps the latent's dtype, so the division is exact to
~1e-5) and compare it with the oracle's CFG-combined prediction for the same request, same step.

Tolerance (bf16 kernels + bf16 CFG arithmetic vs the fp32 oracle), per request and step:
    cosine >= COS_MIN and max-abs error <= MAXABS_FRAC * max|oracle prediction|.
The CFG combine u + g (c - u) amplifies the bf16 rounding of the two model outputs by up to
(2g - 1), which is why the max-abs bound is wider than the model-forward tests' 6 %.
"""
import torch

COS_MIN = 0.999
MAXABS_FRAC = 0.10


def snapshot(reqs):
    """{request_id: (latent fp32 cpu, step index)} before a step."""
    return {r.request_id: (r.sampling_params.latents.float().cpu().clone(), r.scheduler_states._step_index)
            for rs in reqs.values() for r in rs}


def implied_prediction(before, r, sigmas):
    x, k = before[r.request_id]
    xn = r.sampling_params.latents.float().cpu()
    dt = float(sigmas[k + 1]) - float(sigmas[k])
    return (xn - x) / dt


def metrics(pred, ref):
    cos = torch.nn.functional.cosine_similarity(pred.flatten().double(), ref.flatten().double(), dim=0).item()
    err = (pred - ref).abs().max().item()
    return cos, err, ref.abs().max().item()


def ok(cos, err, scale):
    return cos >= COS_MIN and err <= MAXABS_FRAC * scale


def _f(t):
    return t.float().cpu()


def oracle_sd3_prediction(sd, cfg, r, x, t, cfg_on, guidance, round_input=True):
    """CFG-combined MMDiT prediction of the oracle for request r at latent x, timestep t."""
    from oracle import schedulers as osch
    from oracle import sd3_mmdit as o3
    sp, po = r.sampling_params, r.prepare_output
    xin = x.to(torch.bfloat16).float() if round_input else x  # the model input is bf16 on both sides
    tt = torch.tensor([float(t)])
    if cfg_on:
        ehs = torch.cat([_f(sp.negative_prompt_embeds), _f(sp.prompt_embeds)])
        pooled = torch.cat([_f(po.negative_pooled_prompt_embeds), _f(po.pooled_prompt_embeds)])
        out = o3.sd3_forward(sd, cfg, {"x": torch.cat([xin, xin])}, ehs, pooled, tt.repeat(2))["x"]
        return osch.cfg_combine(out, guidance)
    return o3.sd3_forward(sd, cfg, {"x": xin}, _f(sp.prompt_embeds), _f(po.pooled_prompt_embeds), tt)["x"]


def oracle_sdxl_prediction(sd, oc, r, x, sigma, t, cfg_on, guidance, round_input=True):
    """CFG-combined UNet prediction of the oracle (input scaled by 1/sqrt(sigma^2+1) in the
    latent's dtype, as EulerDiscreteScheduler.batch_scale_model_input does)."""
    from oracle import schedulers as osch
    from oracle import sdxl_unet as ox
    sp, po = r.sampling_params, r.prepare_output
    xin = osch.batch_scale_model_input(x, [sigma])
    if round_input:
        xin = xin.to(torch.bfloat16).float()
    tt = torch.tensor([float(t)])
    if cfg_on:
        ehs = torch.cat([_f(sp.negative_prompt_embeds), _f(sp.prompt_embeds)])
        te = torch.cat([_f(po.negative_pooled_prompt_embeds), _f(po.pooled_prompt_embeds)])
        ids = torch.cat([_f(po.negative_add_time_ids), _f(po.add_time_ids)])
        out = ox.unet_forward(sd, oc, {"x": torch.cat([xin, xin])}, tt.repeat(2), ehs, te, ids)["x"]
        return osch.cfg_combine(out, guidance)
    return ox.unet_forward(sd, oc, {"x": xin}, tt, _f(sp.prompt_embeds), _f(po.pooled_prompt_embeds),
                           _f(po.add_time_ids))["x"]


def check_step(reqs, before, sigmas_of, oracle_pred, report=None, label=""):
    """After one denoising_step: implied prediction of every request vs oracle_pred(r, x, k).
    Returns the list of (request_id, cos, err, scale); asserts the tolerance unless `report` is a
    list (then the rows are appended and the caller decides)."""
    rows = []
    for rs in reqs.values():
        for r in rs:
            x, k = before[r.request_id]
            pred = implied_prediction(before, r, sigmas_of(r))
            ref = oracle_pred(r, x, k)
            cos, err, scale = metrics(pred, ref)
            rows.append((r.request_id, cos, err, scale))
            if report is None:
                assert ok(cos, err, scale), (label, r.request_id, k, cos, err / scale)
                assert r.scheduler_states._step_index == k + 1 and r.scheduler_states.timestep_idx == k + 1
    if report is not None:
        report.extend(rows)
    return rows


# ------------------------------------------------------------------ reference-generated step fixtures
def load_step_fixture(kind, tag, device="cpu", latent_dtype=torch.float32, embed_dtype=torch.bfloat16):
    """tests/golden/step_{sd3,sdxl}.npz (tools/make_golden.py: the reference's own denoising_step
    around the oracle model) -> (requests dict as the step takes it, fixture, guidance).
    Requests start at DIFFERENT step indices, as continuous batching produces them."""
    import os
    from types import SimpleNamespace
    import numpy as np
    from oracle import schedulers as osch
    from sduss_b200.schedulers import SchedulerStates
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"step_{kind}.npz"))
    steps = int(z["num_inference_steps"])
    if kind == "sd3":
        sig, ts = osch.flow_match_sigmas(steps)
    else:
        sig, ts, _ = osch.euler_sigmas(steps)
    ids = sorted(int(k.rsplit("_", 1)[1]) for k in z.files if k.startswith(f"{tag}_res_"))
    reqs = {}
    for i in ids:
        res = str(int(z[f"{tag}_res_{i}"]))
        t = lambda name, dt: torch.from_numpy(z[f"{tag}_{name}_{i}"]).to(device=device, dtype=dt)
        st = SchedulerStates(sig.clone(), steps, ts.clone().to(device))
        st._step_index, st.timestep_idx = (int(v) for v in z[f"{tag}_idx0_{i}"])
        po = SimpleNamespace(pooled_prompt_embeds=t("pp", embed_dtype),
                             negative_pooled_prompt_embeds=t("npp", embed_dtype))
        if kind == "sdxl":
            po.add_time_ids = t("ids", embed_dtype)
            po.negative_add_time_ids = t("ids", embed_dtype)
        sp = SimpleNamespace(latents=t("x0", latent_dtype), prompt_embeds=t("pe", embed_dtype),
                             negative_prompt_embeds=t("npe", embed_dtype), num_inference_steps=steps)
        reqs.setdefault(res, []).append(SimpleNamespace(request_id=i, sampling_params=sp, prepare_output=po,
                                                        scheduler_states=st))
    return reqs, z, float(z["guidance"]), sig, ts


def fixture_weights(kind):
    """The tiny-config weights the fixture was generated with (seeded init, bf16-representable),
    verified against the checksum stored in the fixture."""
    import os
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"step_{kind}.npz"))
    if kind == "sd3":
        from oracle import sd3_mmdit as o3
        cfg = o3.sd3_tiny_config()
        sd = {k: v.to(torch.bfloat16).float() for k, v in o3.init_sd3_weights(cfg, 0).items()}
    else:
        from oracle import sdxl_unet as ox
        cfg = ox.sdxl_tiny_config()
        sd = {k: v.to(torch.bfloat16).float() for k, v in ox.init_unet_weights(cfg, 0).items()}
    chk = float(sum(v.double().abs().sum().item() for v in sd.values()))
    want = float(z["weights_checksum"])
    assert abs(chk - want) <= 1e-9 * want, "seeded weight init differs from the one the fixture was made with"
    return cfg, sd
