"""Step orchestration pinned against THE REFERENCE'S OWN CODE, on CPU.

tests/golden/step_{sd3,sdxl}.npz hold two consecutive steps of a mixed-resolution batch computed by
the reference's `denoising_step` methods (pipeline_stable_diffusion_3_esymred.py:232-388,
pipeline_stable_diffusion_xl_esymred.py:260-403) with the reference's scheduler classes, around the
oracle model (tools/make_golden.py). Two things are checked here without a GPU:

  1. the oracle-side step helper the GPU parity tests use (tests/_parity.py) reproduces those
     latents, i.e. the checker itself follows the reference's orchestration;
  2. the PRODUCT's host logic -- sduss_b200.pipelines._step: plan composition, gather / CFG /
     conditioning order, per-request timesteps and sigma pairs, output offsets, state side
     effects -- reproduces them when its kernels are replaced by CPU stand-ins that read the very
     descriptor structs the C ABI would receive (through their host pointers) and the model by the
     oracle forward. No product arithmetic runs here; the GPU twin is tests/test_step_golden_gpu.py.
"""
import ctypes

import numpy as np
import pytest
import torch

import _parity as P


@pytest.mark.parametrize("kind", ["sd3", "sdxl"])
@pytest.mark.parametrize("tag", ["cfg", "nocfg"])
def test_oracle_step_helper_matches_reference_orchestration(kind, tag):
    cfg, sd = P.fixture_weights(kind)
    reqs, z, g, sig, ts = P.load_step_fixture(kind, tag, embed_dtype=torch.float32)
    cfg_on = tag == "cfg"
    for rs in reqs.values():
        for r in rs:
            x = r.sampling_params.latents
            k = r.scheduler_states._step_index
            for step in (1, 2):
                if kind == "sd3":
                    pred = P.oracle_sd3_prediction(sd, cfg, r, x, ts[k], cfg_on, g, round_input=False)
                else:
                    pred = P.oracle_sdxl_prediction(sd, cfg, r, x, sig[k], ts[k], cfg_on, g, round_input=False)
                x = x + (sig[k + 1] - sig[k]) * pred
                want = torch.from_numpy(z[f"{tag}_x{step}_{r.request_id}"])
                # fp32 on both sides; the reference batches the requests through the model, the helper
                # runs them one by one (different BLAS summation order)
                assert (x - want).abs().max() <= 1e-5 * want.abs().max(), (kind, tag, r.request_id, step)
                k += 1
                assert list(z[f"{tag}_idx{step}_{r.request_id}"]) == [k, k]


# ------------------------------------------------------------------ CPU stand-ins for the C ABI
def _view(ptr, n, dtype):
    ct = {torch.float32: ctypes.c_float, torch.bfloat16: ctypes.c_uint16, torch.float16: ctypes.c_uint16}[dtype]
    a = np.ctypeslib.as_array((ct * n).from_address(ptr))
    t = torch.from_numpy(a)
    return t if dtype == torch.float32 else t.view(dtype)


class _FakeOps:
    """Same call signatures as sduss_b200.ops for the five entry points the step uses; the
    arithmetic is the documented kernel semantics in fp32 (include/sduss_b200.h)."""

    def __init__(self, real):
        self.real = real
        self.calls = []

    def latent_refs(self, rows):
        return self.real.latent_refs(rows)

    def gather_latents(self, refs, dtype, staging, scale_input=False):
        self.calls.append("gather_latents")
        for r in refs:
            x = _view(r.src, r.elems, dtype).float()
            if scale_input:
                x = x / ((torch.tensor(r.sigma) ** 2 + 1) ** 0.5)
            staging[r.off_a:r.off_a + r.elems] = x
            if r.off_b >= 0:
                staging[r.off_b:r.off_b + r.elems] = x

    def write_f32(self, dst, values):
        self.calls.append("write_f32")
        dst[:len(values)] = torch.tensor(values, dtype=torch.float32)

    def gather_rows(self, dst, srcs, bytes_each=None):
        self.calls.append("gather_rows")
        esz = dst.element_size()
        for i, p in enumerate(srcs):
            dst[i].reshape(-1)[:bytes_each // esz] = _view(p, bytes_each // esz, dst.dtype)

    def run_plan(self, model, plan, prologue=None):
        self.calls.append("run_plan")
        prologue(plan)
        model._run(plan)

    def cfg_scheduler_step(self, eps, refs, latent_dtype, guidance, cfg, mode):
        self.calls.append("cfg_scheduler_step")
        for r in refs:
            c = eps[r.off_b:r.off_b + r.elems].float()
            e = c
            if cfg:
                u = eps[r.off_a:r.off_a + r.elems].float()
                e = u + guidance * (c - u)
            x = _view(r.src, r.elems, latent_dtype).float()
            s, sn = torch.tensor(r.sigma), torch.tensor(r.sigma_next)
            if mode == 0:
                xn = x + (sn - s) * e
            else:
                assert mode == 1
                x0 = x - s * e
                xn = x + ((x - x0) / s) * (sn - s)
            _view(r.dst, r.elems, latent_dtype).copy_(xn.to(latent_dtype))


class _FakeModel:
    """plan_for / project_context / _run with the product's plan attributes, on CPU tensors; the
    forward is the oracle (raw text embeddings are 'projected' by the identity and handed to it)."""
    device = torch.device("cpu")

    def __init__(self, kind, cfg, sd):
        self.kind, self.cfg, self.sd, self.plans = kind, cfg, sd, {}

    def project_context(self, ehs):
        return ehs.clone()

    def plan_for(self, comp, ctx_len):
        from types import SimpleNamespace
        pl = self.plans.get((comp, ctx_len))
        if pl is None:
            C = self.cfg.in_channels
            L = sum(n for _, n, _, _ in comp)
            numel = [n * C * h * w for _, n, h, w in comp]
            off = np.concatenate([[0], np.cumsum(numel)])
            ctx_dim = self.cfg.joint_attention_dim if self.kind == "sd3" else self.cfg.cross_attention_dim
            pooled_dim = self.cfg.pooled_projection_dim if self.kind == "sd3" else self.cfg.pooled_dim
            pl = SimpleNamespace(
                comp=comp, ctx_len=ctx_len, L=L, flat_in=torch.zeros(int(off[-1])), flat_out=torch.zeros(int(off[-1])),
                in_elem_off={c[0]: int(o) for c, o in zip(comp, off)},
                out_elem_off={c[0]: int(o) for c, o in zip(comp, off)},
                t32=torch.zeros(L), c=torch.zeros(L * ctx_len, ctx_dim, dtype=torch.bfloat16),
                pooled=torch.zeros(L, pooled_dim, dtype=torch.bfloat16),
                text_embeds=torch.zeros(L, pooled_dim, dtype=torch.bfloat16), ids32=torch.zeros(L * 6))
            self.plans[(comp, ctx_len)] = pl
        return pl

    def kv_buffer(self, pl):
        return pl.c

    def _run(self, pl):
        C = self.cfg.in_channels
        hs, o = {}, 0
        for res, n, h, w in pl.comp:
            hs[res] = pl.flat_in[o:o + n * C * h * w].view(n, C, h, w).clone()
            o += n * C * h * w
        ctx = pl.c.float().view(pl.L, pl.ctx_len, -1)
        if self.kind == "sd3":
            from oracle import sd3_mmdit as o3
            out = o3.sd3_forward(self.sd, self.cfg, hs, ctx, pl.pooled.float(), pl.t32.clone())
        else:
            from oracle import sdxl_unet as ox
            out = ox.unet_forward(self.sd, self.cfg, hs, pl.t32.clone(), ctx, pl.text_embeds.float(),
                                  pl.ids32.view(pl.L, 6).clone())
        o = 0
        for res, n, h, w in pl.comp:
            pl.flat_out[o:o + n * C * h * w] = out[res].reshape(-1)
            o += n * C * h * w


@pytest.mark.parametrize("kind", ["sd3", "sdxl"])
@pytest.mark.parametrize("tag", ["cfg", "nocfg"])
def test_step_host_logic_matches_reference_on_cpu(kind, tag, monkeypatch):
    from sduss_b200 import ops, pipelines
    cfg, sd = P.fixture_weights(kind)
    reqs, z, g, sig, ts = P.load_step_fixture(kind, tag)
    fake = _FakeOps(ops)
    monkeypatch.setattr(pipelines, "ops", fake)
    model = _FakeModel(kind, cfg, sd)
    if kind == "sd3":
        pipe = pipelines.B200StableDiffusion3Pipeline(model, scheduler=None)
        step = lambda: pipe.denoising_step(reqs, tag == "cfg", g, True, 256)
    else:
        from types import SimpleNamespace
        pipe = pipelines.B200StableDiffusionXLPipeline(model, scheduler=SimpleNamespace(prediction_type="epsilon"))
        step = lambda: pipe.denoising_step(reqs, tag == "cfg", 0.0, g, None, {}, None, None, None, True, 256)
    # the step validates that latents live on the GPU; this CPU harness lifts exactly that check
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    for k in (1, 2):
        step()
        for rs in reqs.values():
            for r in rs:
                want = torch.from_numpy(z[f"{tag}_x{k}_{r.request_id}"])
                got = r.sampling_params.latents
                assert got.shape == want.shape and got.dtype == torch.float32
                assert (got - want).abs().max() <= 1e-5 * want.abs().max(), (kind, tag, r.request_id, k)
                assert [r.scheduler_states._step_index, r.scheduler_states.timestep_idx] == \
                    list(z[f"{tag}_idx{k}_{r.request_id}"])
    n_rows = 3 if kind == "sdxl" else 2
    # a step = the prologue launches (inside run_plan), the forward, one fused CFG + scheduler launch
    assert fake.calls == (["run_plan", "gather_latents", "write_f32"] + ["gather_rows"] * n_rows +
                          ["cfg_scheduler_step"]) * 2
    assert pipe._cond.hits > 0  # second step re-used the first step's projected conditioning
