"""`denoising_step` on the B200 path -- the drop-in boundary sduss' runner calls
(`self.pipeline.denoising_step(**input_dict)`, sduss/worker/runner/_model_runner.py:250-252).

Mirrors, argument for argument and side effect for side effect:
  ESyMReDStableDiffusion3Pipeline.denoising_step
      sduss/model_executor/diffusers/pipelines/stable_diffusion_3/pipeline_stable_diffusion_3_esymred.py:232-388
  ESyMReDStableDiffusionXLPipeline.denoising_step
      sduss/model_executor/diffusers/pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl_esymred.py:260-403
i.e. resolutions in ascending order, per resolution [uncond..., cond...] under CFG, the model
forward on the dict of latents, CFG combine, per-request scheduler update, then
`req.sampling_params.latents` / `req.scheduler_states` are advanced in place.

Requests are duck-typed like sduss' RunnerRequest (worker/runner/wrappers.py:19-36):
  req.request_id, req.sampling_params.{latents, prompt_embeds, negative_prompt_embeds},
  req.prepare_output.{pooled_prompt_embeds, negative_pooled_prompt_embeds[, add_time_ids,
  negative_add_time_ids]}, req.scheduler_states (see schedulers.SchedulerStates).
"""
from typing import Dict, List

import numpy as np
import torch

from . import ops


def _next_timestep(st) -> float:
    ts = getattr(st, "_timesteps_host", None)
    if ts is None:  # one device->host read per request lifetime, not per step
        ts = st._timesteps_host = st.timesteps.detach().float().cpu().numpy()
    return float(ts[st.timestep_idx])


class B200PipelineOutput:
    """Stands in for StableDiffusion{XL,3}EsymredPipelineOutput (images, nsfw_content_detected)."""

    def __init__(self, images, nsfw_content_detected=None):
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected


class _StepState:
    """Per batch-composition host staging for the fused CFG + scheduler kernel."""

    def __init__(self, plan, comp, elems_per_latent: Dict[str, int], cfg: bool, device):
        R = sum(n // (2 if cfg else 1) for _, n, _, _ in comp)
        desc, x_off = [], 0
        for res, n, _, _ in comp:
            per = n // (2 if cfg else 1)
            e = elems_per_latent[res]
            base = plan.out_elem_off[res]
            for i in range(per):
                u_off = base + i * e
                c_off = base + (per + i) * e if cfg else u_off
                desc.append((x_off, e, u_off, c_off))
                x_off += e
        self.R, self.total = R, x_off
        self.max_elems = max(d[1] for d in desc)
        self.desc = torch.tensor(desc, dtype=torch.int64).to(device)
        self.device = device
        self.spans = [(d[0], d[1]) for d in desc]


class B200DenoisingPipelineBase:
    SUPPORT_MIXED_PRECISION = True
    SUPPORT_RESOLUTIONS = [256, 512, 768, 1024]
    step_mode = 0           # b200_cfg_scheduler_step mode
    default_guidance = 7.0

    def __init__(self, model, scheduler, vae=None):
        self.model = model
        self.scheduler = scheduler
        self.vae = vae  # optional sduss_b200.vae.B200VAEDecoder (post_inference, row f-4)

    # -- post stage --------------------------------------------------------------------
    @torch.inference_mode()
    def post_inference(self, worker_reqs: Dict[str, List], output_type: str = "pt") -> None:
        """VAE decode of the finished requests' latents; sets `req.output.images` like
        ESyMReDStableDiffusionXLPipeline.post_inference (pipeline_stable_diffusion_xl_esymred.py:
        406-462) and its SD3 twin (pipeline_stable_diffusion_3_esymred.py:391-415), but for ALL
        resolutions of the batch in one pass instead of one `vae.decode` per resolution.
        output_type "pt": float image [3, H, W] in [0, 1]; "np": [H, W, 3]; "pil": PIL.Image."""
        if self.vae is None:
            raise RuntimeError("post_inference needs a B200VAEDecoder (pipeline built without vae)")
        if output_type == "latent":
            raise NotImplementedError("latent output is not supported (as in the reference)")
        res_list = self._sorted_res(worker_reqs)
        lat = {res: torch.cat([r.sampling_params.latents for r in worker_reqs[res]], dim=0) for res in res_list}
        images = self.vae.decode(lat, _borrow=True)
        for res in res_list:
            img = (images[res].float() / 2 + 0.5).clamp(0, 1)  # VaeImageProcessor.denormalize
            for i, req in enumerate(worker_reqs[res]):
                one = img[i]
                if output_type in ("np", "pil"):
                    one = one.permute(1, 2, 0).cpu().numpy()
                if output_type == "pil":
                    from PIL import Image
                    one = Image.fromarray((one * 255).round().astype("uint8"))
                req.output = B200PipelineOutput(images=one)

    # -- helpers -----------------------------------------------------------------------
    @staticmethod
    def _sorted_res(reqs: Dict[str, List]) -> List[str]:
        return sorted((r for r in reqs if len(reqs[r]) > 0), key=lambda s: int(s))

    def _state(self, plan, cfg: bool):
        # lives on the plan, so it goes away with it when the plan cache evicts the plan
        states = plan.__dict__.setdefault("_step_states", {})
        st = states.get(cfg)
        if st is None:
            elems = {res: t[0].numel() for res, t in plan.stage_out.items()}
            st = states[cfg] = _StepState(plan, plan.comp, elems, cfg, self.model.device)
        return st

    def _finish(self, plan, reqs_sorted, cfg: bool, guidance: float):
        """CFG combine + scheduler update in one launch; writes latents / states back."""
        st = self._state(plan, cfg)
        flat = [r for _, rs in reqs_sorted for r in rs]
        x = torch.cat([r.sampling_params.latents.reshape(-1) for r in flat]).to(torch.bfloat16)
        # (sigma, sigma_next) per request. A fresh host tensor per call: the step is asynchronous
        # (CUDA-graph replay), so a reused pinned staging buffer would be overwritten by the next
        # call before this call's copy has run.
        sig = np.empty((st.R, 2), dtype=np.float32)
        for i, r in enumerate(flat):
            ss = r.scheduler_states
            sig[i, 0] = float(ss.sigmas[ss._step_index])
            sig[i, 1] = float(ss.sigmas[ss._step_index + 1])
        sig_dev = torch.from_numpy(sig).to(st.device)
        out = torch.empty_like(x)
        ops.cfg_scheduler_step(plan.flat_out, x, out, st.desc, sig_dev, st.R, st.max_elems,
                               guidance, cfg, self.step_mode)
        for (off, n), r in zip(st.spans, flat):
            ss = r.scheduler_states
            ss._step_index += 1
            ss.update_states_one_step()
            lat = r.sampling_params.latents
            r.sampling_params.latents = out[off:off + n].view(lat.shape).to(lat.dtype)


class B200StableDiffusion3Pipeline(B200DenoisingPipelineBase):
    """SD3 / SD3.5: flow-match Euler, guidance 7.0 (…_3_esymred_utils.py:169)."""
    step_mode = 0
    default_guidance = 7.0

    @property
    def transformer(self):
        return self.model

    @torch.inference_mode()
    def denoising_step(self, runner_reqs: Dict[str, List], do_classifier_free_guidance: bool = True,
                       guidance_scale: float = 7.0, is_sliced: bool = True, patch_size: int = 256) -> None:
        res_list = self._sorted_res(runner_reqs)
        cfg = do_classifier_free_guidance
        lat_in, embeds, pooled, ts = {}, [], [], []
        for res in res_list:
            reqs = runner_reqs[res]
            lat = torch.cat([r.sampling_params.latents for r in reqs], dim=0)
            t = [_next_timestep(r.scheduler_states) for r in reqs]
            if cfg:
                lat_in[res] = torch.cat([lat, lat], dim=0)
                embeds += [r.sampling_params.negative_prompt_embeds for r in reqs]
                embeds += [r.sampling_params.prompt_embeds for r in reqs]
                pooled += [r.prepare_output.negative_pooled_prompt_embeds for r in reqs]
                pooled += [r.prepare_output.pooled_prompt_embeds for r in reqs]
                ts += t + t
            else:
                lat_in[res] = lat
                embeds += [r.sampling_params.prompt_embeds for r in reqs]
                pooled += [r.prepare_output.pooled_prompt_embeds for r in reqs]
                ts += t
        ids = {res: [str(r.request_id) for r in runner_reqs[res]] for res in res_list}
        self.model(hidden_states=lat_in, timestep=torch.tensor(ts, dtype=torch.float32).to(self.model.device, non_blocking=True),
                   encoder_hidden_states=torch.cat(embeds, dim=0),
                   pooled_projections=torch.cat(pooled, dim=0), return_dict=False,
                   is_sliced=is_sliced, patch_size=patch_size, input_indices=ids, _borrow=True)
        plan = self.model._plan(lat_in, embeds[0].shape[1])
        self._finish(plan, [(res, runner_reqs[res]) for res in res_list], cfg, guidance_scale)


class B200StableDiffusionXLPipeline(B200DenoisingPipelineBase):
    """SDXL-base: Euler discrete (epsilon), guidance 5.0 (…_xl_esymred_utils.py:197)."""
    step_mode = 1
    default_guidance = 5.0

    @property
    def unet(self):
        return self.model

    @torch.inference_mode()
    def denoising_step(self, worker_reqs: Dict[str, List], do_classifier_free_guidance: bool = True,
                       guidance_rescale: float = 0.0, guidance_scale: float = 5.0,
                       timestep_cond=None, extra_step_kwargs: Dict = None,
                       cross_attention_kwargs=None, ip_adapter_image=None,
                       ip_adapter_image_embeds=None, is_sliced: bool = True,
                       patch_size: int = 256) -> None:
        if guidance_rescale > 0.0:
            raise NotImplementedError("guidance_rescale > 0 is not fused (reference default is 0.0)")
        assert timestep_cond is None and cross_attention_kwargs is None
        res_list = self._sorted_res(worker_reqs)
        cfg = do_classifier_free_guidance
        lat_in, embeds, pooled, ids, ts = {}, [], [], [], []
        pt = getattr(getattr(self.scheduler, "config", None), "prediction_type",
                     getattr(self.scheduler, "prediction_type", "epsilon"))
        self.step_mode = 1 if pt == "epsilon" else 2
        for res in res_list:
            reqs = worker_reqs[res]
            lat = torch.cat([r.sampling_params.latents for r in reqs], dim=0)
            t = [_next_timestep(r.scheduler_states) for r in reqs]
            if cfg:
                lat = torch.cat([lat, lat], dim=0)
                embeds += [r.sampling_params.negative_prompt_embeds for r in reqs]
                embeds += [r.sampling_params.prompt_embeds for r in reqs]
                pooled += [r.prepare_output.negative_pooled_prompt_embeds for r in reqs]
                pooled += [r.prepare_output.pooled_prompt_embeds for r in reqs]
                for r in reqs:  # reference interleaves (neg, pos) per request here: deviation D4
                    ids += [r.prepare_output.negative_add_time_ids, r.prepare_output.add_time_ids]
                ts += t + t
            else:
                embeds += [r.sampling_params.prompt_embeds for r in reqs]
                pooled += [r.prepare_output.pooled_prompt_embeds for r in reqs]
                ids += [r.prepare_output.add_time_ids for r in reqs]
                ts += t
            # scale_model_input: x / sqrt(sigma^2 + 1) per request (one launch per resolution)
            lat_in[res] = self.scheduler.batch_scale_model_input(reqs, lat, None)
        index = {res: [str(r.request_id) for r in worker_reqs[res]] for res in res_list}
        t_dev = torch.tensor(ts, dtype=torch.float32).to(self.model.device, non_blocking=True)
        self.model(lat_in, t_dev, encoder_hidden_states=torch.cat(embeds, dim=0),
                   added_cond_kwargs={"text_embeds": torch.cat(pooled, dim=0),
                                      "time_ids": torch.cat(ids, dim=0)},
                   return_dict=False, is_sliced=is_sliced, patch_size=patch_size,
                   input_indices=index, _borrow=True)
        plan = self.model._plan(lat_in, embeds[0].shape[1])
        self._finish(plan, [(res, worker_reqs[res]) for res in res_list], cfg, guidance_scale)
