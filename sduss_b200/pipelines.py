"""`denoising_step` on the B200 path -- the drop-in boundary sduss' runner calls
(`self.pipeline.denoising_step(**input_dict)`, sduss/worker/runner/_model_runner.py:250-252).

Mirrors, argument for argument and side effect for side effect:
  ESyMReDStableDiffusion3Pipeline.denoising_step
      sduss/model_executor/diffusers/pipelines/stable_diffusion_3/pipeline_stable_diffusion_3_esymred.py:232-388
  ESyMReDStableDiffusionXLPipeline.denoising_step
      sduss/model_executor/diffusers/pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl_esymred.py:260-403
i.e. resolutions in ascending order, per resolution [uncond..., cond...] under CFG, the model
forward on the latents, CFG combine, per-request scheduler update, then
`req.sampling_params.latents` / `req.scheduler_states` are advanced in place.

What is different from the reference is how the step talks to the GPU. The reference re-`cat`s
latents, prompt embeddings and pooled embeddings of all requests every step, builds the timestep
and sigma tensors on the host and copies them over (a pageable host->device copy drains the
stream), and recomputes everything that depends only on the prompt. Here one step is:

    b200_gather_latents      request latents -> the plan's bf16 input buffer (+ CFG duplicate,
                             + Euler input scaling), descriptors passed BY VALUE
    b200_write_f32           the per-latent timesteps, by value
    b200_gather_rows (x2-3)  the requests' cached conditioning -> the plan's packed buffers
    <CUDA-graph replay of the model forward>
    b200_cfg_scheduler_step  CFG combine + scheduler update, sigma pairs by value, new latents
                             written in the latents' own dtype

No torch op, no host->device copy, no synchronisation. The prompt-only work (SD3: context_embedder
on the 333 x 4096 embeddings; SDXL: text K/V of all 70 cross-attention layers) is done when a
request is first seen and cached for its remaining steps (`_CondCache`).

Requests are duck-typed like sduss' RunnerRequest (worker/runner/wrappers.py:19-36):
  req.request_id, req.sampling_params.{latents, prompt_embeds, negative_prompt_embeds},
  req.prepare_output.{pooled_prompt_embeds, negative_pooled_prompt_embeds[, add_time_ids,
  negative_add_time_ids]}, req.scheduler_states (see schedulers.SchedulerStates).
"""
import collections
import weakref
from typing import Dict, List

import numpy as np
import torch

from . import ops


class _HostTables:
    """Host copies of each request's scheduler tables (timesteps live on the GPU in sduss:
    `to_device`, scheduling_euler_discrete.py:57-66). One device->host read per request lifetime
    instead of one per step; kept in a side table keyed by the states object (weakly), never on
    the foreign object itself (the reference pickles / `to_numpy()`s its states for IPC)."""

    def __init__(self):
        self._weak = weakref.WeakKeyDictionary()
        self._strong = collections.OrderedDict()  # for state classes that cannot be weak-referenced

    @staticmethod
    def _host(a):
        if torch.is_tensor(a):
            a = a.detach().float().cpu().numpy()
        return np.asarray(a, dtype=np.float32)

    def get(self, st):
        try:
            t = self._weak.get(st)
            if t is None:
                t = self._weak[st] = (self._host(st.timesteps), self._host(st.sigmas))
            return t
        except TypeError:
            t = self._strong.get(id(st))
            if t is None or t[0] is not st:
                t = self._strong[id(st)] = (st, self._host(st.timesteps), self._host(st.sigmas))
                while len(self._strong) > 1024:
                    self._strong.popitem(last=False)
            return t[1:]


class _CondEntry:
    __slots__ = ("refs", "versions", "ctx", "pooled", "ids")


def _version(t):
    try:
        return t._version
    except RuntimeError:   # inference tensors (made under torch.inference_mode) carry no version counter
        return -1


class _CondCache:
    """Per-(request, CFG branch) conditioning that does not change over the request's steps: the
    projected text context, the pooled embedding as bf16, SDXL's time ids as fp32. An entry is valid
    while the request still holds the very same source tensors, unmodified (weak references + the
    tensors' version counters: a recycled request id or an in-place edit is a miss)."""

    def __init__(self, max_entries=512):
        self.entries = collections.OrderedDict()
        self.max_entries = max_entries
        self.hits = self.misses = 0

    def lookup(self, key, sources):
        e = self.entries.get(key)
        if e is not None and all(r() is s for r, s in zip(e.refs, sources)) and \
                e.versions == tuple(_version(s) for s in sources):
            self.entries.move_to_end(key)
            self.hits += 1
            return e
        self.misses += 1
        return None

    def store(self, key, sources, ctx, pooled, ids):
        e = _CondEntry()
        e.refs = tuple(weakref.ref(s) for s in sources)
        e.versions = tuple(_version(s) for s in sources)
        e.ctx, e.pooled, e.ids = ctx, pooled, ids
        self.entries[key] = e
        self.entries.move_to_end(key)
        while len(self.entries) > self.max_entries:
            self.entries.popitem(last=False)
        return e


class B200PipelineOutput:
    """Stands in for StableDiffusion{XL,3}EsymredPipelineOutput (images, nsfw_content_detected)."""

    def __init__(self, images, nsfw_content_detected=None):
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected


class B200DenoisingPipelineBase:
    SUPPORT_MIXED_PRECISION = True
    SUPPORT_RESOLUTIONS = [256, 512, 768, 1024]
    step_mode = 0           # b200_cfg_scheduler_step mode
    default_guidance = 7.0
    scale_input = False     # Euler: x / sqrt(sigma^2 + 1) fused into the latent gather
    has_time_ids = False

    def __init__(self, model, scheduler, vae=None):
        self.model = model
        self.scheduler = scheduler
        self.vae = vae  # optional sduss_b200.vae.B200VAEDecoder (post_inference, row f-4)
        self._tables = _HostTables()
        self._cond = _CondCache()

    # -- prepare stage -----------------------------------------------------------------
    def attach_text_encoders(self, prompt_encoder, tokenizers):
        """prompt_encoder: sduss_b200.text_encoders.B200PromptEncoder; tokenizers: the pipeline's
        tokenizer, tokenizer_2[, tokenizer_3] (callables with the transformers tokenizer call
        signature returning an object / dict with `input_ids`)."""
        self.prompt_encoder, self.tokenizers = prompt_encoder, list(tokenizers)

    @staticmethod
    def _ids(tok, prompts, max_length):
        out = tok(prompts, padding="max_length", max_length=max_length, truncation=True, return_tensors="pt")
        return out["input_ids"] if isinstance(out, dict) else out.input_ids

    def _encode(self, prompts, max_sequence_length):
        """prompts: one list of strings per tokenizer -> (prompt_embeds, pooled)."""
        lens = [77, 77, max_sequence_length]
        ids = [self._ids(t, p, n) for t, p, n in zip(self.tokenizers, prompts, lens)]
        return self.prompt_encoder.encode(*ids)

    @torch.inference_mode()
    def _prepare(self, reqs: Dict[str, List], second_prompts, cfg: bool, generator, max_sequence_length,
                 latent_channels, init_noise_sigma_of, force_zero_negative=False, extra=None) -> None:
        """Shared body of prepare_inference: text encoders for the positive (and, under CFG, negative)
        prompts of all requests in ONE pass each, initial latents, per-request scheduler tables, and
        the per-request hand-over the reference does at the end of prepare_inference
        (pipeline_stable_diffusion_3_esymred.py:219-229, ..._xl_esymred.py:244-258)."""
        from types import SimpleNamespace
        if getattr(self, "prompt_encoder", None) is None:
            raise RuntimeError("prepare_inference needs text encoders (attach_text_encoders)")
        res_list = sorted(r for r in reqs if len(reqs[r]) > 0)     # the reference sorts the keys as strings
        flat = [r for res in res_list for r in reqs[res]]
        pick = lambda sp, names: next((getattr(sp, n) for n in names if getattr(sp, n, None) is not None), "")
        pos = [[pick(r.sampling_params, names) for r in flat] for names in second_prompts["pos"]]
        emb, pooled = self._encode(pos, max_sequence_length)
        if cfg:
            if force_zero_negative:
                nemb, npooled = torch.zeros_like(emb), torch.zeros_like(pooled)
            else:
                neg = [[pick(r.sampling_params, names) for r in flat] for names in second_prompts["neg"]]
                nemb, npooled = self._encode(neg, max_sequence_length)
        dev = self.model.device
        self.scheduler.batch_set_timesteps(flat, device=dev)
        for i, r in enumerate(flat):
            sp = r.sampling_params
            if getattr(sp, "latents", None) is None:
                h, w = sp.height // 8, sp.width // 8
                lat = torch.randn((1, latent_channels, h, w), generator=generator, device=dev, dtype=torch.float32)
                sp.latents = (lat * init_noise_sigma_of(r)).to(emb.dtype)
            elif not sp.latents.is_cuda:
                sp.latents = sp.latents.to(dev)
            sp.prompt_embeds = emb[i:i + 1]
            sp.negative_prompt_embeds = nemb[i:i + 1] if cfg else None
            r.prepare_output = SimpleNamespace(
                pooled_prompt_embeds=pooled[i:i + 1],
                negative_pooled_prompt_embeds=npooled[i:i + 1] if cfg else None, **(extra or {}))

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

    def _conditioning(self, flat, cfg):
        """[(entry_uncond | None, entry_cond)] per request, computing the misses in one batch."""
        out, todo = [], []
        for r in flat:
            sp, po = r.sampling_params, r.prepare_output
            pair = []
            for neg in ((True, False) if cfg else (False,)):
                emb = sp.negative_prompt_embeds if neg else sp.prompt_embeds
                pooled = po.negative_pooled_prompt_embeds if neg else po.pooled_prompt_embeds
                src = [emb, pooled]
                if self.has_time_ids:
                    src.append(po.negative_add_time_ids if neg else po.add_time_ids)
                key = (r.request_id, neg)
                e = self._cond.lookup(key, src)
                if e is None:
                    todo.append((key, src, len(out), len(pair)))
                pair.append(e)
            out.append(pair)
        if todo:
            dev = self.model.device
            raw = torch.cat([s[0].reshape(1, s[0].shape[-2], s[0].shape[-1]) for _, s, _, _ in todo], 0)
            ctx = self.model.project_context(raw.to(device=dev, dtype=torch.bfloat16))
            for j, (key, src, i, k) in enumerate(todo):
                pooled = src[1].reshape(-1).to(device=dev, dtype=torch.bfloat16).contiguous()
                ids = src[2].reshape(-1).to(device=dev, dtype=torch.float32).contiguous() \
                    if self.has_time_ids else None
                out[i][k] = self._cond.store(key, src, ctx[j], pooled, ids)
        return out

    def _step(self, reqs: Dict[str, List], cfg: bool, guidance: float) -> None:
        model = self.model
        res_list = self._sorted_res(reqs)
        flat = [r for res in res_list for r in reqs[res]]
        lat0 = flat[0].sampling_params.latents
        dtype, dup = lat0.dtype, (2 if cfg else 1)
        comp = []
        for res in res_list:
            lat = reqs[res][0].sampling_params.latents
            comp.append((res, dup * len(reqs[res]), lat.shape[-2], lat.shape[-1]))
        plan = model.plan_for(tuple(comp), flat[0].sampling_params.prompt_embeds.shape[-2])
        conds = self._conditioning(flat, cfg)

        # per-request descriptors: where the latent goes in the model input (uncond copy, cond
        # copy), where its predictions come out, its sigma pair and its timestep
        total = sum(r.sampling_params.latents.numel() for r in flat)
        new = torch.empty((total,), device=model.device, dtype=dtype)
        esz = new.element_size()
        gather, step, ts, views = [], [], [], []
        ctx_ptrs, pooled_ptrs, id_ptrs = [], [], []
        x_off = 0
        for res in res_list:
            rs = reqs[res]
            per = len(rs)
            t_res, c_res, p_res, i_res = [], [[], []], [[], []], [[], []]
            for i, r in enumerate(rs):
                lat = r.sampling_params.latents
                if lat.dtype != dtype or not lat.is_contiguous() or not lat.is_cuda:
                    raise ValueError("all request latents must be contiguous CUDA tensors of one dtype")
                n = lat.numel()
                ss = r.scheduler_states
                tsteps, sigmas = self._tables.get(ss)
                s, sn = float(sigmas[ss._step_index]), float(sigmas[ss._step_index + 1])
                a_in = plan.in_elem_off[res] + i * n
                a_out = plan.out_elem_off[res] + i * n
                gather.append((lat.data_ptr(), 0, n, a_in, (a_in + per * n) if cfg else -1, s, sn))
                step.append((lat.data_ptr(), new.data_ptr() + x_off * esz, n,
                             a_out if cfg else -1, (a_out + per * n) if cfg else a_out, s, sn))
                views.append((r, x_off, n, lat.shape))
                x_off += n
                t_res.append(float(tsteps[ss.timestep_idx]))
                for k, e in enumerate(conds[len(views) - 1]):
                    c_res[k].append(e.ctx.data_ptr())
                    p_res[k].append(e.pooled.data_ptr())
                    if self.has_time_ids:
                        i_res[k].append(e.ids.data_ptr())
            for k in range(dup):  # [uncond requests..., cond requests...] per resolution
                ts += t_res
                ctx_ptrs += c_res[k]
                pooled_ptrs += p_res[k]
                id_ptrs += i_res[k]
        g_refs, s_refs = ops.latent_refs(gather), ops.latent_refs(step)
        e0 = conds[0][0]
        ctx_bytes = e0.ctx.numel() * 2
        pooled_bytes = e0.pooled.numel() * 2

        def prologue(pl):
            ops.gather_latents(g_refs, dtype, pl.flat_in, scale_input=self.scale_input)
            ops.write_f32(pl.t32, ts)
            ops.gather_rows(self._ctx_buffer(pl), ctx_ptrs, ctx_bytes)
            ops.gather_rows(self._pooled_buffer(pl), pooled_ptrs, pooled_bytes)
            if self.has_time_ids:
                ops.gather_rows(pl.ids32.view(pl.L, -1), id_ptrs, e0.ids.numel() * 4)

        if getattr(model, "patch_cache_enabled", lambda: False)():
            # Patch cache (row f-3): a latent's kept block outputs / keys / values are usable iff this
            # plan slot served the same request, same CFG branch, at the step right before this one.
            cb = model.cache_for(plan)
            tags = []
            for res in res_list:
                rs = reqs[res]
                for k in range(dup):
                    tags += [(r.request_id, k, r.scheduler_states._step_index) for r in rs]
            valid = [1.0 if cb.tags[l] == t else 0.0 for l, t in enumerate(tags)]
            cb.tags = [(rid, k, step + 1) for rid, k, step in tags]
            base_prologue = prologue

            def prologue(pl):
                base_prologue(pl)
                ops.write_f32(cb.valid, valid)

            ops.run_plan(model, plan, prologue, run=model._run_cached, state=plan.cached_state)
        else:
            ops.run_plan(model, plan, prologue)
        ops.cfg_scheduler_step(plan.flat_out, s_refs, dtype, guidance, cfg, self.step_mode)
        for r, off, n, shape in views:
            ss = r.scheduler_states
            ss._step_index += 1
            ss.update_states_one_step()
            r.sampling_params.latents = new[off:off + n].view(shape)


class B200StableDiffusion3Pipeline(B200DenoisingPipelineBase):
    """SD3 / SD3.5: flow-match Euler, guidance 7.0 (…_3_esymred_utils.py:169)."""
    step_mode = 0
    default_guidance = 7.0

    @property
    def transformer(self):
        return self.model

    @staticmethod
    def _ctx_buffer(pl):
        return pl.c.view(pl.L, -1)

    @staticmethod
    def _pooled_buffer(pl):
        return pl.pooled

    def prepare_inference(self, runner_reqs: Dict[str, List] = None, guidance_scale: float = 7.0,
                          generator=None, pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None,
                          joint_attention_kwargs=None, clip_skip=None, max_sequence_length: int = 256,
                          skip_layer_guidance_scale: float = 2.8) -> None:
        """ESyMReDStableDiffusion3Pipeline.prepare_inference
        (pipeline_stable_diffusion_3_esymred.py:49-230): CLIP-L / CLIP-G / T5 prompt encoding
        (prompt_2 / prompt_3 default to prompt, negatives to ""), N(0, 1) latents, per-request
        flow-match tables."""
        assert clip_skip is None and not joint_attention_kwargs, "clip_skip / LoRA scale are not supported"
        cfg = guidance_scale > 1.0   # diffusers' do_classifier_free_guidance
        names = {"pos": [("prompt",), ("prompt_2", "prompt"), ("prompt_3", "prompt")],
                 "neg": [("negative_prompt",), ("negative_prompt_2", "negative_prompt"),
                         ("negative_prompt_3", "negative_prompt")]}
        self._prepare(runner_reqs, names, cfg, generator, max_sequence_length, self.model.cfg.in_channels,
                      lambda r: 1.0)

    @torch.inference_mode()
    def denoising_step(self, runner_reqs: Dict[str, List], do_classifier_free_guidance: bool = True,
                       guidance_scale: float = 7.0, is_sliced: bool = True, patch_size: int = 256) -> None:
        self._step(runner_reqs, do_classifier_free_guidance, guidance_scale)


class B200StableDiffusionXLPipeline(B200DenoisingPipelineBase):
    """SDXL-base: Euler discrete (epsilon), guidance 5.0 (…_xl_esymred_utils.py:197)."""
    step_mode = 1
    default_guidance = 5.0
    scale_input = True
    has_time_ids = True

    @property
    def unet(self):
        return self.model

    def _ctx_buffer(self, pl):
        return self.model.kv_buffer(pl).view(pl.L, -1)

    @staticmethod
    def _pooled_buffer(pl):
        return pl.text_embeds

    def prepare_inference(self, worker_reqs: Dict[str, List] = None, denoising_end=None,
                          guidance_scale: float = 5.0, eta: float = 0.0, generator=None,
                          pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, ip_adapter_image=None,
                          ip_adapter_image_embeds=None, output_type="pil", return_dict=True,
                          cross_attention_kwargs=None, guidance_rescale: float = 0.0,
                          crops_coords_top_left=(0, 0), negative_original_size=None,
                          negative_crops_coords_top_left=(0, 0), negative_target_size=None, clip_skip=None,
                          force_zeros_for_empty_prompt: bool = True) -> None:
        """ESyMReDStableDiffusionXLPipeline.prepare_inference
        (pipeline_stable_diffusion_xl_esymred.py:56-258): CLIP-L / CLIP-G prompt encoding (the reference
        passes no negative prompt: zeros under SDXL-base's force_zeros_for_empty_prompt), initial
        latents scaled by init_noise_sigma, per-request Euler tables, add_time_ids built for
        original / target size (1024, 1024) whatever the request resolution (deviation D3)."""
        assert clip_skip is None and not cross_attention_kwargs and denoising_end is None
        assert ip_adapter_image is None and ip_adapter_image_embeds is None
        cfg = guidance_scale > 1.0
        # the reference calls encode_prompt(prompt=prompt, prompt_2=None, negative_prompt=None, ...):
        # prompt_2 and the requests' negative prompts are NOT used (xl_esymred.py:118-133)
        names = {"pos": [("prompt",), ("prompt",)], "neg": [(), ()]}
        dev = self.model.device
        ids = torch.tensor([[1024., 1024., float(crops_coords_top_left[0]), float(crops_coords_top_left[1]),
                             1024., 1024.]], device=dev, dtype=torch.bfloat16)
        if negative_original_size is not None and negative_target_size is not None:
            nids = torch.tensor([[float(v) for v in (*negative_original_size, *negative_crops_coords_top_left,
                                                     *negative_target_size)]], device=dev, dtype=torch.bfloat16)
        else:
            nids = ids
        self._prepare(worker_reqs, names, cfg, generator, 77, self.model.cfg.in_channels,
                      lambda r: self.scheduler.init_noise_sigma, force_zero_negative=force_zeros_for_empty_prompt,
                      extra={"add_time_ids": ids, "negative_add_time_ids": nids, "timestep_cond": None,
                             "extra_step_kwargs": {}})

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
        pt = getattr(getattr(self.scheduler, "config", None), "prediction_type",
                     getattr(self.scheduler, "prediction_type", "epsilon"))
        if pt not in ("epsilon", "v_prediction"):
            raise NotImplementedError(f"prediction_type {pt} is not supported by the fused step kernel")
        self.step_mode = 1 if pt == "epsilon" else 2
        # add_time_ids order: the reference interleaves (neg, pos) per request while the latents are
        # blocked [uncond..., cond...] (deviation D4, harmless there because neg == pos ids); here
        # every latent gets the ids of its own branch.
        self._step(worker_reqs, do_classifier_free_guidance, guidance_scale)
