"""Batched per-request schedulers, B200 side.

Mirrors the reference's operator interface for this path -- same method names, argument
meaning, return values and side effects:
  EulerDiscreteScheduler.batch_set_timesteps / batch_scale_model_input / batch_step
      sduss/model_executor/diffusers/schedulers/scheduling_euler_discrete.py:72-274
  FlowMatchEulerDiscreteScheduler.batch_set_timesteps / batch_step
      sduss/model_executor/diffusers/schedulers/scheduling_flow_match_euler_discrete.py:70-203
The arithmetic runs in one launch of b200_euler_scale_input / b200_cfg_scheduler_step over
all requests instead of ~10 ATen launches + 2 host->device copies per resolution.

In a sduss deployment the `Batch*Mixin` classes are mixed into the reference's scheduler
classes (which inherit set_timesteps from diffusers); the standalone classes below carry
their own sigma tables (diffusers 0.32.1 set_timesteps semantics) so that tests and bench.py
run without diffusers.
"""
from typing import List

import numpy as np
import torch

from . import ops


class SchedulerStates:
    """Same fields/methods as the reference's *SchedulerStates wrappers
    (scheduling_euler_discrete.py:14-66)."""

    def __init__(self, sigmas, num_inference_steps, timesteps):
        self.sigmas = sigmas                # CPU fp32 [steps + 1]
        self.num_inference_steps = num_inference_steps
        self.timesteps = timesteps          # [steps]
        self._step_index = 0
        self.timestep_idx = 0

    def update_states_one_step(self):
        self.timestep_idx += 1
        assert self.timestep_idx <= self.timesteps.shape[0]

    def get_next_timestep(self):
        return self.timesteps[self.timestep_idx]

    def get_step_idx(self):
        assert self.timestep_idx == self._step_index
        return self.timestep_idx

    def to_device(self, device):
        self.sigmas = self.sigmas.to("cpu")
        self.timesteps = self.timesteps.to(device=device)


def _sigma_pairs(reqs):
    rows = []
    for r in reqs:
        st = r.scheduler_states
        rows.append((float(st.sigmas[st._step_index]), float(st.sigmas[st._step_index + 1])))
    return rows


class BatchStepMixin:
    """batch_step via the fused kernel. `mode`: 0 flow match, 1 Euler epsilon, 2 Euler v.
    The sigma pairs travel by value in the kernel arguments (no torch.tensor(...).to(device))."""

    _mode = 0

    def _step_mode(self) -> int:
        return self._mode

    def _batch_step(self, reqs, model_outputs, samples):
        assert model_outputs.shape == samples.shape and model_outputs.shape[0] == len(reqs)
        R, n = samples.shape[0], samples[0].numel()
        x = samples.contiguous()
        e = model_outputs.contiguous()
        if e.dtype not in (torch.bfloat16, torch.float32):
            e = e.float()  # fp16 model outputs: exact upcast; no CFG arithmetic happens here
        out = torch.empty_like(x)
        xs, es = x.element_size(), n
        refs = ops.latent_refs((x.data_ptr() + i * n * xs, out.data_ptr() + i * n * xs, n, -1, i * es, s, sn)
                               for i, (s, sn) in enumerate(_sigma_pairs(reqs)))
        ops.cfg_scheduler_step(e, refs, x.dtype, 1.0, False, self._step_mode())
        for r in reqs:
            r.scheduler_states._step_index += 1
        return out.to(model_outputs.dtype)  # "cast sample back to model compatible dtype"


class BatchEulerMixin(BatchStepMixin):
    _mode = 1
    prediction_type = "epsilon"

    def _step_mode(self):
        pt = getattr(getattr(self, "config", None), "prediction_type", self.prediction_type)
        if pt == "epsilon":
            return 1
        if pt == "v_prediction":
            return 2
        raise NotImplementedError(f"prediction_type {pt} is not supported by the fused step kernel")

    def batch_scale_model_input(self, worker_reqs, samples: torch.Tensor, timestep_list=None):
        sig = [float(r.scheduler_states.sigmas[r.scheduler_states._step_index]) for r in worker_reqs]
        if samples.shape[0] == len(sig) * 2:  # classifier free
            sig = sig + sig
        assert samples.shape[0] == len(sig)
        n = samples[0].numel()
        x = samples.contiguous()
        y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)  # the model's input dtype
        xs = x.element_size()
        refs = ops.latent_refs((x.data_ptr() + i * n * xs, 0, n, i * n, -1, s, 0.0)
                               for i, s in enumerate(sig))
        ops.gather_latents(refs, x.dtype, y.view(-1), scale_input=True)
        return y.to(samples.dtype)

    def batch_step(self, worker_reqs, model_outputs, timestep_list, samples, s_churn=0.0,
                   s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, generator=None, return_dict=True):
        if s_churn != 0.0 or s_tmin != 0.0 or s_tmax != float("inf") or s_noise != 1.0 or generator is not None:
            raise NotImplementedError("We do not support custom parameters at this time.")
        return self._batch_step(worker_reqs, model_outputs, samples)


class BatchFlowMatchMixin(BatchStepMixin):
    _mode = 0

    def batch_step(self, runner_reqs, model_outputs, samples, timesteps=None, s_churn=0.0,
                   s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, generator=None, return_dict=True):
        return self._batch_step(runner_reqs, model_outputs, samples)


def _group_by_steps(reqs):
    groups = {}
    for r in sorted(reqs, key=lambda q: q.sampling_params.num_inference_steps):
        groups.setdefault(r.sampling_params.num_inference_steps, []).append(r)
    return groups


class B200EulerDiscreteScheduler(BatchEulerMixin):
    """Standalone SDXL scheduler: scaled-linear betas 0.00085..0.012, 'leading' spacing,
    steps_offset 1, epsilon prediction (SDXL-base scheduler_config.json)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 steps_offset=1, prediction_type="epsilon"):
        self.prediction_type = prediction_type
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                               dtype=torch.float32) ** 2
        ac = torch.cumprod(1.0 - betas, dim=0)
        self._sig_all = (((1 - ac) / ac) ** 0.5).numpy()
        self.init_noise_sigma = None

    def set_timesteps(self, num_inference_steps, device=None):
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32)
        ts += self.steps_offset
        sig = np.interp(ts, np.arange(0, len(self._sig_all)), self._sig_all)
        sig = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig)
        self.timesteps = torch.from_numpy(ts).to(device=device)
        self.num_inference_steps = num_inference_steps
        self.init_noise_sigma = float((sig.max() ** 2 + 1) ** 0.5)

    def batch_set_timesteps(self, worker_reqs, device):
        for steps, group in _group_by_steps(worker_reqs).items():
            self.set_timesteps(steps, device=device)
            for r in group:
                r.scheduler_states = SchedulerStates(self.sigmas, steps, self.timesteps)


class B200FlowMatchEulerDiscreteScheduler(BatchFlowMatchMixin):
    """Standalone SD3.5 scheduler: shift 3.0, no dynamic shifting (deviation D7)."""

    def __init__(self, num_train_timesteps=1000, shift=3.0):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        t = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        s = t / num_train_timesteps
        s = shift * s / (1 + (shift - 1) * s)
        self.sigma_max, self.sigma_min = float(s[0]), float(s[-1])

    def set_timesteps(self, num_inference_steps, device=None):
        n = self.num_train_timesteps
        ts = np.linspace(self.sigma_max * n, self.sigma_min * n, num_inference_steps)
        sig = ts / n
        sig = self.shift * sig / (1 + (self.shift - 1) * sig)
        sig = torch.from_numpy(sig).to(torch.float32)
        self.timesteps = (sig * n).to(device=device)
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.num_inference_steps = num_inference_steps

    def batch_set_timesteps(self, runner_reqs, device):
        for steps, group in _group_by_steps(runner_reqs).items():
            self.set_timesteps(steps, device=device)
            for r in group:
                r.scheduler_states = SchedulerStates(self.sigmas, steps, self.timesteps)
