"""ORACLE (test infrastructure): the batched Euler (SDXL) and flow-match Euler (SD3) scheduler
arithmetic of sduss, restated on CPU.

In-tree reference (pinned by tests/golden/sched_*.npz, which tools/make_golden.py generates by
executing the reference's own methods):
  batch_scale_model_input  sduss/model_executor/diffusers/schedulers/scheduling_euler_discrete.py:161-184
  batch_step (Euler)       scheduling_euler_discrete.py:187-274
  batch_step (flow match)  scheduling_flow_match_euler_discrete.py:159-203
  CFG combine              pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl_esymred.py:382-385
Third-party (diffusers==0.32.1 set_timesteps, NOT in /root/reference; restated from the
published algorithm, SURVEY.md A11/A12 -- sigma tables are "parity unpinned"):
  euler_sigmas(), flow_match_sigmas().
"""
import numpy as np
import torch


def euler_sigmas(num_inference_steps: int, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, steps_offset=1):
    """EulerDiscreteScheduler.set_timesteps with SDXL's config: scaled_linear betas,
    timestep_spacing='leading', interpolation linear, final sigma zero.
    Returns (sigmas[steps+1] fp32, timesteps[steps] fp32, init_noise_sigma)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                           dtype=torch.float32) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    step_ratio = num_train_timesteps // num_inference_steps
    timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.float32)
    timesteps += steps_offset
    sig_all = (((1 - alphas_cumprod) / alphas_cumprod) ** 0.5).numpy()
    sigmas = np.interp(timesteps, np.arange(0, len(sig_all)), sig_all)
    sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
    init_noise_sigma = float((sigmas.max() ** 2 + 1) ** 0.5)
    return torch.from_numpy(sigmas), torch.from_numpy(timesteps), init_noise_sigma


def flow_match_sigmas(num_inference_steps: int, num_train_timesteps=1000, shift=3.0):
    """FlowMatchEulerDiscreteScheduler (shift=3, no dynamic shifting: deviation D7)."""
    t = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
    s = t / num_train_timesteps
    s = shift * s / (1 + (shift - 1) * s)
    sigma_max, sigma_min = float(s[0]), float(s[-1])
    ts = np.linspace(sigma_max * num_train_timesteps, sigma_min * num_train_timesteps,
                     num_inference_steps)
    sig = ts / num_train_timesteps
    sig = shift * sig / (1 + (shift - 1) * sig)
    sig = torch.from_numpy(sig).to(torch.float32)
    timesteps = sig * num_train_timesteps
    sigmas = torch.cat([sig, torch.zeros(1)])
    return sigmas, timesteps


def batch_scale_model_input(samples: torch.Tensor, sigmas) -> torch.Tensor:
    """x / sqrt(sigma^2 + 1); sigmas: per-request list, repeated x2 when samples hold CFG pairs.
    The reference builds the sigma tensor in samples.dtype (scheduling_euler_discrete.py:175)."""
    s = torch.tensor(data=list(sigmas), dtype=samples.dtype)
    if samples.shape[0] == s.shape[0] * 2:
        s = s.repeat(2)
    s = s.reshape([samples.shape[0]] + [1] * (samples.ndim - 1))
    return samples / ((s ** 2 + 1) ** 0.5)


def euler_batch_step(model_outputs, samples, sigmas, sigmas_next, prediction_type="epsilon"):
    x = samples.to(torch.float32)
    shape = [model_outputs.shape[0]] + [1] * (model_outputs.ndim - 1)
    s = torch.tensor(data=list(sigmas)).reshape(shape)
    sn = torch.tensor(data=list(sigmas_next), dtype=s.dtype).reshape(shape)
    if prediction_type == "epsilon":
        x0 = x - s * model_outputs
    elif prediction_type == "v_prediction":
        x0 = model_outputs * (-s / (s ** 2 + 1) ** 0.5) + (x / (s ** 2 + 1))
    elif prediction_type in ("sample", "original_sample"):
        x0 = model_outputs
    else:
        raise ValueError(prediction_type)
    derivative = (x - x0) / s
    prev = x + derivative * (sn - s)
    return prev.to(model_outputs.dtype)


def flow_match_batch_step(model_outputs, samples, sigmas, sigmas_next):
    x = samples.to(torch.float32)
    shape = [model_outputs.shape[0]] + [1] * (model_outputs.ndim - 1)
    s = torch.tensor(data=list(sigmas)).reshape(shape)
    sn = torch.tensor(data=list(sigmas_next)).reshape(shape)
    prev = x + (sn - s) * model_outputs
    return prev.to(model_outputs.dtype)


def cfg_combine(noise_pred: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    """[uncond..., cond...] -> uncond + g * (cond - uncond)."""
    u, c = noise_pred.chunk(2)
    return u + guidance_scale * (c - u)
