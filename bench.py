#!/usr/bin/env python
"""bench.py -- mixed-resolution denoise steps/s on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--model both|sd3|sdxl] [--serve ...]

Headline workload (N=1): BASELINE.json configs[1]: SD3.5-medium MMDiT denoising step over a mixed
batch of three requests (512^2 + 768^2 + 1024^2), CFG on (6 latents, 14 848 image tokens + 6x333
context tokens), bf16, random-init weights, synthetic latents / embeddings. One "step" = one
`denoising_step` call over the whole batch: gather, model forward, CFG combine, flow-match update,
write-back (reference: pipeline_stable_diffusion_3_esymred.py:232-388).
The same line carries an `sdxl` sub-record (BASELINE's metric names both models): configs[0]'s
shape on the GPU -- SDXL-base UNet, 512^2 + 1024^2, CFG on (4 latents) -- with its own value, e2e,
roofline (dominant kernel there: the tcgen05 GEMM) and cpu_baseline.

`value`  = steps/s with the request state resident in HBM (as in sduss, where request latents
           and embeddings live on the GPU between steps), all N GPUs summed (replicas, weak).
`e2e`    = the same call with every request tensor in pinned HOST memory: per step the latents,
           prompt embeddings and pooled embeddings go host->device and the updated latents
           come back device->host, all inside the timed region.
`roofline` = dominant kernel measured live with CUDA events around each of its launches inside
           timed steps (eager, one stream, so nothing runs beside the kernel being timed);
           `traffic` = DRAM bytes of one launch from the committed ncu capture
           (profiles/roofline_traffic.json), null when there is no capture for this kernel build.
`cpu_baseline` = the oracle (PyTorch fp32 restatement of the reference path; the reference itself
           cannot run on CPU nor without diffusers) on the host cores: ONE REAL
           step of the same workload (all requests, CFG, scheduler update), no extrapolation.
`--impl reference` = the same oracle arm as its own bench line: real steps of the headline
           workload on all host cores, as many of the requested K as fit a wall-clock budget
           (a step is ~20 s on 16 cores); `steps` / `warmup` in its line are the numbers actually
           run, the requested ones are reported beside them. Imports nothing from sduss_b200.
`--serve` = serving replay (BASELINE configs[2] / [3]): see serve_main().
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SD3_SPEC = {"512": 1, "768": 1, "1024": 1}
SDXL_SPEC = {"512": 1, "1024": 1}  # BASELINE configs[0] shape: 512^2 + 1024^2, CFG -> 4 latents
SD3_STEPS, SDXL_STEPS = 28, 50     # BASELINE configs[3] / [2]
SD3_GUIDANCE, SDXL_GUIDANCE = 7.0, 5.0
METRIC = {"sd3": "mixed-res denoise steps/s (SD3.5-medium, 512^2+768^2+1024^2, CFG)",
          "sdxl": "mixed-res denoise steps/s (SDXL-base, 512^2+1024^2, CFG)"}
WORKLOAD = {"sd3": "BASELINE configs[1]: SD3.5-medium MMDiT denoise step, mixed batch 512^2+768^2+1024^2 "
                   "(1 request each), CFG on -> 6 latents, bf16, random-init",
            "sdxl": "BASELINE configs[0] shape on B200: SDXL-base UNet denoise step, 512^2 + 1024^2 "
                    "(1 request each), CFG on -> 4 latents, bf16, random-init"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
        "fallback (B200_PROFILING.md)"


def committed_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU oracle leg
# (the only code in this file that touches oracle/; nothing here imports sduss_b200)
def _oracle_sd3_step_fn(sd32):
    from oracle import schedulers as osch
    from oracle import sd3_mmdit as o3
    cfg = o3.sd35_medium_config()
    sd = sd32 if sd32 is not None else o3.init_sd3_weights(cfg, 0)
    g = torch.Generator().manual_seed(0)
    lat = {r: torch.randn(n, 16, int(r) // 8, int(r) // 8, generator=g) for r, n in SD3_SPEC.items()}
    L = sum(SD3_SPEC.values())
    ehs = torch.randn(2 * L, 333, 4096, generator=g)      # rows: per resolution [uncond..., cond...]
    pooled = torch.randn(2 * L, 2048, generator=g)
    sig, ts = osch.flow_match_sigmas(SD3_STEPS)
    state = {"lat": lat, "k": 0}

    def step():
        k = state["k"] % SD3_STEPS
        x = state["lat"]
        out = o3.sd3_forward(sd, cfg, {r: torch.cat([v, v]) for r, v in x.items()}, ehs, pooled,
                             ts[k:k + 1].repeat(2 * L))
        new = {}
        for r, v in x.items():
            eps = osch.cfg_combine(out[r], SD3_GUIDANCE)
            new[r] = osch.flow_match_batch_step(eps, v, sig[k:k + 1].repeat(v.shape[0]),
                                                sig[k + 1:k + 2].repeat(v.shape[0]))
        state["lat"], state["k"] = new, k + 1

    fl = sum(2 * n * o3.sd3_flops_per_latent(cfg, int(r)) for r, n in SD3_SPEC.items())
    return step, fl


def _oracle_sdxl_step_fn(sd32):
    from oracle import schedulers as osch
    from oracle import sdxl_unet as ox
    cfg = ox.sdxl_base_config()
    sd = sd32 if sd32 is not None else ox.init_unet_weights(cfg, 0)
    g = torch.Generator().manual_seed(0)
    sig, ts, init = osch.euler_sigmas(SDXL_STEPS)
    lat = {r: torch.randn(n, 4, int(r) // 8, int(r) // 8, generator=g) * init for r, n in SDXL_SPEC.items()}
    L = sum(SDXL_SPEC.values())
    ehs, te = torch.randn(2 * L, 77, 2048, generator=g), torch.randn(2 * L, 1280, generator=g)
    ids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * (2 * L))
    state = {"lat": lat, "k": 0}

    def step():
        k = state["k"] % SDXL_STEPS
        x = state["lat"]
        xin = {r: osch.batch_scale_model_input(torch.cat([v, v]), [sig[k]] * v.shape[0]) for r, v in x.items()}
        out = ox.unet_forward(sd, cfg, xin, ts[k:k + 1].repeat(2 * L), ehs, te, ids)
        new = {}
        for r, v in x.items():
            eps = osch.cfg_combine(out[r], SDXL_GUIDANCE)
            new[r] = osch.euler_batch_step(eps, v, [sig[k]] * v.shape[0], [sig[k + 1]] * v.shape[0])
        state["lat"], state["k"] = new, k + 1

    fl = sum(2 * n * ox.unet_flops_per_latent(cfg, int(r)) for r, n in SDXL_SPEC.items())
    return step, fl


def cpu_oracle(model, sd32=None, steps=1, warmup=0, budget_s=None):
    """Real steps of the bench workload through the oracle on all host cores. Runs `warmup` untimed
    and up to `steps` timed steps; with a budget, stops starting new steps once it is used up (at
    least one step is always timed). Returns the cpu_baseline object + the counts actually run."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, step_fl = (_oracle_sd3_step_fn if model == "sd3" else _oracle_sdxl_step_fn)(sd32)
    t_start = time.perf_counter()
    elapsed = lambda: time.perf_counter() - t_start
    done_w, last, times = 0, None, []
    for _ in range(warmup if budget_s is None else min(warmup, 1)):  # a budgeted run affords one warm-up
        t0 = time.perf_counter()
        step()
        last = time.perf_counter() - t0
        done_w += 1
    for _ in range(steps):
        if times and budget_s is not None and elapsed() + last > budget_s:
            break
        t0 = time.perf_counter()
        step()
        last = time.perf_counter() - t0
        times.append(last)
    mean = sum(times) / len(times)
    spec = SD3_SPEC if model == "sd3" else SDXL_SPEC
    return {"value": 1.0 / mean, "unit": "denoise steps/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} full step(s) of the same workload ({'+'.join(k + '^2' for k in spec)}, CFG, "
                      f"scheduler update) through the fp32 oracle after {done_w} warm-up step(s); no extrapolation",
            "seconds_per_step": mean, "steps_timed": len(times), "warmup_run": done_w,
            "step_tflop": step_fl / 1e12, "tflops": step_fl / 1e12 / mean}


def step_tflop(model):
    """Algorithmic TFLOP of one step of the bench workload (analytic: SURVEY.md 8d formulas)."""
    if model == "sd3":
        from oracle import sd3_mmdit as o3
        cfg = o3.sd35_medium_config()
        return sum(2 * n * o3.sd3_flops_per_latent(cfg, int(r)) for r, n in SD3_SPEC.items()) / 1e12
    from oracle import sdxl_unet as ox
    cfg = ox.sdxl_base_config()
    return sum(2 * n * ox.unet_flops_per_latent(cfg, int(r)) for r, n in SDXL_SPEC.items()) / 1e12


def run_reference(args, rank, world):
    """The reference arm: the oracle on the host cores (the reference's own path needs a GPU,
    diffusers and xformers: DESIGN.md section 5). Rank 0 only."""
    if rank != 0:
        return
    budget = float(os.environ.get("SDUSS_B200_REF_BUDGET_S", "150"))
    model = "sdxl" if args.model == "sdxl" else "sd3"
    res = cpu_oracle(model, None, steps=max(1, args.steps), warmup=max(0, args.warmup), budget_s=budget)
    line = {"impl": "reference", "metric": METRIC[model],
            "value": res["value"], "unit": "denoise steps/s", "n_gpus": args.gpus,
            "steps": res["steps_timed"], "warmup": res["warmup_run"],
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1000.0 * res["seconds_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD[model].replace("bf16", "fp32 on the host cores"),
                       "arm": "oracle (fp32 restatement of the reference path) on all host cores; every step is a "
                              f"real full step of the workload; no new step starts once {budget:.0f} s of wall clock "
                              "are used (one warm-up at most), so `steps` / `warmup` are the counts actually run"},
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": "denoise steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 leg
def tensor_bytes(reqs):
    h2d = d2h = 0
    for rs in reqs.values():
        for r in rs:
            sp, po = r.sampling_params, r.prepare_output
            for t in [sp.latents, sp.prompt_embeds, sp.negative_prompt_embeds] + \
                    [v for v in vars(po).values() if torch.is_tensor(v)]:
                h2d += t.numel() * t.element_size()
            d2h += sp.latents.numel() * sp.latents.element_size()
    return h2d, d2h


def attn_flops_per_step(cfg, spec, ctx=333):
    """4*Sq*Skv*d*heads per latent per attention; 24 joint + 13 image-only."""
    H, d = cfg.num_attention_heads, cfg.attention_head_dim
    fl = 0.0
    for r, n in spec.items():
        S = (int(r) // 16) ** 2
        joint = 4.0 * (S + ctx) ** 2 * d * H
        selfa = 4.0 * S * S * d * H
        fl += 2 * n * (cfg.num_layers * joint + len(cfg.dual_attention_layers) * selfa)
    return fl


def build_pipeline(model, dev):
    from sduss_b200 import synthetic
    if model == "sd3":
        from sduss_b200.pipelines import B200StableDiffusion3Pipeline
        from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
        from sduss_b200.sd3_transformer import B200SD3Transformer2DModel, SD3Config
        cfg = SD3Config()
        sd = synthetic.random_sd3_state_dict(cfg, dev, seed=0)
        sched = B200FlowMatchEulerDiscreteScheduler()
        pipe = B200StableDiffusion3Pipeline(B200SD3Transformer2DModel(sd, cfg, device=dev), sched)
        make = lambda spec, steps, seed, pin=False: synthetic.make_sd3_requests(
            cfg, spec, steps, sched, dev, seed=seed, pin_host=pin)
        call = lambda reqs: pipe.denoising_step(reqs, True, SD3_GUIDANCE, True, 256)
    else:
        from sduss_b200.pipelines import B200StableDiffusionXLPipeline
        from sduss_b200.schedulers import B200EulerDiscreteScheduler
        from sduss_b200.unet import B200UNet, UNetConfig
        cfg = UNetConfig()
        cfg.context_len = 77
        sd = synthetic.random_unet_state_dict(cfg, dev, seed=0)
        sched = B200EulerDiscreteScheduler()
        pipe = B200StableDiffusionXLPipeline(B200UNet(sd, cfg, device=dev), sched)
        make = lambda spec, steps, seed, pin=False: synthetic.make_sdxl_requests(
            cfg, spec, steps, sched, dev, seed=seed, pin_host=pin)
        call = lambda reqs: pipe.denoising_step(reqs, True, 0.0, SDXL_GUIDANCE, None, {}, None, None, None, True, 256)
    return cfg, sd, pipe, make, call


def measure_model(model, args, rank, world, dev, dist):
    """value / e2e / roofline / launches of one model on this rank; rank 0 returns the record."""
    from sduss_b200 import ops
    cfg, sd, pipe, make, call = build_pipeline(model, dev)
    spec = SD3_SPEC if model == "sd3" else SDXL_SPEC
    horizon = 6 * (args.steps + args.warmup) + 64  # scheduler tables long enough for every call below

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps

    # ---- device-resident steps
    reqs = make(spec, horizon, rank)
    step = lambda: call(reqs)
    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    # (a) per-launch CUDA events: eager launches on one stream, so no other kernel of the step runs
    #     beside the one being timed
    ops.profile = {}
    pipe.model.two_streams = False
    n0 = ops.launch_count
    ms_prof = timed(step, args.steps, 0)
    launches = ops.launch_count - n0
    pipe.model.two_streams = True
    prof, ops.profile = ops.profile, None
    tags = prof.pop("tags", [])
    torch.cuda.synchronize()
    per_kernel = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in prof.items()}
    shapes = {}
    for n, t, ev in tags:   # per (M, N, K, epilogue | stride): ms per step, calls per step
        k = (n.replace("b200_", "").replace("_bf16", ""),) + tuple(t)
        e = shapes.setdefault(k, [0.0, 0])
        e[0] += ev[0].elapsed_time(ev[1])
        e[1] += 1
    top_shapes = [{"kernel": k[0], "shape": list(k[1:]), "ms_per_step": round(v[0] / args.steps, 3),
                   "calls_per_step": v[1] / args.steps,
                   "tflops": round(2.0 * k[1] * k[2] * k[3] * v[1] / (v[0] / 1e3) / 1e12, 1)
                   if v[0] > 0 and k[0] != "attn_varlen" else None}   # attention: (q rows, kv rows, units)
                  for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])
                  if k[0] == "attn_varlen" or v[0] / args.steps >= 0.4][:16]
    gemm_fl = sum(2.0 * t[0] * t[1] * t[2] for n, t, _ in tags if n == "b200_gemm_bf16") / args.steps
    conv_fl = sum(2.0 * t[0] * t[1] * t[2] for n, t, _ in tags if n == "b200_conv3x3_bf16") / args.steps
    # (b) the timed region proper: the step as the product runs it (CUDA-graph replay of the forward)
    ms_step = min(ms_prof, timed(step, args.steps, 2))
    clk = clocks.stop()

    # ---- e2e: every request tensor in pinned host memory, H2D + D2H inside the timed region
    hreqs = make(spec, horizon, rank, True)
    h2d, d2h = tensor_bytes(hreqs)

    def e2e_step():
        dreqs = {}
        for res, rs in hreqs.items():
            dreqs[res] = []
            for r in rs:
                sp, po = r.sampling_params, r.prepare_output
                dreqs[res].append(type(r)(
                    request_id=r.request_id, scheduler_states=r.scheduler_states,
                    sampling_params=type(sp)(**{k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v)
                                                for k, v in vars(sp).items()}),
                    prepare_output=type(po)(**{k: v.to(dev, non_blocking=True) for k, v in vars(po).items()})))
        call(dreqs)
        for res, rs in hreqs.items():
            for r, d in zip(rs, dreqs[res]):
                r.sampling_params.latents.copy_(d.sampling_params.latents, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the host owns the result before the next step

    ms_e2e = timed(e2e_step, args.steps, args.warmup)
    if rank != 0:
        return None, sd

    pk, pk_src = peaks()
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    gemm_ms, attn_ms = per_kernel.get("b200_gemm_bf16", 0.0), per_kernel.get("b200_attn_varlen_bf16", 0.0)
    if model == "sd3":
        attn_fl = attn_flops_per_step(cfg, spec)
        ach = attn_fl / (attn_ms / 1e3) / 1e12 if attn_ms > 0 else None
        roof = {"kernel": "attn_fwd_kernel (b200_attn_varlen_bf16)", "bound": "tensor", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                "traffic": committed_traffic("attn_fwd_kernel:sd3_joint"),
                "traffic_unit": "dram bytes read + written by one joint-attention launch of this workload "
                                "(ncu --set full; algorithmic Q, K, V in + O out: 207 MB)",
                "peak_source": pk_src + ", sustained bf16", "kernel_ms_per_step": attn_ms,
                "algorithmic_tflop_per_step": attn_fl / 1e12,
                "gemm": {"ms_per_step": gemm_ms, "tflop_per_step": gemm_fl / 1e12,
                         "achieved": gemm_fl / 1e12 / (gemm_ms / 1e3) if gemm_ms else None,
                         "frac": gemm_fl / 1e12 / (gemm_ms / 1e3) / peak if gemm_ms else None}}
    else:
        ach = gemm_fl / 1e12 / (gemm_ms / 1e3) if gemm_ms else None
        roof = {"kernel": "gemm_bf16_kernel (b200_gemm_bf16: transformer linears, 1x1 shortcuts, conv_in)",
                "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": (ach / peak) if ach else None,
                "traffic": committed_traffic("gemm_bf16_kernel:sdxl_2560x1280x5120"),
                "traffic_unit": "dram bytes read + written by one [2560,1280]x[5120,1280]^T launch",
                "peak_source": pk_src + ", sustained bf16", "kernel_ms_per_step": gemm_ms,
                "algorithmic_tflop_per_step": gemm_fl / 1e12,
                "conv": {"ms_per_step": per_kernel.get("b200_conv3x3_bf16", 0.0), "tflop_per_step": conv_fl / 1e12},
                "attn": {"ms_per_step": attn_ms + per_kernel.get("b200_attn_cross_short_bf16", 0.0),
                         "self_ms_per_step": attn_ms,
                         "cross_ms_per_step": per_kernel.get("b200_attn_cross_short_bf16", 0.0)}}
    rec = {"metric": METRIC[model], "value": world * 1000.0 / ms_step, "unit": "denoise steps/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": WORKLOAD[model], "requests_per_step": sum(spec.values()),
                      "guidance": SD3_GUIDANCE if model == "sd3" else SDXL_GUIDANCE,
                      "l2": "weights (5 GB bf16) and activations streamed every step exceed the 126 MB L2; no flush",
                      "parallelism": f"dp{world} replicas, no collective"},
           "req_steps_per_s": world * sum(spec.values()) * 1000.0 / ms_step,
           "e2e": {"value": world * 1000.0 / ms_e2e, "unit": "denoise steps/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
           "gpu_launches": launches, "clocks": clk, "roofline": roof,
           "kernel_ms_per_step": {k: round(v, 3) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
           "top_shapes": top_shapes}
    return rec, sd


def run_b200(args, rank, world, local_rank):
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    models = ["sd3", "sdxl"] if args.model == "both" else [args.model]
    recs = {}
    for m in models:
        rec, sd = measure_model(m, args, rank, world, dev, dist)
        if rank == 0 and world == 1 and not args.no_cpu:
            sd32 = {k: v.float().cpu() for k, v in sd.items()}
            del sd
            rec["cpu_baseline"] = cpu_oracle(m, sd32, steps=1, warmup=0)
            del sd32
        if rank == 0:
            rec["config"]["step_tflop"] = step_tflop(m)
            rec["config"]["model_tflops_per_gpu"] = rec["config"]["step_tflop"] / (rec["ms_per_step"] / 1e3)
        recs[m] = rec
        gc.collect()
        torch.cuda.empty_cache()
    serve = {}
    if not args.no_serve:
        # BASELINE configs[2] / [3]: serving replay on the N workers (tools/serve_replay.py); the
        # offered load (4 req/s per GPU) is above one B200's capacity, so req_s is the capacity.
        from tools import serve_replay as sr
        for m in models:
            w = sr._Worker(m, dev, seed=1000 * rank)
            w.warm()
            serve[m] = sr.serve_run(m, w, sr.load_trace(m, args.serve_requests * world, 4.0 * world),
                                    rank, world, dist, "bench")
            del w
            gc.collect()
            torch.cuda.empty_cache()
    if rank == 0:
        line = recs[models[0]]
        for m in models[1:]:
            line[m] = recs[m]
        if serve:
            line["serve"] = serve
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--model", default="both", choices=["both", "sd3", "sdxl"],
                    help="both = SD3.5-medium headline (BASELINE configs[1]) + `sdxl` sub-record (configs[0] shape)")
    ap.add_argument("--serve", action="store_true", help="serving replay instead of the step bench")
    ap.add_argument("--no-serve", action="store_true", help="skip the `serve` sub-records of the step bench")
    ap.add_argument("--serve-requests", type=int, default=48, help="requests per GPU of the `serve` sub-records")
    ap.add_argument("--qps", type=float, default=0.0, help="--serve: offered requests/s per GPU (default 4)")
    ap.add_argument("--requests", type=int, default=0, help="--serve: requests per GPU (default 48)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    if args.serve:
        from tools.serve_replay import serve_main
        serve_main(args, rank, world, local_rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
