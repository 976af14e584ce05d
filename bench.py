#!/usr/bin/env python
"""bench.py -- mixed-resolution denoise steps/s on B200 (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1]: SD3.5-medium MMDiT denoising step over a mixed batch
of three requests (512^2 + 768^2 + 1024^2), CFG on (6 latents, 14 848 image tokens + 6x333
context tokens), bf16, random-init weights, synthetic latents / embeddings. One "step" = one
`denoising_step` call over the whole batch: gather, model forward, CFG combine, flow-match
update, write-back (reference: pipeline_stable_diffusion_3_esymred.py:232-388).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--model sd3|sdxl]

`value`  = steps/s with the request state resident in HBM (as in sduss, where request latents
           and embeddings live on the GPU between steps), all N GPUs summed (replicas, weak).
`e2e`    = the same call with every request tensor in pinned HOST memory: per step the latents,
           prompt embeddings and pooled embeddings go host->device and the updated latents
           come back device->host, all inside the timed region.
`roofline` = dominant kernel (packed varlen attention) measured live with CUDA events around
           each of its launches inside the timed steps.
`cpu_baseline` / `--impl reference` = the oracle (PyTorch fp32 restatement of the reference
           path; the reference itself cannot run on CPU nor without diffusers) on the host
           cores, on a bounded sample (one 512^2 request with CFG through the full model),
           scaled to whole steps by algorithmic FLOPs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPEC = {"512": 1, "768": 1, "1024": 1}
SD3_STEPS = 28  # BASELINE.json configs[3]: 28-step flow matching
GUIDANCE = 7.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
        "fallback (B200_PROFILING.md)"


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


def sd3_step_flops():
    from sduss_b200.sd3_transformer import SD3Config
    sys.path.insert(0, ROOT)
    from oracle.sd3_mmdit import sd3_flops_per_latent, sd35_medium_config
    cfg = sd35_medium_config()
    per = {r: sd3_flops_per_latent(cfg, int(r)) for r in SPEC}
    return sum(2 * n * per[r] for r, n in SPEC.items()), per


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


# ----------------------------------------------------------------------------- CPU oracle leg
def cpu_oracle_steps_per_s(state_dict_fp32=None, repeats=1):
    """Times the oracle on the host cores on a bounded sample: the 512^2 request with CFG
    (2 latents, 1024+333 tokens each) through the full 24-layer SD3.5-medium MMDiT + CFG +
    flow-match update; whole-step throughput = sample rate * FLOPs(sample) / FLOPs(step)."""
    from oracle import schedulers as osch
    from oracle import sd3_mmdit as o3
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = o3.sd35_medium_config()
    sd = state_dict_fp32 if state_dict_fp32 is not None else o3.init_sd3_weights(cfg, 0)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 16, 64, 64, generator=g)
    ehs = torch.randn(2, 333, 4096, generator=g)
    pooled = torch.randn(2, 2048, generator=g)
    sig, ts = osch.flow_match_sigmas(SD3_STEPS)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = o3.sd3_forward(sd, cfg, {"512": torch.cat([lat, lat])}, ehs, pooled, ts[:1].repeat(2))
        eps = osch.cfg_combine(out["512"], GUIDANCE)
        osch.flow_match_batch_step(eps, lat, sig[:1], sig[1:2])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    step_fl, per = sd3_step_flops()
    sample_fl = 2 * per["512"]
    steps_per_s = (1.0 / best) * (sample_fl / step_fl)
    return {"value": steps_per_s, "unit": "denoise steps/s", "cores": cores, "kind": "port",
            "sample": f"1x512^2 request with CFG (2 latents) through the full SD3.5-medium oracle, "
                      f"fp32, {best:.2f} s; scaled by FLOPs {sample_fl / 1e12:.2f}/{step_fl / 1e12:.2f} T",
            "sample_seconds": best}


def run_reference(args, rank, world):
    if rank != 0:
        return
    res = cpu_oracle_steps_per_s(repeats=max(1, min(args.steps, 2)))
    line = {"impl": "reference", "metric": "mixed-res denoise steps/s (SD3.5-medium, 512^2+768^2+1024^2, CFG)",
            "value": res["value"], "unit": "denoise steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / res["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: SD3.5-medium MMDiT denoise step, mixed batch "
                                   "512^2+768^2+1024^2 (1 request each), CFG on -> 6 latents, random-init",
                       "arm": "oracle (fp32 restatement of the reference path) on the host cores; each "
                              "step is a bounded sample scaled by FLOPs"},
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
            for t in (sp.latents, sp.prompt_embeds, sp.negative_prompt_embeds,
                      po.pooled_prompt_embeds, po.negative_pooled_prompt_embeds):
                h2d += t.numel() * t.element_size()
            d2h += sp.latents.numel() * sp.latents.element_size()
    return h2d, d2h


def run_b200(args, rank, world, local_rank):
    from sduss_b200 import ops
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel, SD3Config
    from sduss_b200.synthetic import make_sd3_requests, random_sd3_state_dict

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    cfg = SD3Config()
    sd = random_sd3_state_dict(cfg, dev, seed=0)
    model = B200SD3Transformer2DModel(sd, cfg, device=dev)
    sched = B200FlowMatchEulerDiscreteScheduler()
    pipe = B200StableDiffusion3Pipeline(model, sched)

    def fresh(pin_host):
        return make_sd3_requests(cfg, SPEC, 4 * (args.steps + args.warmup) + 32, sched, dev,
                                 seed=rank, pin_host=pin_host)

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

    # ---- device-resident steps (value) with per-launch events on the dominant kernel
    reqs = fresh(False)
    step = lambda: pipe.denoising_step(reqs, True, GUIDANCE, True, 256)
    clocks = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step()
    barrier()
    clocks.start()
    # per-launch CUDA events (eager launches, one stream, so that no other kernel of the step
    # runs beside the one being timed); the step itself is re-timed below as the product runs it
    ops.profile = {}
    model.two_streams = False
    n0 = ops.launch_count
    ms_step = timed(step, args.steps, 0)
    launches = ops.launch_count - n0
    model.two_streams = True
    prof, ops.profile = ops.profile, None
    prof.pop("tags", None)
    torch.cuda.synchronize()
    attn_ms = sum(a.elapsed_time(b) for a, b in prof.get("b200_attn_varlen_bf16", []))
    gemm_ms = sum(a.elapsed_time(b) for a, b in prof.get("b200_gemm_bf16", []))
    n_attn = len(prof.get("b200_attn_varlen_bf16", []))
    # the timed region proper: CUDA-graph replay of the two-stream schedule, no per-launch events
    ms_step = min(ms_step, timed(step, args.steps, 1))
    clk = clocks.stop()

    # ---- e2e: every request tensor in pinned host memory, H2D + D2H inside the timed region
    hreqs = fresh(True)
    h2d, d2h = tensor_bytes(hreqs)

    def e2e_step():
        dreqs = {}
        for res, rs in hreqs.items():
            dreqs[res] = []
            for r in rs:
                sp, po = r.sampling_params, r.prepare_output
                d = type(r)(request_id=r.request_id, scheduler_states=r.scheduler_states,
                            sampling_params=type(sp)(**{k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v)
                                                        for k, v in vars(sp).items()}),
                            prepare_output=type(po)(**{k: v.to(dev, non_blocking=True) for k, v in vars(po).items()}))
                dreqs[res].append(d)
        pipe.denoising_step(dreqs, True, GUIDANCE, True, 256)
        for res, rs in hreqs.items():
            for r, d in zip(rs, dreqs[res]):
                r.sampling_params.latents.copy_(d.sampling_params.latents, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the host owns the result before the next step

    ms_e2e = timed(e2e_step, args.steps, args.warmup)

    if rank != 0:
        return
    pk, pk_src = peaks()
    step_fl, _ = sd3_step_flops()
    attn_fl = attn_flops_per_step(cfg, SPEC)
    achieved = attn_fl * args.steps / (attn_ms / 1e3) / 1e12 if attn_ms > 0 else None
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    line = {
        "metric": "mixed-res denoise steps/s (SD3.5-medium, 512^2+768^2+1024^2, CFG)",
        "value": world * 1000.0 / ms_step, "unit": "denoise steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: SD3.5-medium MMDiT denoise step, mixed batch "
                               "512^2+768^2+1024^2 (1 request each), CFG on -> 6 latents, bf16, random-init",
                   "requests_per_step": sum(SPEC.values()), "guidance": GUIDANCE,
                   "l2": "weights (4.9 GB bf16) and activations streamed every step exceed the 126 MB L2; no flush",
                   "parallelism": f"dp{world} replicas, no collective",
                   "step_tflop": step_fl / 1e12,
                   "model_tflops_per_gpu": step_fl / 1e12 / (ms_step / 1e3)},
        "req_steps_per_s": world * sum(SPEC.values()) * 1000.0 / ms_step,
        "e2e": {"value": world * 1000.0 / ms_e2e, "unit": "denoise steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"kernel": "attn_fwd_kernel (b200_attn_varlen_bf16)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": (achieved / peak) if achieved else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one joint-attention launch of
                     # this workload (ncu --set full of the final round-1 kernel,
                     # profiles/r01_ncu_attn_final.txt); algorithmic bytes of that launch (Q, K, V in,
                     # O out): 207 MB
                     "traffic": 185323776 + 42830848, "traffic_unit": "bytes per joint-attention launch",
                     "peak_source": pk_src + ", sustained bf16",
                     "launches_timed": n_attn, "kernel_ms_per_step": attn_ms / args.steps,
                     "gemm_ms_per_step": gemm_ms / args.steps,
                     "algorithmic_tflop_per_step": attn_fl / 1e12},
    }
    if world == 1 and not args.no_cpu:
        sd32 = {k: v.float().cpu() for k, v in sd.items()}
        line["cpu_baseline"] = cpu_oracle_steps_per_s(sd32)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- SDXL leg
SDXL_SPEC = {"512": 1, "1024": 1}  # BASELINE configs[0] shape: 512^2 + 1024^2, CFG -> 4 latents


def sdxl_cpu_oracle_steps_per_s(sd32):
    """Oracle SDXL UNet, fp32, host cores: the 512^2 request with CFG (2 latents) + scale input +
    CFG + Euler update; scaled to the whole 512^2+1024^2 step by FLOPs."""
    from oracle import schedulers as osch
    from oracle import sdxl_unet as ox
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ox.sdxl_base_config()
    g = torch.Generator().manual_seed(0)
    sig, ts, init = osch.euler_sigmas(50)
    lat = torch.randn(1, 4, 64, 64, generator=g) * init
    ehs, te = torch.randn(2, 77, 2048, generator=g), torch.randn(2, 1280, generator=g)
    ids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * 2)
    t0 = time.perf_counter()
    xin = osch.batch_scale_model_input(torch.cat([lat, lat]), [sig[0]])
    out = ox.unet_forward(sd32, cfg, {"512": xin}, ts[:1].repeat(2), ehs, te, ids)
    osch.euler_batch_step(osch.cfg_combine(out["512"], 5.0), lat, [sig[0]], [sig[1]])
    dt = time.perf_counter() - t0
    per = {r: ox.unet_flops_per_latent(cfg, int(r)) for r in SDXL_SPEC}
    step_fl = sum(2 * n * per[r] for r, n in SDXL_SPEC.items())
    return {"value": (1.0 / dt) * (2 * per["512"] / step_fl), "unit": "denoise steps/s", "cores": cores,
            "kind": "port", "sample_seconds": dt,
            "sample": f"1x512^2 request with CFG (2 latents) through the full SDXL-base oracle, fp32, {dt:.2f} s; "
                      f"scaled by FLOPs {2 * per['512'] / 1e12:.2f}/{step_fl / 1e12:.2f} T"}, step_fl


def run_b200_sdxl(args, rank, world, local_rank):
    from sduss_b200 import ops
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline
    from sduss_b200.schedulers import B200EulerDiscreteScheduler
    from sduss_b200.synthetic import make_sdxl_requests, random_unet_state_dict
    from sduss_b200.unet import B200UNet, UNetConfig
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = UNetConfig()
    cfg.context_len = 77
    sd = random_unet_state_dict(cfg, dev, seed=0)
    model = B200UNet(sd, cfg, device=dev)
    sched = B200EulerDiscreteScheduler()
    pipe = B200StableDiffusionXLPipeline(model, sched)
    reqs = make_sdxl_requests(cfg, SDXL_SPEC, 4 * (args.steps + args.warmup) + 32, sched, dev, seed=rank)
    step = lambda: pipe.denoising_step(reqs, True, 0.0, 5.0, None, {}, None, None, None, True, 256)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            step()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ops.profile = {}
    n0 = ops.launch_count
    ms_prof = timed(args.steps)
    launches = ops.launch_count - n0
    prof, ops.profile = ops.profile, None
    prof.pop("tags", None)
    clk = clocks.stop()
    torch.cuda.synchronize()
    per_kernel = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in prof.items()}
    step()  # first un-profiled call captures the CUDA graph of the forward
    ms_step = min(ms_prof, timed(args.steps))
    if rank != 0:
        return
    pk, pk_src = peaks()
    line = {"metric": "mixed-res denoise steps/s (SDXL-base, 512^2+1024^2, CFG)",
            "value": world * 1000.0 / ms_step, "unit": "denoise steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[0] shape on B200: SDXL-base UNet denoise step, 512^2 + 1024^2 "
                                   "(1 request each), CFG on -> 4 latents, bf16, random-init",
                       "parallelism": f"dp{world} replicas, no collective"},
            "req_steps_per_s": world * sum(SDXL_SPEC.values()) * 1000.0 / ms_step,
            "gpu_launches": launches, "clocks": clk,
            "kernel_ms_per_step": {k: round(v, 3) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])}}
    if world == 1 and not args.no_cpu:
        sd32 = {k: v.float().cpu() for k, v in sd.items()}
        line["cpu_baseline"], step_fl = sdxl_cpu_oracle_steps_per_s(sd32)
        line["config"]["step_tflop"] = step_fl / 1e12
        line["config"]["model_tflops_per_gpu"] = step_fl / 1e12 / (ms_step / 1e3)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--model", default="sd3", choices=["sd3", "sdxl"],
                    help="sd3 = BASELINE configs[1] (default, the headline); sdxl = configs[0] shape on B200")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    if args.model == "sdxl":
        run_b200_sdxl(args, rank, world, local_rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
