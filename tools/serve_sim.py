"""Replay harness for BASELINE configs[2]/[3] on one GPU: Poisson arrivals of mixed-resolution
requests into sduss' `fcfs_mixed`-style continuous batching (every waiting request joins the
running batch at the next step boundary, up to max_batchsize; sduss/worker/scheduler/policy),
each request running `steps` denoising steps through this repo's drop-in `denoising_step`.
Prepare (text encoders) is not simulated; with SERVE_VAE=1 every finished request goes through the
post stage on the B200 VAE decoder (row f-4) before it counts as served.
Reports requests/s, latency percentiles, the number of distinct batch compositions and what
building their plans cost.
Usage: python tools/serve_sim.py sd3|sdxl [qps] [n_requests] [max_batchsize]"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

which = sys.argv[1] if len(sys.argv) > 1 else "sd3"
qps = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
n_req = int(sys.argv[3]) if len(sys.argv) > 3 else 120
max_bs = int(sys.argv[4]) if len(sys.argv) > 4 else 12
dev = torch.device("cuda")
if which == "sd3":
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline as P
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler as S
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel as M, SD3Config
    from sduss_b200.synthetic import make_sd3_requests as make, random_sd3_state_dict
    cfg = SD3Config(); model = M(random_sd3_state_dict(cfg, dev), cfg, device=dev); sch = S(); steps = 28
    pipe = P(model, sch)
    step = lambda reqs: pipe.denoising_step(reqs, True, 7.0, True, 256)
else:
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline as P
    from sduss_b200.schedulers import B200EulerDiscreteScheduler as S
    from sduss_b200.unet import B200UNet as M, UNetConfig
    from sduss_b200.synthetic import make_sdxl_requests as make, random_unet_state_dict
    cfg = UNetConfig(); cfg.context_len = 77
    model = M(random_unet_state_dict(cfg, dev), cfg, device=dev); sch = S(); steps = 50
    pipe = P(model, sch)
    step = lambda reqs: pipe.denoising_step(reqs, True, 0.0, 5.0, None, {}, None, None, None, True, 256)

with_vae = os.environ.get("SERVE_VAE", "0") == "1"
if with_vae:
    from sduss_b200.synthetic import random_vae_state_dict
    from sduss_b200.vae import B200VAEDecoder, VAEDecoderConfig
    vcfg = VAEDecoderConfig() if which != "sd3" else VAEDecoderConfig(
        latent_channels=16, scaling_factor=1.5305, shift_factor=0.0609, use_post_quant_conv=False)
    pipe.vae = B200VAEDecoder(random_vae_state_dict(vcfg, dev), vcfg, device=dev)
rng = random.Random(0)
arrivals, t = [], 0.0
for i in range(n_req):
    t += rng.expovariate(qps)
    arrivals.append((t, rng.choice(["512", "768", "1024"])))
pending = list(arrivals)
# requests are created up front (the prepare stage is not part of this path)
objs = [make(cfg, {res: 1}, steps, sch, dev, seed=100 + i)[res][0] for i, (_, res) in enumerate(arrivals)]
for i, o in enumerate(objs):
    o.request_id = i
running, done_lat, plan_time, n_steps = [], [], 0.0, 0
torch.cuda.synchronize()
t0 = time.perf_counter()
next_i = 0
while next_i < n_req or running:
    now = time.perf_counter() - t0
    while next_i < n_req and arrivals[next_i][0] <= now and len(running) < max_bs:
        running.append((next_i, arrivals[next_i][0])); next_i += 1
    if not running:
        time.sleep(max(0.0, arrivals[next_i][0] - now)); continue
    batch = {}
    for i, _ in running:
        batch.setdefault(arrivals[i][1], []).append(objs[i])
    n_plans = len(model._plans)
    ts = time.perf_counter()
    step(batch)
    if len(model._plans) != n_plans or model._plans.evictions:   # a new composition: plan build + capture
        torch.cuda.synchronize(); plan_time += time.perf_counter() - ts if len(model._plans) != n_plans else 0.0
    n_steps += 1
    keep, finished = [], {}
    for i, ta in running:
        if objs[i].scheduler_states._step_index >= steps:
            finished.setdefault(arrivals[i][1], []).append(objs[i])
        else:
            keep.append((i, ta))
    if finished:
        if with_vae:
            pipe.post_inference(finished, "pt")   # all resolutions that finished at this step, one pass
        torch.cuda.synchronize()
        for i, ta in running:
            if objs[i].scheduler_states._step_index >= steps:
                done_lat.append(time.perf_counter() - t0 - ta)
    running = keep
torch.cuda.synchronize()
wall = time.perf_counter() - t0
lat = np.asarray(done_lat)
print(f"{which}: {n_req} requests at {qps} req/s offered, {steps} steps each, max batch {max_bs}"
      + (", VAE decode of finished requests included" if with_vae else ""))
print(f"  wall {wall:.1f} s -> {n_req / wall:.2f} req/s served, {n_steps} batch steps ({n_steps / wall:.1f} steps/s, "
      f"{n_req * steps / wall:.0f} request-steps/s)")
print(f"  latency mean {lat.mean():.2f} s  p50 {np.percentile(lat, 50):.2f}  p99 {np.percentile(lat, 99):.2f}")
print(f"  distinct compositions (plans) {len(model._plans)}, evictions {model._plans.evictions}, "
      f"steps that built a plan (eager run + graph capture) took {plan_time:.1f} s "
      f"({100 * plan_time / wall:.0f} % of wall), "
      f"plan memory {model._plans.total_bytes() / 2**30:.1f} GiB")
