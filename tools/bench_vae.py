"""VAE decode stage (row f-4) at full size: SDXL VAE (4-ch latents) or SD3 VAE (16-ch), a
mixed-resolution batch of finished requests decoded in one pass. Reports GPU time per decode
(CUDA-graph replay, CUDA events), achieved TFLOP/s against the oracle's FLOP count, images/s, and a
per-kernel / per-shape breakdown (eager pass with events around every launch)."""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops
from sduss_b200.synthetic import random_vae_state_dict, vae_decode_flops
from sduss_b200.vae import B200VAEDecoder, VAEDecoderConfig

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="sdxl", choices=["sdxl", "sd3"])
ap.add_argument("--spec", default="512:1,1024:1", help="resolution:count,...")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda")
cfg = VAEDecoderConfig() if a.model == "sdxl" else VAEDecoderConfig(
    latent_channels=16, scaling_factor=1.5305, shift_factor=0.0609, use_post_quant_conv=False)
sd = random_vae_state_dict(cfg, dev)
model = B200VAEDecoder(sd, cfg, device=dev)
spec = {r: int(n) for r, n in (kv.split(":") for kv in a.spec.split(","))}
g = torch.Generator().manual_seed(0)
lat = {r: (torch.randn(n, cfg.latent_channels, int(r) // 8, int(r) // 8, generator=g) * 0.8).bfloat16().to(dev)
       for r, n in sorted(spec.items(), key=lambda kv: int(kv[0]))}
flops = sum(n * vae_decode_flops(cfg, int(r) // 8, int(r) // 8) for r, n in spec.items())
for _ in range(3):
    out = model.decode(lat, _borrow=True)
torch.cuda.synchronize()
assert all(torch.isfinite(v.float()).all() for v in out.values())
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.iters):
    model.decode(lat, _borrow=True)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / a.iters
n_img = sum(spec.values())
print(f"{a.model} VAE decode {a.spec}: {ms:.2f} ms per batch ({n_img} images, {n_img / ms * 1e3:.1f} images/s), "
      f"{flops / 1e12:.2f} TFLOP -> {flops / ms / 1e9:.0f} TFLOP/s; plan memory "
      f"{ops.PlanCache.plan_bytes(next(iter(model._plans.values()))) / 2**30:.2f} GiB")
model.use_graphs = False
ops.profile = {}
model.decode(lat, _borrow=True)
torch.cuda.synchronize()
prof, ops.profile = ops.profile, None
tags = prof.pop("tags", [])
tot = {k: sum(s.elapsed_time(e) for s, e in v) for k, v in prof.items()}
print("kernel totals (ms per decode, eager):")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]): print(f"  {k:32s} {v:8.3f}  launches {len(prof[k])}")
by = collections.defaultdict(lambda: [0.0, 0])
for name, tag, (s, e) in tags:
    by[(name, tag)][0] += s.elapsed_time(e); by[(name, tag)][1] += 1
print("per shape (ms, calls, TFLOP/s):  tag = (M, N, K, epi|stride)")
for (name, tag), (t, n) in sorted(by.items(), key=lambda kv: -kv[1][0])[:24]:
    print(f"  {name[5:-5]:10s} {str(tag):32s} {t:8.3f} ms  x{n:3d}  {2.0 * tag[0] * tag[1] * tag[2] * n / t / 1e9:7.0f} TF/s")
