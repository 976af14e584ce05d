"""Per-kernel / per-shape breakdown of one denoising step (eager mode, CUDA events per launch)."""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops

ap = argparse.ArgumentParser(); ap.add_argument("--model", default="sd3"); ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda")
if a.model == "sd3":
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline as P
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler as S
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel as M, SD3Config
    from sduss_b200.synthetic import make_sd3_requests, random_sd3_state_dict
    cfg = SD3Config(); model = M(random_sd3_state_dict(cfg, dev), cfg, device=dev); sch = S()
    pipe = P(model, sch); reqs = make_sd3_requests(cfg, {"512": 1, "768": 1, "1024": 1}, 200, sch, dev)
    step = lambda: pipe.denoising_step(reqs, True, 7.0, True, 256)
else:
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline as P
    from sduss_b200.schedulers import B200EulerDiscreteScheduler as S
    from sduss_b200.unet import B200UNet as M, UNetConfig
    from sduss_b200.synthetic import make_sdxl_requests, random_unet_state_dict
    cfg = UNetConfig(); cfg.context_len = 77
    model = M(random_unet_state_dict(cfg, dev), cfg, device=dev); sch = S()
    pipe = P(model, sch); reqs = make_sdxl_requests(cfg, {"512": 1, "1024": 1}, 200, sch, dev)
    step = lambda: pipe.denoising_step(reqs, True, 0.0, 5.0, None, {}, None, None, None, True, 256)
model.use_graphs = False
for _ in range(2): step()
torch.cuda.synchronize()
ops.profile = {}
for _ in range(a.steps): step()
torch.cuda.synchronize()
prof, ops.profile = ops.profile, None
tags = prof.pop("tags", [])
tot = {k: sum(s.elapsed_time(e) for s, e in v) / a.steps for k, v in prof.items()}
print("kernel totals (ms/step):")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]): print(f"  {k:32s} {v:8.3f}  launches/step {len(prof[k]) // a.steps}")
print(f"  {'SUM':32s} {sum(tot.values()):8.3f}")
by = collections.defaultdict(lambda: [0.0, 0])
for name, tag, (s, e) in tags:
    by[(name, tag)][0] += s.elapsed_time(e) / a.steps; by[(name, tag)][1] += 1
print("per shape (ms/step, calls/step, TFLOP/s):  tag = (M, N, K, epi|stride)")
for (name, tag), (ms, n) in sorted(by.items(), key=lambda kv: -kv[1][0])[:40]:
    fl = 2.0 * tag[0] * tag[1] * tag[2] * (n / a.steps)
    print(f"  {name[5:-5]:10s} {str(tag):32s} {ms:8.3f} ms  x{n // a.steps:3d}  {fl / ms / 1e9:7.0f} TF/s")
