"""Stress test of batch invariance through the SD3 drop-in (tiny config): the same requests are
stepped alone and with a changing set of batch mates; every mismatch of the final latents is
counted. Usage: python tools/stress_invariance.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import sd3_mmdit as o3
from sduss_b200.pipelines import B200StableDiffusion3Pipeline
from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
from sduss_b200.sd3_transformer import B200SD3Transformer2DModel
from sduss_b200.synthetic import make_sd3_requests

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cuda = torch.device("cuda")
cfg = o3.sd3_tiny_config()
sd = {k: v.to(torch.bfloat16).float() for k, v in o3.init_sd3_weights(cfg, 0).items()}
model = B200SD3Transformer2DModel(sd, cfg, device="cuda")
sched = B200FlowMatchEulerDiscreteScheduler()
pipe = B200StableDiffusion3Pipeline(model, sched)
bad = 0
for it in range(n_iter):
    # churn the caching allocator so that fresh workspaces see different stale contents
    junk = [torch.full((1 << 20,), float("nan"), device=cuda, dtype=torch.bfloat16) for _ in range(4)]
    del junk
    a = make_sd3_requests(cfg, {"256": 1, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=5 + it)
    b = make_sd3_requests(cfg, {"256": 1, "512": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=5 + it)
    extra = make_sd3_requests(cfg, {"768": 1}, 28, sched, cuda, ctx_len=cfg.context_len, seed=600 + it)
    for _ in range(3):
        pipe.denoising_step(a, True, 7.0, True, 256)
    pipe.denoising_step(b, True, 7.0, True, 256)
    pipe.denoising_step({**b, **extra}, True, 7.0, True, 256)
    pipe.denoising_step(b, True, 7.0, True, 256)
    torch.cuda.synchronize()
    for res in a:
        x, y = a[res][0].sampling_params.latents, b[res][0].sampling_params.latents
        if not torch.equal(x, y):
            bad += 1
            print(f"iter {it} res {res}: {int((x != y).sum())} of {x.numel()} elements differ, max abs {float((x.float() - y.float()).abs().max()):.4g}")
print(f"mismatching (iteration, request) pairs: {bad} of {2 * n_iter}")
