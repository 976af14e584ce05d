"""Runs a few denoising steps (eager, no CUDA graph) -- the command profiled by ncu."""
import argparse, os, sys
os.environ.setdefault("SDUSS_B200_NO_GRAPH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
for _ in range(a.steps): step()
torch.cuda.synchronize()
print("done")
