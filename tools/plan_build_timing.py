"""Where the time of building a plan (new batch composition) goes: table construction, eager
warm-up run, CUDA-graph capture, first replay. python tools/plan_build_timing.py sd3|sdxl"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "sd3"
dev = torch.device("cuda")
if which == "sd3":
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline as P
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler as S
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel as M, SD3Config
    from sduss_b200.synthetic import make_sd3_requests as make, random_sd3_state_dict
    cfg = SD3Config(); model = M(random_sd3_state_dict(cfg, dev), cfg, device=dev); sch = S()
    pipe = P(model, sch)
    step = lambda reqs: pipe.denoising_step(reqs, True, 7.0, True, 256)
else:
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline as P
    from sduss_b200.schedulers import B200EulerDiscreteScheduler as S
    from sduss_b200.unet import B200UNet as M, UNetConfig
    from sduss_b200.synthetic import make_sdxl_requests as make, random_unet_state_dict
    cfg = UNetConfig(); cfg.context_len = 77
    model = M(random_unet_state_dict(cfg, dev), cfg, device=dev); sch = S()
    pipe = P(model, sch)
    step = lambda reqs: pipe.denoising_step(reqs, True, 0.0, 5.0, None, {}, None, None, None, True, 256)

def timed(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {label:38s} {(time.perf_counter() - t) * 1e3:8.1f} ms"); return r

orig_run_eager, orig_graph = ops._run_eager, torch.cuda.graph
ops._run_eager = lambda m, p: timed("eager warm-up run (model._run)", lambda: orig_run_eager(m, p))
for spec in ({"512": 1}, {"512": 2, "768": 1, "1024": 1}, {"512": 4, "768": 4, "1024": 4}, {"512": 4, "768": 4, "1024": 3}):
    reqs = make(cfg, spec, 50, sch, dev, seed=1)
    print(f"{which} composition {spec}:")
    timed("step 1 (plan build + eager + capture)", lambda: step(reqs))
    timed("step 2 (first replay)", lambda: step(reqs))
    timed("step 3 (replay)", lambda: step(reqs))
