"""Row f-1: regenerate the reference's exp/profile/unet_time_<model>.csv on
B200 with this repo's denoising step. Same columns ("512 num, 768 num, 1024 num, avg unet time"),
same meaning: seconds for 50 denoising steps of a batch holding that many requests per
resolution (CFG on), which is what sduss' schedule predictor is trained on
(sduss/worker/scheduler/policy/ESyMReD.py:45-51, exp/profile/unet_time_*.csv).
Usage (GPU): python tools/profile_compositions.py sd3|sdxl [n_compositions] [out.csv]"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

which = sys.argv[1]
n_comp = int(sys.argv[2]) if len(sys.argv) > 2 else 70
out = sys.argv[3] if len(sys.argv) > 3 else f"gpurun_out/unet_time_{which}_b200.csv"
dev = torch.device("cuda")
STEPS = 50


def compositions(n, seed=0):
    """Singletons and pure batches first (they anchor STANDALONE and the per-resolution slopes),
    then random mixes from the reference's domain: <= 12 / 8 / 5 requests per resolution and
    <= 15 requests per batch (exp/profile/unet_time_*.csv)."""
    comps = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1), (2, 0, 0), (0, 2, 0), (0, 0, 2), (4, 0, 0), (0, 4, 0),
             (0, 0, 4), (8, 0, 0), (12, 0, 0), (0, 8, 0), (0, 0, 5), (6, 4, 2), (10, 3, 2)]
    rng = random.Random(seed)
    seen = set(comps)
    while len(comps) < n:
        c = (rng.randint(0, 12), rng.randint(0, 8), rng.randint(0, 5))
        if 0 < sum(c) <= 15 and c not in seen:
            seen.add(c)
            comps.append(c)
    return comps[:n]


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

rows = []
for a, b, c in compositions(n_comp):
    spec = {k: v for k, v in (("512", a), ("768", b), ("1024", c)) if v}
    reqs = make(cfg, spec, 60, sch, dev, seed=1)
    for _ in range(3):   # eager + graph capture + one replay
        step(reqs)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(4):
        step(reqs)
    e.record(); torch.cuda.synchronize()
    sec50 = s.elapsed_time(e) / 4 / 1e3 * STEPS
    rows.append((a, b, c, sec50))
    print(f"{a},{b},{c},{sec50:.6f}", flush=True)
    model._plans.clear(); del reqs
    torch.cuda.empty_cache()
with open(out, "w") as f:
    f.write("512 num, 768 num, 1024 num, avg unet time\n")
    for r in rows:
        f.write(f"{r[0]},{r[1]},{r[2]},{r[3]}\n")
