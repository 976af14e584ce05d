"""Patch cache, SDXL variant, on the full SDXL-base architecture (random-init weights):
what the mechanism costs and saves. A 20-step trajectory of the given resolutions (1 request each, CFG)
with the cache off, with the cache on and everything flagged (the pure overhead: 5 decisions per step and
the persistent buffers), and with the rule "recompute iff input MSE > tau" at quantiles of the MSE the
blocks saw -- ms per step, share of the patches recomputed, and how far the applied update drifts from
the exact trajectory (cosine per step, worst).
python tools/patch_cache_sdxl.py [resolution ...]      (default 1024)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from sduss_b200 import ops

res_list = sys.argv[1:] or ["1024"]
dev = torch.device("cuda")
cfg, sd, pipe, make, call = bench.build_pipeline("sdxl", dev)
del sd
model = pipe.model
STEPS = 20
spec = {r: 1 for r in res_list}
share, mses = [], []


def plan():
    return [p for p in model._plans.values() if p.cache is not None][0]


def rec(k, reqs):
    cb = plan().cache
    ms = [m.float() for m in cb.masks().values()]
    share.append(float(torch.cat(ms).mean()))
    if k > 0:
        mses.append(torch.cat([st.mse for st in cb.blocks.values() if st is not None]).cpu().numpy())


def trajectory(record=None):
    reqs = make(spec, 50, 3)
    preds, ms = [], []
    for k in range(STEPS):
        x0 = {r: rs[0].sampling_params.latents.float().clone() for r, rs in reqs.items()}
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        call(reqs)
        e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
        preds.append(torch.cat([(rs[0].sampling_params.latents.float() - x0[r]).flatten() for r, rs in reqs.items()]))
        if record is not None:
            record(k, reqs)
    return preds, ms


off_preds, ms_off = trajectory()
off = float(np.mean(ms_off[4:]))
print(f"## SDXL-base, {'+'.join(res_list)} (1 request each, CFG), {STEPS} steps, one B200")
print(f"cache off: {off:.2f} ms/step (mean of steps 4..{STEPS - 1})")
model.enable_patch_cache(ops.DeviceForest.threshold_rule(-1.0, dev), refresh=4)
p_all, ms_all = trajectory(rec)
print(f"cache on, every patch flagged: {np.mean(ms_all[4:]):.2f} ms/step ({np.mean(ms_all[4:]) / off:.3f} of off), "
      f"bit-identical to off: {all(torch.equal(a, b) for a, b in zip(p_all, off_preds))}; "
      f"cache buffers {plan().cache.bytes() / 2**30:.2f} GiB")
seen = np.concatenate(mses)
print(f"input MSE seen by the deciding blocks: median {np.median(seen):.3e}, 10 % {np.quantile(seen, 0.1):.3e}, 90 % {np.quantile(seen, 0.9):.3e}")
print(f"{'tau quantile':>12s} {'recomputed':>11s} {'ms/step':>9s} {'vs off':>8s} {'min cos':>8s}")
for q in (0.25, 0.5, 0.75, 0.9, 1.0):
    tau = float(np.quantile(seen, q)) if q < 1.0 else 3e38
    model.enable_patch_cache(ops.DeviceForest.threshold_rule(tau, dev), refresh=4)
    share.clear()
    p_q, ms_q = trajectory(rec)
    c_q = [torch.nn.functional.cosine_similarity(a, b, dim=0).item() for a, b in zip(p_q, off_preds)]
    # (mean, not median: with the forced refresh every fifth step the per-step times are bimodal)
    print(f"{q:12.2f} {100 * np.mean(share[4:]):10.1f}% {np.mean(ms_q[4:]):9.2f} {np.mean(ms_q[4:]) / off:8.3f} {min(c_q):8.4f}"
          f"   per step: " + " ".join(f"{m:.1f}" for m in ms_q[4:]))
