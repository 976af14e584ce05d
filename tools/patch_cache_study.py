"""Patch cache on the full SD3.5-medium architecture (random-init weights):

  1. FIT   run a 28-step trajectory with every patch recomputed while the decision kernel records the
           per-block, per-patch input MSE; label a (block, step, patch) "recompute" when the block's
           OUTPUT for that patch moved by more than 1 % (relative MSE against the previous step,
           the reference's `...-threshold0.01` predictors); fit an sklearn RandomForestClassifier on
           [block, timestep, input MSE] -- the re-fit of the reference's cuML forests
           (exp/sd3-state-threshold0.01.pkl cannot be unpickled without cuml) -- and store it
           flattened in sduss_b200/data/patch_cache_sd3_b200.npz.
  2. MEASURE  the same trajectory with the cache off / on (fitted forest): ms per step, share of
           patches recomputed, and how far the cached trajectory's predictions drift from the exact
           ones (cosine per step).
python tools/patch_cache_study.py [resolution ...]      (default 1024)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from sduss_b200 import ops

res_list = sys.argv[1:] or ["1024"]
dev = torch.device("cuda")
cfg, sd, pipe, make, call = bench.build_pipeline("sd3", dev)
del sd
model = pipe.model
STEPS = 28
spec = {r: 1 for r in res_list}
DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sduss_b200", "data",
                    "patch_cache_sd3_b200.npz")


def trajectory(record=None):
    reqs = make(spec, STEPS, 3)
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


# ---- 1. FIT: everything recomputed (rule: MSE > -1), features from the decision kernel, labels from xout
model.enable_patch_cache(ops.DeviceForest.threshold_rule(-1.0, dev), refresh=2)
rows, prev_out = [], {}


def record(k, reqs):
    pl = next(iter(model._plans.values()))
    cb = pl.cache
    mse = cb.mse.cpu().numpy()                         # [blocks, patches] input MSE vs previous step
    t = pl.t32.cpu().numpy()[cb.patch_latent.cpu().numpy()]
    for i, xo in enumerate(cb.xout):
        cur = xo.float().view(cb.n_patches, -1)
        if i in prev_out and k > 0:
            rel = ((cur - prev_out[i]) ** 2).mean(1) / (prev_out[i] ** 2).mean(1).clamp_min(1e-12)
            for p in range(cb.n_patches):
                rows.append((i, float(t[p]), float(mse[i, p]), float(rel[p])))
        prev_out[i] = cur.clone()


exact_preds, ms_all = trajectory(record)
data = np.asarray(rows, dtype=np.float64)
y = (data[:, 3] > 0.01).astype(int)
print(f"dataset: {len(data)} rows (block, timestep, input MSE -> output moved > 1 %): {y.mean() * 100:.1f} % recompute")
from sklearn.ensemble import RandomForestClassifier
rf = RandomForestClassifier(n_estimators=16, max_depth=8, random_state=0).fit(data[:, :3].astype(np.float32), y)
print(f"forest: 16 trees, depth <= 8, training accuracy {rf.score(data[:, :3].astype(np.float32), y):.3f}")
f = ops.DeviceForest.from_sklearn(rf, dev)
np.savez_compressed(DATA, feature=f.t[0].cpu().numpy(), threshold=f.t[1].cpu().numpy(), left=f.t[2].cpu().numpy(),
                    right=f.t[3].cpu().numpy(), value=f.t[4].cpu().numpy(), roots=f.t[5].cpu().numpy(),
                    meta=np.frombuffer(("SD3.5-medium random-init, " + "+".join(res_list) + ", 28 steps, label: block output "
                                        "moved > 1 % relative MSE; features [block, timestep, input MSE]").encode(), dtype=np.uint8))
print("wrote", DATA)
import shutil
os.makedirs(os.path.join(os.path.dirname(DATA), "..", "..", "gpurun_out"), exist_ok=True)
shutil.copy(DATA, os.path.join(os.path.dirname(DATA), "..", "..", "gpurun_out", os.path.basename(DATA)))

# ---- 2. MEASURE
model.enable_patch_cache(None)
off_preds, ms_off = trajectory()
model.enable_patch_cache(ops.DeviceForest.from_npz(DATA, dev), refresh=2)
share = []


def rec2(k, reqs):
    pl = next(iter(model._plans.values()))
    share.append(float(pl.cache.mask.float().mean()))


on_preds, ms_on = trajectory(rec2)
cos = [torch.nn.functional.cosine_similarity(a, b, dim=0).item() for a, b in zip(on_preds, off_preds)]
print(f"## SD3.5-medium, {'+'.join(res_list)} (1 request each, CFG), 28 steps, one B200")
print(f"cache off: {np.median(ms_off[3:]):.2f} ms/step (median of steps 3..27)")
print(f"cache on : {np.median(ms_on[3:]):.2f} ms/step, patches recomputed {100 * np.mean(share[1:]):.1f} % on average "
      f"(step 0: {100 * share[0]:.0f} %)")
print("applied-update cosine, cached vs exact trajectory, per step:")
print("  " + " ".join(f"{c:.4f}" for c in cos))
print("recomputed share per step:")
print("  " + " ".join(f"{s:.2f}" for s in share))

# ---- 3. what the mechanism itself saves: the rule "recompute iff MSE > tau" at quantiles of the observed MSE
print("## mechanism at imposed skip rates (rule: recompute iff input MSE > tau; tau = quantile of the MSE seen in 1.)")
print(f"{'tau quantile':>12s} {'recomputed':>11s} {'ms/step':>9s} {'vs off':>8s} {'min cos':>8s}")
off = np.median(ms_off[3:])
for q in (0.0, 0.25, 0.5, 0.75, 0.9):
    tau = float(np.quantile(data[:, 2], q)) if q > 0 else -1.0
    model.enable_patch_cache(ops.DeviceForest.threshold_rule(tau, dev), refresh=2)
    share.clear()
    p_q, ms_q = trajectory(rec2)
    c_q = [torch.nn.functional.cosine_similarity(a, b, dim=0).item() for a, b in zip(p_q, off_preds)]
    print(f"{q:12.2f} {100 * np.mean(share[1:]):10.1f}% {np.median(ms_q[3:]):9.2f} {np.median(ms_q[3:]) / off:8.3f} {min(c_q):8.4f}")
