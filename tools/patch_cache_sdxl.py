"""Patch cache, SDXL variant, on the full SDXL-base architecture (random-init weights):
what the mechanism costs and saves, and the re-fit of the reference's two cuML forests
(ESYMRED_DOWNSAMPLE_PATH / ESYMRED_UPSAMPLE_PATH, cache_manager.py:27-36) with scikit-learn: while every
patch is recomputed the decision kernels record the feature rows [block, timestep, input MSE (, skip-tensor
MSEs)]; a (block, step, patch) is labelled "recompute" when the block's last same-level output moved by
more than 1 % (relative MSE against the previous step); one RandomForestClassifier for the down / mid
blocks, one for the up blocks, stored flattened in sduss_b200/data/patch_cache_sdxl_{down,up}_b200.npz. A 20-step trajectory of the given resolutions (1 request each, CFG)
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


DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sduss_b200", "data")
INDEX = {"down_blocks.0": 0, "down_blocks.1": 1, "down_blocks.2": 2, "mid_block": 3, "up_blocks.0": 4,
         "up_blocks.1": 5, "up_blocks.2": 6}
rows_fit, prev_out, fitting = {"down": [], "up": []}, {}, [True]


def last_output(cb, key):
    """The block's last output on the level its decision was taken on (before its sampler)."""
    names = [n for n in cb.bufs if n.startswith(key + ".") and n.endswith(".out")]
    res = sorted(n for n in names if ".resnets." in n)
    att = sorted(n for n in names if ".attentions." in n)
    return cb.bufs[(att or res)[-1]] if key != "mid_block" else cb.bufs["mid_block.resnets.1.out"]


def rec(k, reqs):
    pl = plan()
    cb = pl.cache
    ms = [m.float() for m in cb.masks().values()]
    share.append(float(torch.cat(ms).mean()))
    if k > 0:
        mses.append(torch.cat([st.mse for st in cb.blocks.values() if st is not None]).cpu().numpy())
    if not fitting[0]:
        return
    for key, st in cb.blocks.items():
        if st is None:
            continue
        cur = last_output(cb, key).float().view(st.n, -1)
        if key in prev_out and k > 0:
            rel = ((cur - prev_out[key]) ** 2).mean(1) / (prev_out[key] ** 2).mean(1).clamp_min(1e-12)
            t = pl.t32.cpu().numpy()[st.patch_latent.cpu().numpy()]
            feats = [st.mse.cpu().numpy()]
            up = key.startswith("up_blocks")
            if up:
                feats += list(cb.bufs[key + ".rmse"].cpu().numpy())
            for p in range(st.n):
                rows_fit["up" if up else "down"].append([INDEX[key], float(t[p])] + [float(f[p]) for f in feats]
                                                        + [float(rel[p])])
        prev_out[key] = cur.clone()


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
fitting[0] = False
from sklearn.ensemble import RandomForestClassifier
forests = {}
for kind in ("down", "up"):
    data = np.asarray(rows_fit[kind], dtype=np.float64)
    X, y = data[:, :-1].astype(np.float32), (data[:, -1] > 0.01).astype(int)
    rf = RandomForestClassifier(n_estimators=16, max_depth=8, random_state=0).fit(X, y)
    print(f"{kind} forest: {len(data)} rows x {X.shape[1]} features, {100 * y.mean():.1f} % labelled recompute, "
          f"training accuracy {rf.score(X, y):.3f}")
    f = forests[kind] = ops.DeviceForest.from_sklearn(rf, dev)
    path = os.path.join(DATA, f"patch_cache_sdxl_{kind}_b200.npz")
    np.savez_compressed(path, feature=f.t[0].cpu().numpy(), threshold=f.t[1].cpu().numpy(), left=f.t[2].cpu().numpy(),
                        right=f.t[3].cpu().numpy(), value=f.t[4].cpu().numpy(), roots=f.t[5].cpu().numpy(),
                        meta=np.frombuffer(("SDXL-base random-init, " + "+".join(res_list) + f", {STEPS} steps, label: block output "
                                            "moved > 1 % relative MSE; features [block, timestep, input MSE"
                                            + (", MSE of the 3 skip tensors]" if kind == "up" else "]")).encode(), dtype=np.uint8))
    import shutil
    shutil.copy(path, os.path.join(DATA, "..", "..", "gpurun_out", os.path.basename(path)))
model.enable_patch_cache(forests["down"], forests["up"], refresh=4)
share.clear()
p_f, ms_f = trajectory(rec)
c_f = [torch.nn.functional.cosine_similarity(a, b, dim=0).item() for a, b in zip(p_f, off_preds)]
print(f"cache on, fitted forests: {np.mean(ms_f[4:]):.2f} ms/step, patches recomputed {100 * np.mean(share[4:]):.1f} %, "
      f"min cos vs exact {min(c_f):.4f}")
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
