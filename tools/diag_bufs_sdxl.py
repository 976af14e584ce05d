"""Diagnostic: SDXL forward twice on identical inputs, no syncs in between launches; every named
intermediate buffer of the plan is snapshotted after each call and compared (creation order =
dataflow order of first use)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200.unet import B200UNet, UNetConfig
from sduss_b200.synthetic import random_unet_state_dict
cuda = torch.device("cuda")
cfg = UNetConfig(); cfg.context_len = 77
spec = {}
for kv in (sys.argv[1] if len(sys.argv) > 1 else "512:2,1024:2").split(","):
    k, v = kv.split(":"); spec[k] = int(v)
model = B200UNet(random_unet_state_dict(cfg, cuda, seed=0), cfg, device=cuda)
g = torch.Generator().manual_seed(3)
hs = {r: torch.randn(n, 4, int(r) // 8, int(r) // 8, generator=g).to(cuda, torch.bfloat16) for r, n in spec.items()}
L = sum(spec.values())
ehs = torch.randn(L, 77, cfg.cross_attention_dim, generator=g).to(cuda, torch.bfloat16)
te = torch.randn(L, cfg.pooled_dim, generator=g).to(cuda, torch.bfloat16)
ids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * L).to(cuda, torch.bfloat16)
t = torch.full((L,), 999.0, device=cuda)
snaps = []
for i in range(3):
    model(hs, t, encoder_hidden_states=ehs, added_cond_kwargs={"text_embeds": te, "time_ids": ids})
    torch.cuda.synchronize()
    pl = next(iter(model._plans.values()))
    snaps.append({k: v.clone() for k, v in pl.bufs.items()})
names = list(snaps[0].keys())
shown = 0
for k in names:
    a, b, c = snaps[0][k], snaps[1][k], snaps[2][k]
    d01 = int((a != b).sum()); d12 = int((b != c).sum())
    if d01 or d12:
        rows = (a != b).any(dim=1).nonzero().flatten()
        print(f"{k:60s} shape {tuple(a.shape)}  call0!=call1: {d01}  call1!=call2: {d12}  first rows {rows[:5].tolist()} last {rows[-3:].tolist()} nrows {len(rows)}")
        shown += 1
        if shown >= 14: break
print("buffers:", len(names), "differing shown:", shown)
