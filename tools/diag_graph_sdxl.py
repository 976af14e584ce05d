"""Diagnostic: SDXL UNet forward repeated on identical inputs (eager first call, then CUDA-graph
replays unless SDUSS_B200_NO_GRAPH=1): element mismatches vs the first call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200.unet import B200UNet, UNetConfig
from sduss_b200.synthetic import random_unet_state_dict
cuda = torch.device("cuda")
size = sys.argv[1] if len(sys.argv) > 1 else "full"
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if size == "tiny":
    from dataclasses import asdict
    from oracle import sdxl_unet as ox
    oc = ox.sdxl_tiny_config()
    d = asdict(oc); ctx = d.pop("context_len")
    cfg = UNetConfig(**d); cfg.context_len = ctx
    sd = {k: v.to(torch.bfloat16) for k, v in ox.init_unet_weights(oc, 0).items()}
    spec = {"256": 2, "512": 2}
else:
    cfg = UNetConfig(); cfg.context_len = 77
    sd = random_unet_state_dict(cfg, cuda, seed=0)
    spec = {"512": 2, "1024": 2}
model = B200UNet(sd, cfg, device=cuda)
g = torch.Generator().manual_seed(3)
hs = {r: (torch.randn(n, 4, int(r) // 8, int(r) // 8, generator=g)).to(cuda, torch.bfloat16) for r, n in spec.items()}
L = sum(spec.values())
ehs = torch.randn(L, cfg.context_len, cfg.cross_attention_dim, generator=g).to(cuda, torch.bfloat16)
te = torch.randn(L, cfg.pooled_dim, generator=g).to(cuda, torch.bfloat16)
ids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * L).to(cuda, torch.bfloat16)
t = torch.full((L,), 999.0, device=cuda)
def run():
    out = model(hs, t, encoder_hidden_states=ehs, added_cond_kwargs={"text_embeds": te, "time_ids": ids})[0]
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in out.items()}
outs = [run() for _ in range(runs)]
for i, o in enumerate(outs):
    print(i, {k: (int((o[k] != outs[0][k]).sum()), float((o[k].float() - outs[0][k].float()).abs().max())) for k in o},
          "vs prev", {k: int((o[k] != outs[i - 1][k]).sum()) for k in o} if i else "")
