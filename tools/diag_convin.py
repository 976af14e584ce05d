import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops
from sduss_b200.unet import B200UNet, UNetConfig
from sduss_b200.synthetic import random_unet_state_dict
cuda = torch.device("cuda")
cfg = UNetConfig(); cfg.context_len = 77
model = B200UNet(random_unet_state_dict(cfg, cuda, seed=0), cfg, device=cuda)
g = torch.Generator().manual_seed(3)
spec = {"512": 2, "1024": 2}
hs = {r: torch.randn(n, 4, int(r) // 8, int(r) // 8, generator=g).to(cuda, torch.bfloat16) for r, n in spec.items()}
pl = model._plan(hs, 77)
for res, _, _, _ in pl.comp:
    pl.stage_in[res].copy_(hs[res])
l0 = pl.levels[0]
w = model.w
outs, colss = [], []
for i in range(4):
    cols = torch.empty(l0.T, model.k_in_pad, device=cuda, dtype=torch.bfloat16)
    ops.pack_im2col3x3(pl.in_ptr, l0.desc, pl.L, l0.max_pixels, cfg.in_channels, cols)
    x = torch.empty(l0.T, 320, device=cuda, dtype=torch.bfloat16)
    ops.gemm(cols, w["conv_in.weight"], x, bias=w["conv_in.bias"])
    torch.cuda.synchronize()
    colss.append(cols); outs.append(x)
print("cols diffs vs 0:", [int((c != colss[0]).sum()) for c in colss])
print("gemm diffs vs 0:", [int((o != outs[0]).sum()) for o in outs])
# gemm alone on identical cols
y = []
for i in range(4):
    x = torch.empty(l0.T, 320, device=cuda, dtype=torch.bfloat16)
    ops.gemm(colss[0], w["conv_in.weight"], x, bias=w["conv_in.bias"])
    torch.cuda.synchronize(); y.append(x)
print("gemm-only diffs vs 0:", [int((o != y[0]).sum()) for o in y])
ref = (colss[0].float() @ w["conv_in.weight"].float().t() + w["conv_in.bias"].float())
for i, o in enumerate(y):
    err = (o.float() - ref).abs()
    bad = (err > 0.05 * ref.abs().clamp(min=1.0)).nonzero()
    print(i, "max err", float(err.max()), "bad elements", len(bad), bad[:6].tolist())
d = (y[1] != y[0]).nonzero()
print("sample diffs:", [(int(r), int(c), float(y[0][r, c]), float(y[1][r, c]), float(ref[r, c])) for r, c in d[:8]])
print("nan in cols:", bool(torch.isnan(colss[0].float()).any()), "cols beyond 36 nonzero:", int((colss[0][:, 36:] != 0).sum()))
