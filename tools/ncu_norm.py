"""Launches each normalisation kernel a few times on one model shape (no CUDA graph), for
`ncu --set full -k regex:'gn_|ln_mod' ...` captures (tools/ncu_summary.py turns them into text)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops
from sduss_b200.layout import LevelLayout

dev = torch.device("cuda")
# GroupNorm + SiLU, SDXL level 0 (C = 320) and level 1 (C = 1280), config-1 latents
for lvl, C in ((0, 320), (1, 1280)):
    sizes = [(64 >> lvl, 64 >> lvl)] * 2 + [(128 >> lvl, 128 >> lvl)] * 2
    lay = LevelLayout(sizes, dev)
    x = torch.randn(lay.T, C, device=dev).bfloat16(); y = torch.empty_like(x)
    g, b = torch.randn(C, device=dev).bfloat16(), torch.randn(C, device=dev).bfloat16()
    ws = ops.groupnorm_workspace(lay.T, lay.L, dev)
    for _ in range(2):
        ops.groupnorm_nhwc(x, y, g, b, lay.row_group, lay.lat_chunks, lay.L, ws, silu=True)
# AdaLN-modulated LayerNorm, SD3.5-M config-2 image tokens
T, D, L = 14848, 1536, 6
x = torch.randn(T, D, device=dev).bfloat16(); y = torch.empty_like(x); y2 = torch.empty_like(x)
rg = (torch.arange(T, device=dev, dtype=torch.int32) * L // T).int().contiguous()
mod = torch.randn(L, 6 * D, device=dev).bfloat16()
for _ in range(2):
    ops.layernorm_mod(x, y, eps=1e-6, mod=mod, row_group=rg, shift_col=0, scale_col=D)
    ops.layernorm_mod(x, y, eps=1e-6, mod=mod, row_group=rg, shift_col=0, scale_col=D, y2=y2, shift2_col=2 * D, scale2_col=3 * D)
torch.cuda.synchronize()
print("ok")
