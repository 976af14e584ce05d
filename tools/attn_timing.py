"""Phase-level cycle breakdown of the attention softmax warps (needs an ATT_TIMING build:
SDUSS_B200_NVCC_EXTRA=-DATT_TIMING python -m sduss_b200.build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sduss_b200 import ops
from sduss_b200._lib import lib
dev = torch.device("cuda"); H = 24; C = H * 64; ctx = 333
img = [1024, 1024, 2304, 2304, 4096, 4096]
Ta, Tb = sum(img), len(img) * ctx
qa = torch.randn(Ta, 3 * C, device=dev).bfloat16(); qb = torch.randn(Tb, 3 * C, device=dev).bfloat16()
oa = torch.empty(Ta, C, device=dev, dtype=torch.bfloat16); ob = torch.empty(Tb, C, device=dev, dtype=torch.bfloat16)
seqs, ra = [], 0
for i, s in enumerate(img):
    seqs.append((ra, s, i * ctx, ctx, ra, s, i * ctx, ctx)); ra += s
plan = ops.build_attn_plan(seqs, dev, H)
sa = ops.attn_source(q=qa, k=qa, k_col=C, v=qa, v_col=2 * C, out=oa)
sb = ops.attn_source(q=qb, k=qb, k_col=C, v=qb, v_col=2 * C, out=ob)
for _ in range(3): ops.attn_varlen(sa, sb, *plan, 0.125)
torch.cuda.synchronize()
lib.b200_attn_debug_buffer.restype = ctypes.c_void_p
ptr = lib.b200_attn_debug_buffer()
n = 148
host = torch.empty(n * 32, dtype=torch.int64)
import ctypes as C_
cudart = C_.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
cudart.cudaMemcpy(C_.c_void_p(host.data_ptr()), C_.c_void_p(ptr), C_.c_size_t(n * 32 * 8), 2)
d = host.numpy().reshape(-1, 2, 16)
names = (["wait_s", "ref update / first max", "exp row (2 halves)", "st_p+arrive", "-", "-", "unit epilogue (total)"] if ops.ATTN_Q_TILE == 512 else
         ["wait_s", "ld_s+free", "max(+token)", "exp", "wait_o(+resc)", "st_p+arrive", "unit epilogue (total)"])
for t in (0, 1):
    x = d[:, t, :]
    x = x[x[:, 7] > 0]
    tiles = x[:, 7].sum()
    print(f"tile {'AB'[t]}: CTAs {len(x)}, steps {tiles}, cycles/step total {x[:, 8].sum() / tiles:.0f}")
    for i, nme in enumerate(names):
        print(f"   {nme:24s} {x[:, i].sum() / tiles:8.1f} clk/step")
    if x[:, 10].sum() > 0:
        units = x[:, 10].sum()
        print(f"   wait_s on the FIRST step of a unit: {x[:, 9].sum() / units:8.0f} clk per unit ({units} units, "
              f"{100 * x[:, 9].sum() / max(1, x[:, 0].sum()):.0f} % of all wait_s); steady-state wait_s "
              f"{(x[:, 0].sum() - x[:, 9].sum()) / max(1, tiles - units):.1f} clk/step; longest single wait_o {x[:, 11].max()} clk")
