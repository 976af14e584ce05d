"""A handful of eager launches of one GEMM shape (for `ncu --set full -k regex:gemm_bf16`).
python tools/ncu_one_gemm.py M N K [resid]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
resid = len(sys.argv) > 4
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
b = torch.randn(N, device="cuda").bfloat16()
x = torch.randn(M, N, device="cuda").bfloat16()
for _ in range(6):
    if resid:
        ops.gemm(a, w, x, bias=b, epi=ops.EPI_GATE_RESID, resid=x)
    else:
        ops.gemm(a, w, x, bias=b)
torch.cuda.synchronize()
