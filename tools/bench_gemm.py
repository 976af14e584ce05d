"""Micro-benchmark of the tcgen05 GEMM vs torch.matmul (cuBLAS) on the hot-path shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops

def timeit(fn, n=20):
    """GPU time per call, replayed from a CUDA graph (host launch cost exceeds the small shapes)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

dev = torch.device("cuda")
for (M, N, K) in [(14848, 4608, 1536), (14848, 1536, 1536), (14848, 6144, 1536), (14848, 1536, 6144),
                  (1998, 4608, 1536), (40960, 320, 2880), (8192, 8192, 8192),
                  # SDXL level-2 / level-1 token GEMMs (config-1: 2560 / 10240 rows)
                  (2560, 1280, 1280), (2560, 1280, 5120), (1998, 1536, 1536), (1998, 1536, 6144), (2560, 3840, 1280), (2560, 10240, 1280),
                  (10240, 640, 640), (10240, 640, 2560), (10240, 1920, 640)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * .05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t1 = timeit(lambda: ops.gemm(a, w, out=out))
    t2 = timeit(lambda: torch.matmul(a, w.t(), out=out))
    fl = 2 * M * N * K / 1e9
    print(f"M={M} N={N} K={K}: b200 {t1:.3f} ms {fl/t1:.0f} TF/s | cublas {t2:.3f} ms {fl/t2:.0f} TF/s", flush=True)

print("## gate * y + residual epilogue (EPI_GATE_RESID; out == resid buffer as in the models) vs the plain epilogue", flush=True)
for (M, N, K, L) in [(14848, 1536, 1536, 6), (14848, 1536, 6144, 6), (1998, 1536, 1536, 6), (2560, 1280, 1280, 0),
                     (2560, 1280, 5120, 0), (10240, 640, 640, 0), (10240, 640, 2560, 0)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * .05).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    x = torch.randn(M, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    kw = {}
    if L:
        kw = dict(gate=torch.randn(L, N, device=dev).bfloat16(),
                  row_group=(torch.arange(M, device=dev) * L // M).int())
    t0 = timeit(lambda: ops.gemm(a, w, out=out, bias=bias))
    t1 = timeit(lambda: ops.gemm(a, w, out=x, bias=bias, epi=ops.EPI_GATE_RESID, resid=x, **kw))
    fl = 2 * M * N * K / 1e9
    print(f"M={M} N={N} K={K}: plain {t0*1e3:.1f} us {fl/t0:.0f} TF/s | gate+resid {t1*1e3:.1f} us {fl/t1:.0f} TF/s", flush=True)
