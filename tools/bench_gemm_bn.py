"""Tile width A/B inside one process: the library's choice vs 128 / 192 / 256-wide tiles forced
(SDUSS_B200_FORCE_BN, read per call) on the SDXL / SD3 token GEMM shapes; 20 launches per CUDA graph, medians."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops

REPS = 5


def graph_of(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    return g, n


def time_graph(g, n):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


dev = torch.device("cuda")
shapes = [(2560, 3840, 1280, 0), (2560, 1280, 1280, 2), (2560, 1280, 5120, 2), (2560, 10240, 1280, 4),
          (10240, 1920, 640, 0), (10240, 640, 640, 2), (10240, 640, 2560, 2), (10240, 5120, 640, 4),
          (5120, 3840, 1280, 0), (5120, 1280, 1280, 2), (7680, 3840, 1280, 0), (1024, 3840, 1280, 0),
          (14848, 4608, 1536, 0), (14848, 1536, 1536, 2), (2048, 4608, 1536, 0), (2048, 1536, 6144, 2)]
print(f"{'M':>6s} {'N':>6s} {'K':>5s} epi | {'default':>8s} | {'BN128':>8s} {'BN192':>8s} {'BN256':>8s}  (us)")
for (M, N, K, epi) in shapes:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * .05).bfloat16()
    No = N // 2 if epi == 4 else N
    out = torch.empty(M, No, device=dev, dtype=torch.bfloat16)
    resid = torch.randn(M, No, device=dev).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    kw = dict(bias=bias, w_static=True)
    if epi == 2:
        kw.update(epi=ops.EPI_GATE_RESID, resid=resid)
    elif epi == 4:
        kw.update(epi=ops.EPI_GEGLU)
    gs = {}
    ref = None
    for bn in ("0", "128", "192", "256"):
        os.environ["SDUSS_B200_FORCE_BN"] = bn
        gs[bn] = graph_of(lambda: ops.gemm(a, w, out, **kw))
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        else:
            assert torch.equal(ref, out), (M, N, K, bn)
    os.environ["SDUSS_B200_FORCE_BN"] = "0"
    t = {bn: [] for bn in gs}
    for _ in range(REPS):
        for bn, g in gs.items():
            t[bn].append(time_graph(*g))
    m = {bn: statistics.median(v) for bn, v in t.items()}
    print(f"{M:6d} {N:6d} {K:5d} {epi:3d} | {m['0']:8.1f} | {m['128']:8.1f} {m['192']:8.1f} {m['256']:8.1f}", flush=True)
