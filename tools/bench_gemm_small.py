"""Sub-wave GEMMs (a single 512^2 request): W-multicast pairs below the usual tile-count threshold
(SDUSS_B200_MC_MIN_PAIRS), in-process A/B, 20 back-to-back launches per timing, medians of 5."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops


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
shapes = [(512, 1280, 1280), (512, 3840, 1280), (512, 10240, 1280), (512, 1280, 5120), (2048, 640, 640), (2048, 1920, 640),
          (2048, 5120, 640), (2048, 640, 2560), (1024, 1280, 1280), (1024, 1280, 5120), (4096, 640, 640), (666, 1536, 1536),
          (2048, 1536, 1536), (2048, 4608, 1536), (2048, 6144, 1536), (2048, 1536, 6144), (154, 1536, 1536)]
print(f"{'M':>6s} {'N':>6s} {'K':>5s} | {'default us':>10s} | {'pairs>=1 us':>11s} | {'cublas us':>9s} | pairs/default")
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * .05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); out2 = torch.empty_like(out)
    os.environ.pop("SDUSS_B200_MC_MIN_PAIRS", None)
    g0 = graph_of(lambda: ops.gemm(a, w, out=out, w_static=True))
    os.environ["SDUSS_B200_MC_MIN_PAIRS"] = "1"
    g1 = graph_of(lambda: ops.gemm(a, w, out=out2, w_static=True))
    os.environ.pop("SDUSS_B200_MC_MIN_PAIRS", None)
    gc = graph_of(lambda: torch.matmul(a, w.t(), out=out))
    t0, t1, tc = [], [], []
    for _ in range(5):
        t0.append(time_graph(*g0)); t1.append(time_graph(*g1)); tc.append(time_graph(*gc))
    g0[0].replay(); g1[0].replay(); torch.cuda.synchronize()
    m0, m1, mc = statistics.median(t0), statistics.median(t1), statistics.median(tc)
    print(f"{M:6d} {N:6d} {K:5d} | {m0:10.1f} | {m1:11.1f} | {mc:9.1f} | {m1 / m0:6.3f} {'bit-equal' if torch.equal(out, out2) else 'MISMATCH'}", flush=True)
