"""A/B of the GEMM's CTA-pair modes inside ONE process (clock / box differences cancel): for every
hot-path shape, alternate cta_group::2 (default) and the round-1 W-multicast pairs (SDUSS_B200_NO_2CTA=1)
REPS times, each timing = 20 back-to-back launches replayed from a CUDA graph; medians, plus cuBLAS."""
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
shapes = [(14848, 4608, 1536), (14848, 1536, 1536), (14848, 6144, 1536), (14848, 1536, 6144), (1998, 4608, 1536),
          (1998, 1536, 1536), (1998, 1536, 6144), (1998, 6144, 1536), (8192, 8192, 8192), (40960, 320, 2880),
          (2560, 1280, 1280), (2560, 1280, 5120), (2560, 3840, 1280), (2560, 10240, 1280),
          (10240, 640, 640), (10240, 640, 2560), (10240, 1920, 640), (10240, 5120, 640),
          (512, 1280, 1280), (2048, 640, 640), (4096, 1536, 1536)]
print(f"{'M':>6s} {'N':>6s} {'K':>5s} | {'2cta us':>8s} {'TF/s':>6s} | {'mc us':>8s} {'TF/s':>6s} | {'cublas us':>9s} {'TF/s':>6s} | 2cta/mc  best/cublas")
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * .05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    os.environ["SDUSS_B200_NO_2CTA"] = "0"
    g2 = graph_of(lambda: ops.gemm(a, w, out=out, w_static=True))
    os.environ["SDUSS_B200_NO_2CTA"] = "1"
    g1 = graph_of(lambda: ops.gemm(a, w, out=out, w_static=True))
    os.environ["SDUSS_B200_NO_2CTA"] = "0"
    os.environ["SDUSS_B200_QUAD"] = "1"
    out4 = torch.empty_like(out)
    g4 = graph_of(lambda: ops.gemm(a, w, out=out4, w_static=True))
    os.environ["SDUSS_B200_QUAD"] = "0"
    gc = graph_of(lambda: torch.matmul(a, w.t(), out=out))
    t2, t1, tc, t4 = [], [], [], []
    for _ in range(REPS):
        t2.append(time_graph(*g2)); t1.append(time_graph(*g1)); tc.append(time_graph(*gc)); t4.append(time_graph(*g4))
    m2, m1, mc = statistics.median(t2), statistics.median(t1), statistics.median(tc)
    fl = 2 * M * N * K / 1e6
    m4 = statistics.median(t4)
    g2[0].replay(); torch.cuda.synchronize(); ok = torch.equal(out, out4)
    print(f"{M:6d} {N:6d} {K:5d} | {m2:8.1f} {fl/m2:6.0f} | {m1:8.1f} {fl/m1:6.0f} | {mc:9.1f} {fl/mc:6.0f} | {m2/m1:7.3f}  {min(m1,m2)/mc:7.3f} | quad {m4:7.1f} us {fl/m4:6.0f} TF/s ({m4/m2:5.3f} of 2cta) {'bit-equal' if ok else 'MISMATCH'}", flush=True)
