"""Where a GEMM launch spends its time (debug build: python tools/build_variant.py trace gemm_sm100.cu -DGEMM_TRACE,
run with SDUSS_B200_LIB=sduss_b200/variants/libsduss_b200_trace.so). For CTA 0 of each of 20 back-to-back
launches (CUDA graph, programmatic dependent launch): cycles from kernel entry to the end of the prologue,
to the return of griddepcontrol.wait, to the first full stage, to the last K step, to the accumulator hand-over,
to the last store, to the exit; and the gap between consecutive launches (globaltimer)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sduss_b200 import ops, _lib

lib = _lib.lib
lib.b200_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
dev = torch.device("cuda")
N_L = 20
for (M, N, K, ws) in [(2560, 1280, 1280, True), (2560, 1280, 1280, False), (2560, 1280, 5120, True), (512, 1280, 1280, True),
                      (2048, 640, 640, True), (2560, 10240, 1280, True), (14848, 1536, 1536, True)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * .05).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    f = lambda: ops.gemm(a, w, out=out, w_static=ws)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N_L): f()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    buf = np.zeros((64, 10), dtype=np.uint64); n = ctypes.c_uint(0)
    assert lib.b200_debug_gemm_trace(buf.ctypes.data, ctypes.byref(n)) == 0
    rows = [(n.value - N_L + i) & 63 for i in range(N_L)]
    t = buf[rows].astype(np.int64)
    d = t[:, 1:8] - t[:, 0:1]
    med = np.median(d[3:], axis=0)
    life = np.median((t[3:, 9] - t[3:, 8]))
    gap = np.median(t[4:, 8] - t[3:-1, 9])
    print(f"M={M} N={N} K={K} w_static={ws}: {s.elapsed_time(e) / N_L * 1e3:.1f} us per launch | CTA 0 cycles from entry: "
          f"prologue {med[0]:.0f}, wait returns {med[1]:.0f}, first stage full {med[2]:.0f}, last K step {med[3]:.0f}, "
          f"accumulator ready {med[4]:.0f}, stores done {med[5]:.0f}, exit {med[6]:.0f} | CTA 0 lifetime {life / 1e3:.2f} us, "
          f"previous exit -> this entry {gap / 1e3:.2f} us", flush=True)
