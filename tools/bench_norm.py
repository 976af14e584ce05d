"""Micro-benchmark of the HBM-bound normalisation kernels on the shapes of the two models:
per-request GroupNorm(+SiLU) on the SDXL levels, LayerNorm (+affine / +AdaLN modulation) rows.
Two numbers per shape: 'hot' = the same buffers every launch (inputs stay in the 126 MB L2, as
they do in the model, where the producer kernel has just written them) and 'cold' = rotating over
enough buffers to exceed L2 (HBM-bound). Algorithmic bytes = one read + one write per output."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops
from sduss_b200.layout import LevelLayout

dev = torch.device("cuda")
PEAK = 6527.8  # MEASURED_PEAKS.json hbm GB/s on this pool


def timeit(fns, iters=30):
    """GPU time per call: the launches are replayed from a CUDA graph (a ctypes call costs ~12 us
    of host time, more than most of these kernels run)."""
    for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fns[i % len(fns)]()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us


def report(name, nbytes, hot, cold):
    print(f"{name:44s} {nbytes / 1e6:7.1f} MB  hot {hot:6.1f} us {nbytes / hot / 1e3:6.0f} GB/s | "
          f"cold {cold:6.1f} us {nbytes / cold / 1e3:6.0f} GB/s = {nbytes / cold / 1e3 / PEAK:4.2f} of HBM peak")


print("## GroupNorm + SiLU (SDXL config-1: 512^2 + 1024^2 with CFG -> latents 64^2 x2, 128^2 x2)")
for lvl, C in ((0, 320), (0, 640), (0, 960), (1, 640), (1, 1280), (1, 1920), (2, 1280), (2, 2560)):
    sizes = [(64 >> lvl, 64 >> lvl)] * 2 + [(128 >> lvl, 128 >> lvl)] * 2
    lay = LevelLayout(sizes, dev)
    nbuf = max(2, int(400e6 / (lay.T * C * 4)) + 1)
    xs = [torch.randn(lay.T, C, device=dev).bfloat16() for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev).bfloat16(), torch.randn(C, device=dev).bfloat16()
    ws = ops.groupnorm_workspace(lay.T, lay.L, dev)
    mk = lambda i: (lambda: ops.groupnorm_nhwc(xs[i], ys[i], g, b, lay.row_group, lay.lat_chunks, lay.L, ws, silu=True))
    report(f"groupnorm T={lay.T} C={C}", 2 * lay.T * C * 2, timeit([mk(0)]), timeit([mk(i) for i in range(nbuf)], 3 * nbuf))
    # statistics already left by the producing convolution's epilogue: finalize + apply, one read of x
    st = torch.ones(lay.n_tiles * 2 * C * 2, device=dev)
    mk2 = lambda i: (lambda: ops.groupnorm_from_conv_stats(xs[i], ys[i], g, b, lay.row_group, st, lay.lat_tiles, lay.L, ws, silu=True))
    report(f"  + stats from the conv epilogue", 2 * lay.T * C * 2, timeit([mk2(0)]), timeit([mk2(i) for i in range(nbuf)], 3 * nbuf))

print("## GroupNorm + SiLU on the VAE decoder levels (512^2 + 1024^2 images: latents 64^2 + 128^2, levels x4 and x8)")
for up, C in ((8, 128), (8, 256), (4, 256), (4, 512), (2, 512)):
    sizes = [(64 * up, 64 * up), (128 * up, 128 * up)]
    lay = LevelLayout(sizes, dev)
    xs = [torch.randn(lay.T, C, device=dev).bfloat16() for _ in range(2)]
    ys = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev).bfloat16(), torch.randn(C, device=dev).bfloat16()
    ws = ops.groupnorm_workspace(lay.T, lay.L, dev)
    mk = lambda i: (lambda: ops.groupnorm_nhwc(xs[i], ys[i], g, b, lay.row_group, lay.lat_chunks, lay.L, ws, silu=True))
    report(f"groupnorm T={lay.T} C={C}", 2 * lay.T * C * 2, timeit([mk(0)], 10), timeit([mk(0), mk(1)], 10))
    st = torch.ones(lay.n_tiles * 2 * C * 2, device=dev)
    mk2 = lambda i: (lambda: ops.groupnorm_from_conv_stats(xs[i], ys[i], g, b, lay.row_group, st, lay.lat_tiles, lay.L, ws, silu=True))
    report(f"  + stats from the conv epilogue", 2 * lay.T * C * 2, timeit([mk2(0)], 10), timeit([mk2(0), mk2(1)], 10))
    del xs, ys, st

print("## LayerNorm rows")
for T, D, kind in ((10240, 640, "affine"), (2560, 1280, "affine"), (14848, 1536, "mod"), (14848, 1536, "dual"),
                   (1998, 1536, "mod"), (59392, 1536, "mod")):
    nbuf = max(2, int(400e6 / (T * D * (6 if kind == "dual" else 4))) + 1)
    xs = [torch.randn(T, D, device=dev).bfloat16() for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    y2 = [torch.empty_like(x) for x in xs] if kind == "dual" else None
    L = 6
    rg = (torch.arange(T, device=dev, dtype=torch.int32) * L // T).int().contiguous()
    mod = torch.randn(L, 6 * D, device=dev).bfloat16()
    g, b = torch.randn(D, device=dev).bfloat16(), torch.randn(D, device=dev).bfloat16()
    if kind == "affine":
        mk = lambda i: (lambda: ops.layernorm_mod(xs[i], ys[i], eps=1e-5, gamma=g, beta=b))
    elif kind == "mod":
        mk = lambda i: (lambda: ops.layernorm_mod(xs[i], ys[i], eps=1e-6, mod=mod, row_group=rg, shift_col=0, scale_col=D))
    else:
        mk = lambda i: (lambda: ops.layernorm_mod(xs[i], ys[i], eps=1e-6, mod=mod, row_group=rg, shift_col=0, scale_col=D,
                                                  y2=y2[i], shift2_col=2 * D, scale2_col=3 * D))
    nbytes = T * D * 2 * (3 if kind == "dual" else 2)
    report(f"layernorm T={T} D={D} {kind}", nbytes, timeit([mk(0)]), timeit([mk(i) for i in range(nbuf)], 3 * nbuf))
