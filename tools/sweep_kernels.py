"""BASELINE configs[4]: attention / conv roofline characterisation over resolution 256^2..1024^2
and batch 1..32 (uniform batches plus the mixed config-1 / config-2 batches). CUDA events,
3 warm-up + 10 timed launches per point. Prints TFLOP/s and the fraction of the measured sustained
bf16 peak (MEASURED_PEAKS.json)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sduss_b200 import ops
from sduss_b200.layout import LevelLayout

dev = torch.device("cuda")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    PEAK = 1400.0


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def attn_point(img_lens, ctx, H):
    C = H * 64
    Ta, Tb = sum(img_lens), len(img_lens) * ctx
    qa = torch.randn(Ta, 3 * C, device=dev).bfloat16()
    oa = torch.empty(Ta, C, device=dev, dtype=torch.bfloat16)
    seqs, ra = [], 0
    if ctx:
        qb = torch.randn(Tb, 3 * C, device=dev).bfloat16()
        ob = torch.empty(Tb, C, device=dev, dtype=torch.bfloat16)
    for i, s in enumerate(img_lens):
        seqs.append((ra, s, i * ctx, ctx, ra, s, i * ctx, ctx) if ctx else (ra, s, 0, 0, ra, s, 0, 0)); ra += s
    plan = ops.build_attn_plan(seqs, dev, H)
    sa = ops.attn_source(q=qa, k=qa, k_col=C, v=qa, v_col=2 * C, out=oa)
    sb = ops.attn_source(q=qb, k=qb, k_col=C, v=qb, v_col=2 * C, out=ob) if ctx else None
    ms = timeit(lambda: ops.attn_varlen(sa, sb, *plan, 0.125))
    fl = sum(4.0 * (x + ctx) ** 2 * 64 * H for x in img_lens)
    return ms, fl / ms / 1e9


def conv_point(sizes, cin, cout):
    lay = LevelLayout(sizes, dev)
    x = torch.randn(lay.T, cin, device=dev).bfloat16()
    w = (torch.randn(cout, 9 * cin, device=dev) * 0.02).bfloat16()
    b = torch.zeros(cout, device=dev).bfloat16()
    out = torch.empty(lay.T, cout, device=dev, dtype=torch.bfloat16)
    maps = ops.conv3x3_encode_maps(x, cin, lay.desc_host, 1)
    omaps = ops.conv3x3_encode_maps(out, cout, lay.desc_host, 1)
    ms = timeit(lambda: ops.conv3x3(maps, lay.tiles, lay.n_tiles, lay.desc, cin, cout, 1, w, out, out_maps=omaps,
                                    epi=ops.EPI_BIAS, bias=b))
    fl = 2.0 * 9 * cin * cout * lay.T
    return ms, fl / ms / 1e9


print(f"# peak = {PEAK:.0f} TFLOP/s (measured sustained bf16)")
print("## SD3.5-medium joint attention (24 heads x 64, 333 context tokens), CFG doubles the latents")
print(f"{'resolution':>10s} {'requests':>8s} {'latents':>8s} {'ms':>8s} {'TFLOP/s':>8s} {'of peak':>8s}")
for res in (256, 512, 768, 1024):
    S = (res // 16) ** 2
    for nreq in (1, 2, 4, 8, 16, 32):
        if S * 2 * nreq > 140000: continue
        ms, tf = attn_point([S] * (2 * nreq), 333, 24)
        print(f"{res:>10d} {nreq:>8d} {2 * nreq:>8d} {ms:8.3f} {tf:8.0f} {tf / PEAK:8.2f}")
ms, tf = attn_point([1024, 1024, 2304, 2304, 4096, 4096], 333, 24)
print(f"{'mixed':>10s} {'config-2':>8s} {6:>8d} {ms:8.3f} {tf:8.0f} {tf / PEAK:8.2f}")
print("## SDXL self-attention, level 1 (10 heads, (res/16)^2 tokens) and level 2 (20 heads, (res/32)^2 tokens)")
print(f"{'resolution':>10s} {'requests':>8s} {'level':>8s} {'ms':>8s} {'TFLOP/s':>8s} {'of peak':>8s}")
for res in (256, 512, 768, 1024):
    for nreq in (1, 4, 16, 32):
        for lvl, H, S in ((1, 10, (res // 16) ** 2), (2, 20, (res // 32) ** 2)):
            ms, tf = attn_point([S] * (2 * nreq), 0, H)
            print(f"{res:>10d} {nreq:>8d} {lvl:>8d} {ms:8.3f} {tf:8.0f} {tf / PEAK:8.2f}")
print("## SDXL resnet 3x3 convolutions (implicit GEMM), latent side = res/8 >> level")
print(f"{'resolution':>10s} {'requests':>8s} {'Cin->Cout':>12s} {'ms':>8s} {'TFLOP/s':>8s} {'of peak':>8s}")
for res in (256, 512, 768, 1024):
    for nreq in (1, 4, 16):
        for lvl, cin, cout in ((0, 320, 320), (1, 640, 640), (2, 1280, 1280), (0, 960, 320), (2, 2560, 1280)):
            side = (res // 8) >> lvl
            if side * side * 2 * nreq * max(cin, cout) * 2 > 8e9: continue
            ms, tf = conv_point([(side, side)] * (2 * nreq), cin, cout)
            print(f"{res:>10d} {nreq:>8d} {f'{cin}->{cout}':>12s} {ms:8.3f} {tf:8.0f} {tf / PEAK:8.2f}")
ms, tf = conv_point([(64, 64), (64, 64), (128, 128), (128, 128)], 320, 320)
print(f"{'mixed':>10s} {'config-1':>8s} {'320->320':>12s} {ms:8.3f} {tf:8.0f} {tf / PEAK:8.2f}")
