"""Micro-benchmark: packed varlen attention on the SD3.5-medium config-2 joint shapes."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops

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
BOUNDED = os.environ.get("BOUNDED", "0") == "1"   # B200AttnExtra.bounded_logits (inputs here: logits ~ N(0, 1))
fn = lambda: ops.attn_varlen(sa, sb, *plan, 0.125, bounded=BOUNDED)
for _ in range(3): fn()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): fn()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
fl = sum(4 * (x + ctx) ** 2 * 64 * H for x in img)
print(f"b200 joint attention{' (bounded logits)' if BOUNDED else ''}: {ms:.3f} ms, {fl / ms / 1e9:.0f} TFLOP/s ({plan[2]} units, max_ctas {plan[4]})")
# torch SDPA per resolution (what the reference does)
def sdpa():
    ra = 0
    for i, x in enumerate(img):
        q = torch.cat([qa[ra:ra + x], qb[i * ctx:(i + 1) * ctx]]).view(1, x + ctx, 3, H, 64)
        torch.nn.functional.scaled_dot_product_attention(q[:, :, 0].transpose(1, 2), q[:, :, 1].transpose(1, 2), q[:, :, 2].transpose(1, 2))
        ra += x
for _ in range(3): sdpa()
torch.cuda.synchronize(); s.record()
for _ in range(10): sdpa()
e.record(); torch.cuda.synchronize()
ms2 = s.elapsed_time(e) / 10
print(f"torch SDPA loop (incl. cat copies): {ms2:.3f} ms, {fl / ms2 / 1e9:.0f} TFLOP/s")
