"""Locates read-before-write / hidden-state bugs: runs the same forward twice on identical inputs
(eager, no PDL) and compares, launch by launch, checksums of every tensor a launch touches.
The first launch whose tensors differ between call 0 and call 1 is printed.
Usage: SDUSS_B200_NO_GRAPH=1 SDUSS_B200_NO_PDL=1 python tools/trace_stale.py [sdxl|sd3] [tiny|full]"""
import os, sys, inspect
os.environ.setdefault("SDUSS_B200_NO_GRAPH", "1")
os.environ.setdefault("SDUSS_B200_NO_PDL", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sduss_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
size = sys.argv[2] if len(sys.argv) > 2 else "full"
cuda = torch.device("cuda")
trace = []

def csum(t):
    return float(t.detach().reshape(-1).to(torch.float64).nan_to_num(nan=12345.0, posinf=1e30, neginf=-1e30).sum().item())

def tensors_of(x, out):
    if torch.is_tensor(x):
        out.append(x)
    elif isinstance(x, (list, tuple)):
        for y in x: tensors_of(y, out)
    elif hasattr(x, "_keep"):
        tensors_of(list(x._keep), out)

def wrap(name, fn):
    def inner(*a, **k):
        r = fn(*a, **k)
        ts = []
        if name == "copy_cols":   # only the written columns of the destination are defined here
            ts = [a[0], a[1][:, :a[2]]]
        else:
            tensors_of(list(a) + list(k.values()) + [r], ts)
        trace.append((name, [(tuple(t.shape), csum(t)) for t in ts if t.is_cuda and t.dtype != torch.uint8]))
        return r
    return inner

for name, fn in list(vars(ops).items()):
    if inspect.isfunction(fn) and fn.__module__ == ops.__name__ and not name.startswith("_") and name not in (
            "run_plan", "graphs_enabled", "attn_source", "build_attn_plan", "check", "groupnorm_workspace", "conv3x3_encode_maps"):
        setattr(ops, name, wrap(name, fn))

g = torch.Generator().manual_seed(3)
if which == "sdxl":
    from sduss_b200.unet import B200UNet, UNetConfig
    from sduss_b200.synthetic import random_unet_state_dict
    if size == "tiny":
        from oracle import sdxl_unet as ox
        from dataclasses import asdict
        oc = ox.sdxl_tiny_config()
        d = asdict(oc); ctx = d.pop("context_len")
        cfg = UNetConfig(**d); cad = cfg.cross_attention_dim
        sd = {k: v.to(torch.bfloat16) for k, v in ox.init_unet_weights(oc, 0).items()}
    else:
        cfg = UNetConfig(); ctx, cad = 77, cfg.cross_attention_dim
        sd = random_unet_state_dict(cfg, cuda, seed=0)
    cfg.context_len = ctx
    model = B200UNet(sd, cfg, device=cuda)
    spec = {"512": 2, "1024": 2} if size == "full" else {"256": 2, "512": 2}
    hs = {r: torch.randn(n, 4, int(r) // 8, int(r) // 8, generator=g).to(cuda, torch.bfloat16) for r, n in spec.items()}
    L = sum(spec.values())
    ehs = torch.randn(L, ctx, cad, generator=g).to(cuda, torch.bfloat16)
    te = torch.randn(L, cfg.pooled_dim, generator=g).to(cuda, torch.bfloat16)
    ids = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * L).to(cuda, torch.bfloat16)
    t = torch.full((L,), 999.0, device=cuda)
    run = lambda: model(hs, t, encoder_hidden_states=ehs, added_cond_kwargs={"text_embeds": te, "time_ids": ids})[0]
else:
    raise SystemExit("only sdxl wired up")

# poison the allocator cache so that fresh workspaces do not start from zeros
junk = [torch.full((1 << 26,), float("nan"), device=cuda, dtype=torch.bfloat16) for _ in range(8)]
del junk
traces = []
for i in range(2):
    trace.clear()
    out = run()
    torch.cuda.synchronize()
    traces.append(list(trace))
a, b = traces
print("launches:", len(a), len(b))
n = 0
for i, ((na, ta), (nb, tb)) in enumerate(zip(a, b)):
    if na != nb or ta != tb:
        print(f"launch {i}: {na}")
        for (sa, ca), (sb, cb) in zip(ta, tb):
            print(f"    {sa}  call0={ca:.6g}  call1={cb:.6g}  {'DIFF' if ca != cb else ''}")
        n += 1
        if n >= 3:
            break
print("differing launches shown:", n)
