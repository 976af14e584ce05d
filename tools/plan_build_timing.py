"""What a NEW batch composition costs and what a step costs the HOST.
  use 1: plan tables + eager run (workspaces, tensor maps)      use 2: stream capture + first replay
  use 3+: graph replay. Also: host time per step (launch side only, no sync) vs GPU time per step.
python tools/plan_build_timing.py sd3|sdxl"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sduss_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "sd3"
dev = torch.device("cuda")
cfg, sd, pipe, make, call = bench.build_pipeline(which, dev)
del sd
model = pipe.model


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3


acc = {}
orig_capture, orig_eager = ops._capture, ops._run_eager


def cap(m, p, run=None):
    t = time.perf_counter(); g = orig_capture(m, p, run); acc["capture"] = (time.perf_counter() - t) * 1e3; return g


def eag(m, p, pro=None, run=None):
    torch.cuda.synchronize(); t = time.perf_counter(); orig_eager(m, p, pro, run); torch.cuda.synchronize()
    acc["eager"] = (time.perf_counter() - t) * 1e3


ops._capture, ops._run_eager = cap, eag
import importlib
plan_cls = importlib.import_module(type(model).__module__)._Plan
orig_init = plan_cls.__init__


def init(self, *a, **k):
    t = time.perf_counter(); orig_init(self, *a, **k); torch.cuda.synchronize(); acc["plan tables"] = (time.perf_counter() - t) * 1e3


plan_cls.__init__ = init
print(f"## {which}: cost of a new batch composition (ms; capture on the 2nd use)")
print(f"{'composition':34s} {'use1':>8s} {'(tables':>9s} {'eager)':>8s} {'use2':>8s} {'(capture)':>10s} {'replay':>8s} {'host/step':>10s}")
for spec in ({"512": 1}, {"768": 1}, {"1024": 1}, {"512": 1, "768": 1, "1024": 1}, {"512": 2, "1024": 2},
             {"512": 4, "768": 4, "1024": 4}, {"512": 3, "768": 5, "1024": 4}):
    reqs = make(spec, 300, 1)
    acc.clear()
    u1 = timed(lambda: call(reqs))
    tables, eager = acc.get("plan tables", 0.0), acc.get("eager", 0.0)
    u2 = timed(lambda: call(reqs))
    capture = acc.get("capture", 0.0)
    rep = min(timed(lambda: call(reqs)) for _ in range(5))
    # host side of a replayed step: how long the launching thread is busy per step
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        call(reqs)
    host = (time.perf_counter() - t) * 1e3 / 20
    torch.cuda.synchronize()
    name = "+".join(f"{n}x{r}" for r, n in spec.items())
    print(f"{name:34s} {u1:8.1f} {tables:9.1f} {eager:8.1f} {u2:8.1f} {capture:10.1f} {rep:8.2f} {host:10.2f}")
