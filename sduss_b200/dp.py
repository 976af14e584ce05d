"""Data-parallel plumbing for the multi-GPU bench: replicas only, no collective on the denoise
path (reference: sduss/dispatcher/policy/greedy.py:16-36 balances requests by
outstanding pixels). Used by bench.py; torch.distributed is only the timing barrier/reduction."""
from typing import Dict, List, Sequence

import torch


def greedy_assign(resolutions: Sequence[int], world: int) -> List[List[int]]:
    """Mirror of GreedyDispath.dispatch_requests: each request (in arrival order) goes to the
    rank with the fewest outstanding pixels (sum of resolution^2). Returns request indices per
    rank."""
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i, r in enumerate(resolutions):
        k = min(range(world), key=lambda j: (load[j], j))
        out[k].append(i)
        load[k] += int(r) * int(r)
    return out


def max_over_ranks(ms: float, device=None) -> float:
    """Max of a per-rank duration over all ranks (identity when not distributed)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_steps_per_s(steps: int, world: int, max_ms_total: float) -> float:
    """Whole-job throughput: every rank ran `steps` steps of its own batch in max_ms_total."""
    return world * steps / (max_ms_total / 1e3)
