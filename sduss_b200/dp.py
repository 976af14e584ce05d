"""Data-parallel plumbing: replicas only, no collective on the denoise path.

sduss places every arriving request on ONE worker and never migrates it: `GreedyDispath`
(sduss/dispatcher/policy/greedy.py:16-36) picks the data-parallel rank with the fewest outstanding
pixels, sum of resolution^2 over its unfinished requests (dispatcher/request_pool.py:95-102).
`greedy_assign` is that rule for a static list; `DispatchBoard` is the same rule running live
between the N single-GPU bench processes of one box (bench.py --serve, tools/serve_replay.py):
a shared-memory board the dispatcher thread of rank 0 writes assignments to and every worker
reports finished pixels on. torch.distributed is only used for barriers and gathering results.
"""
from multiprocessing import shared_memory
from typing import List, Sequence

import numpy as np
import torch


def greedy_assign(resolutions: Sequence[int], world: int) -> List[List[int]]:
    """Mirror of GreedyDispath.dispatch_requests for requests that are all outstanding: each
    request (in arrival order) goes to the rank with the fewest outstanding pixels (sum of
    resolution^2). Returns request indices per rank."""
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i, r in enumerate(resolutions):
        k = min(range(world), key=lambda j: (load[j], j))
        out[k].append(i)
        load[k] += int(r) * int(r)
    return out


class DispatchBoard:
    """Shared-memory state of the live greedy dispatcher. Single-writer fields only, so no lock:
      assign[i]      rank of request i, -1 until dispatched        (written by the dispatcher)
      dispatched[k]  pixels ever sent to rank k                    (written by the dispatcher)
      finished[k]    pixels of the requests rank k has completed   (written by worker k)
    outstanding pixels of rank k = dispatched[k] - finished[k], what get_pixels_all_dp_rank()
    returns in the reference."""

    def __init__(self, name: str, n_requests: int, world: int, create: bool):
        self.n, self.world = n_requests, world
        nbytes = 8 * (n_requests + 2 * world)
        if create:
            try:  # a crashed earlier run may have left the segment behind
                stale = shared_memory.SharedMemory(name=name, create=False)
                stale.close()
                stale.unlink()
            except FileNotFoundError:
                pass
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes)
        if not create:
            # Python < 3.13 registers attached segments with this process' resource tracker, which
            # would unlink the owner's segment at exit; only the creator owns it
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        buf = np.ndarray((n_requests + 2 * world,), dtype=np.int64, buffer=self.shm.buf)
        self.assign = buf[:n_requests]
        self.dispatched = buf[n_requests:n_requests + world]
        self.finished = buf[n_requests + world:]
        self.owner = create
        if create:
            self.assign[:] = -1
            self.dispatched[:] = 0
            self.finished[:] = 0

    def dispatch(self, i: int, resolution: int) -> int:
        """Dispatcher side: place request i (greedy.py:24-34). Returns the chosen rank."""
        load = self.dispatched - self.finished
        k = int(np.argmin(load))  # first minimum: dict order of the reference's pixels_by_dp_rank
        self.dispatched[k] += int(resolution) ** 2
        self.assign[i] = k
        return k

    def report_finished(self, rank: int, resolution: int) -> None:
        self.finished[rank] += int(resolution) ** 2

    def close(self):
        self.assign = self.dispatched = self.finished = None
        self.shm.close()
        if self.owner:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


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
