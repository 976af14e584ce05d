"""Serving replay on N data-parallel B200 workers: BASELINE configs[2] (SDXL-base, 50 steps) and
configs[3] (SD3.5-medium, 28 steps), `bench.py --serve` and the `serve` sub-records of the default
bench line.

What is replayed (reference: tests/server/esymred_test.py:195-215 drives the server with
exp/<model>/qps_<q>.csv rows = arrival ms, resolution, steps):
  * arrivals  : the first rows of the reference's own trace exp/<model>/qps_8.0.csv (committed as
                tests/fixtures/traces/<model>_qps_8.0.csv), arrival times rescaled to the offered
                rate (a Poisson process stays Poisson under time scaling);
  * dispatch  : the reference's GreedyDispath rule (dispatcher/policy/greedy.py:16-36), live: a
                dispatcher thread in rank 0 places each request, at its arrival time, on the rank
                with the fewest outstanding pixels (sduss_b200.dp.DispatchBoard, shared memory);
  * batching  : per rank, `fcfs_mixed`-style continuous batching (worker/scheduler/policy/
                FCFS_Mixed.py): every dispatched request joins the running mixed-resolution batch
                at the next step boundary, up to max_batchsize 12 (scripts/paper/e2e.sh:73);
  * per step  : this repo's drop-in `denoising_step` over the whole mixed batch;
  * post stage: every request that finished at a step goes through the B200 VAE decoder
                (post_inference, row f-4) before it counts as served.
  * prepare   : every dispatched request goes through prepare_inference when it enters its runner
                (CLIP-L / CLIP-G / T5-XXL text encoders at the real model sizes with random-init weights
                on its prompt and negative prompt, initial latents, scheduler tables; row f-4);
Not simulated: tokenizer vocabularies (a hashing tokenizer stands in), HTTP, PIL conversion. One process per GPU, no collective on the data path; torch.distributed only
synchronises the start and gathers the per-rank results.
"""
import json
import os
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = {"sd3": 28, "sdxl": 50}
MAX_BATCH = 12


def load_trace(kind, n, qps):
    """(arrival seconds, resolution) x n from the committed reference trace, rescaled to `qps`."""
    rows = np.loadtxt(os.path.join(ROOT, "tests", "fixtures", "traces", f"{kind}_qps_8.0.csv"),
                      delimiter=",", skiprows=1)
    assert n <= rows.shape[0], f"trace fixture holds {rows.shape[0]} rows"
    t = rows[:n, 0] / 1000.0 * (8.0 / qps)
    return [(float(a), str(int(r))) for a, r in zip(t, rows[:n, 1])]


PROMPTS = ["a photo of an astronaut riding a horse on mars", "an oil painting of a lighthouse in a storm, dramatic light",
           "a bowl of ramen, studio photograph", "isometric illustration of a tiny city floating in the clouds",
           "portrait of an old fisherman, 85mm, shallow depth of field", "a red fox in the snow at dawn"]


class _Worker:
    def __init__(self, kind, dev, seed, prepare=True):
        import bench
        from sduss_b200 import synthetic
        from sduss_b200.vae import B200VAEDecoder, VAEDecoderConfig
        self.kind, self.dev = kind, dev
        self.cfg, sd, self.pipe, self.make, self.call = bench.build_pipeline(kind, dev)
        del sd
        vcfg = VAEDecoderConfig() if kind == "sdxl" else VAEDecoderConfig(
            latent_channels=16, scaling_factor=1.5305, shift_factor=0.0609, use_post_quant_conv=False)
        self.pipe.vae = B200VAEDecoder(synthetic.random_vae_state_dict(vcfg, dev), vcfg, device=dev)
        # a small pool of synthetic conditioning per resolution (stands in for the prepare stage)
        self.pool = {res: [self.make({res: 1}, STEPS[kind], seed + 17 * j)[res][0] for j in range(4)]
                     for res in ("512", "768", "1024")}
        self.seed = seed
        self.with_prepare = prepare
        if prepare:   # the prepare stage's text encoders (row f-4), random-init at the real model sizes
            enc, toks = synthetic.make_prompt_encoder(kind, dev, seed=7)
            self.pipe.attach_text_encoders(enc, toks)
        self.guidance = 7.0 if kind == "sd3" else 5.0

    def admit(self, rid, res):
        """A newly dispatched request enters the runner: prepare stage (text encoders on its prompt
        and negative prompt, initial latents, scheduler tables) when enabled, else synthetic state."""
        if not self.with_prepare:
            return self.new_request(rid, res)
        from types import SimpleNamespace
        sp = SimpleNamespace(prompt=PROMPTS[rid % len(PROMPTS)] + f" #{rid}", prompt_2=None, prompt_3=None,
                             negative_prompt="blurry, low quality", negative_prompt_2=None, negative_prompt_3=None,
                             num_inference_steps=STEPS[self.kind], height=int(res), width=int(res), latents=None)
        r = SimpleNamespace(request_id=rid, sampling_params=sp)
        if self.kind == "sd3":
            self.pipe.prepare_inference({res: [r]}, guidance_scale=self.guidance)
        else:
            self.pipe.prepare_inference({res: [r]}, guidance_scale=self.guidance)
        return r

    def new_request(self, rid, res):
        """A fresh request object with its own latents and scheduler state; conditioning tensors are
        cloned so the step's per-request conditioning cache sees a new prompt (as in serving)."""
        import copy
        proto = self.pool[res][rid % 4]
        r = copy.copy(proto)
        r.request_id = rid
        r.sampling_params = copy.copy(proto.sampling_params)
        r.sampling_params.latents = torch.randn_like(proto.sampling_params.latents)
        r.sampling_params.prompt_embeds = proto.sampling_params.prompt_embeds.clone()
        r.sampling_params.negative_prompt_embeds = proto.sampling_params.negative_prompt_embeds.clone()
        r.scheduler_states = copy.deepcopy(proto.scheduler_states)
        return r

    def sync(self):
        torch.cuda.synchronize()

    def throttle(self, depth=2):
        """The steps are asynchronous; let the host run at most `depth` steps ahead of the GPU so
        that a request that arrives now joins the batch at (nearly) the GPU's next step boundary."""
        q = self.__dict__.setdefault("_inflight", [])
        ev = torch.cuda.Event()
        ev.record()
        q.append(ev)
        if len(q) > depth:
            q.pop(0).synchronize()

    def n_plans(self):
        return len(self.pipe.model._plans)

    def post(self, finished):
        self.pipe.post_inference(finished, "pt")

    def warm(self):
        """Touch the kernels / allocator once per resolution so the replay does not time lazy init."""
        for res in ("512", "768", "1024"):
            r = self.admit(-1, res)
            r = self.admit(-1, res)   # second pass: the encoders' per-batch-size graphs are captured
            for _ in range(2):
                self.call({res: [r]})
            self.pipe.post_inference({res: [r]}, "pt")
        torch.cuda.synchronize()


def serve_run(kind, worker, trace, rank, world, dist, tag):
    """One replay. Returns the result record on rank 0 (None elsewhere)."""
    from sduss_b200.dp import DispatchBoard
    n_req, steps = len(trace), STEPS[kind]
    board_name = f"sduss_b200_{tag}_{os.environ.get('MASTER_PORT', '0')}_{kind}"
    board = DispatchBoard(board_name, n_req, world, create=True) if rank == 0 else None
    if dist is not None:
        dist.barrier()
    if board is None:
        board = DispatchBoard(board_name, n_req, world, create=False)
    # common start time
    t0 = [time.time() + 0.5]
    if dist is not None:
        dist.broadcast_object_list(t0, src=0, **({"device": worker.dev} if worker.dev.type == "cuda" else {}))
    t0 = t0[0]
    stop = threading.Event()

    def dispatcher():
        for i, (ta, res) in enumerate(trace):
            while not stop.is_set():
                dt = t0 + ta - time.time()
                if dt <= 0:
                    break
                time.sleep(min(dt, 0.005))
            board.dispatch(i, int(res))

    th = None
    if rank == 0:
        th = threading.Thread(target=dispatcher, daemon=True)
        th.start()
    import psutil
    psutil.cpu_percent(None)
    proc = psutil.Process()
    cpu0 = proc.cpu_times()
    plans0, n_steps, batch_sizes = worker.n_plans(), 0, []
    running, lat, done_at, scanned, mine_done, finish_t = [], [], [], 0, 0, t0
    waiting = []
    while True:
        # newly dispatched requests, in arrival order
        while scanned < n_req and board.assign[scanned] >= 0:
            if board.assign[scanned] == rank:
                waiting.append(scanned)
            scanned += 1
        while waiting and len(running) < MAX_BATCH:
            i = waiting.pop(0)
            running.append((i, worker.admit(i, trace[i][1])))
        if not running:
            if scanned >= n_req:
                break
            time.sleep(0.001)
            continue
        batch = {}
        for i, r in running:
            batch.setdefault(trace[i][1], []).append(r)
        worker.call(batch)
        worker.throttle()
        n_steps += 1
        batch_sizes.append(len(running))
        keep, fin = [], {}
        for i, r in running:
            if r.scheduler_states._step_index >= steps:
                fin.setdefault(trace[i][1], []).append((i, r))
            else:
                keep.append((i, r))
        if fin:
            worker.post({res: [r for _, r in v] for res, v in fin.items()})
            worker.sync()
            now = time.time()
            for res, v in fin.items():
                for i, _ in v:
                    lat.append(now - (t0 + trace[i][0]))
                    done_at.append(now - t0)
                    board.report_finished(rank, int(res))
                    mine_done += 1
            finish_t = now
        running = keep
    worker.sync()
    stop.set()
    cpu1 = proc.cpu_times()
    wall_rank = finish_t - t0
    host_cpu = psutil.cpu_percent(None)
    rec = {"rank": rank, "served": mine_done, "wall_s": wall_rank, "lat": lat, "done_at": done_at, "steps": n_steps,
           "mean_batch": float(np.mean(batch_sizes)) if batch_sizes else 0.0,
           "new_plans": worker.n_plans() - plans0,
           "proc_cpu_s": (cpu1.user + cpu1.system) - (cpu0.user + cpu0.system), "host_cpu_percent": host_cpu}
    allrec = [rec]
    if dist is not None:
        allrec = [None] * world
        dist.all_gather_object(allrec, rec)
    if th is not None:
        th.join(timeout=1.0)
    if dist is not None:
        dist.barrier()
    board.close()
    if rank != 0:
        return None
    wall = max(r["wall_s"] for r in allrec)
    lats = np.asarray([x for r in allrec for x in r["lat"]])
    offered = n_req / trace[-1][0]
    # steady state: completions between the 20th and the 80th percentile (ramp-up and drain excluded)
    done = np.sort(np.asarray([x for r in allrec for x in r["done_at"]]))
    lo, hi = int(0.2 * len(done)), int(0.8 * len(done))
    steady = (hi - lo) / (done[hi] - done[lo]) if hi > lo and done[hi] > done[lo] else n_req / wall
    return {"req_s": n_req / wall, "steady_req_s": float(steady), "offered_req_s": offered, "requests": n_req, "steps_per_request": steps,
            "wall_s": wall, "latency_s": {"mean": float(lats.mean()), "p50": float(np.percentile(lats, 50)),
                                          "p99": float(np.percentile(lats, 99))},
            "batch_steps_per_s": sum(r["steps"] for r in allrec) / wall,
            "request_steps_per_s": n_req * steps / wall,
            "mean_batch": float(np.mean([r["mean_batch"] for r in allrec])),
            "served_per_rank": [r["served"] for r in allrec],
            "new_compositions_per_rank": [r["new_plans"] for r in allrec],
            "runner_cpu_utilisation_per_rank": [round(r["proc_cpu_s"] / max(r["wall_s"], 1e-9), 3) for r in allrec],
            "host_cpu_percent": allrec[0]["host_cpu_percent"], "host_cores": os.cpu_count(),
            "max_batchsize": MAX_BATCH, "policy": "greedy dispatch + fcfs_mixed continuous batching",
            "stages": ("prepare (CLIP-L/G" + ("/T5-XXL" if kind == "sd3" else "") + " text encoders, random-init) + "
                       if getattr(worker, "with_prepare", False) else "") + "denoising steps + VAE decode (post_inference)",
            "trace": f"reference exp/{kind}/qps_8.0.csv rows 0..{n_req - 1}, arrivals rescaled to {offered:.2f} req/s"}


def serve_main(args, rank, world, local_rank):
    """bench.py --serve [--model sd3|sdxl|both] [--qps Q per GPU] [--requests R per GPU]."""
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    kinds = ["sd3", "sdxl"] if args.model == "both" else [args.model]
    out = {}
    for kind in kinds:
        w = _Worker(kind, dev, seed=1000 * rank)
        w.warm()
        per_gpu_q = args.qps if args.qps > 0 else 4.0
        n = (args.requests if args.requests > 0 else 48) * world
        rec = serve_run(kind, w, load_trace(kind, n, per_gpu_q * world), rank, world, dist, "serve")
        if rank == 0:
            out[kind] = rec
        del w
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if rank == 0:
        k0 = kinds[0]
        line = {"metric": f"served requests/s ({'+'.join(kinds)}; denoise + VAE decode, reference arrival trace)",
                "value": out[k0]["req_s"], "unit": "req/s", "n_gpus": world, "higher_is_better": True,
                "scaling": "weak", "data": "synthetic", "dtype": "bf16", "serve": out}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
