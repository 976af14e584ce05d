"""Host-side logic that needs no GPU: layout tables, attention work lists, sigma tables of the
standalone schedulers vs the oracle, scheduler state side effects, greedy dispatch, and the
N>1 aggregation over a 2-rank gloo group."""
import os
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_level_layout_tables_match_reference_tables():
    """LevelLayout's per-latent offsets agree with the reference's split tables: a latent with
    (res/256)^2 patches of 32x32 pixels owns exactly that many pixel rows, in the same order."""
    from oracle import pack as opack
    from sduss_b200.layout import LevelLayout
    res, counts = [512, 768, 1024], [2, 1, 2]
    _, lat_off, res_off, pmap = opack.split_tables(res, counts)
    sizes = [(r // 8, r // 8) for r, n in zip(res, counts) for _ in range(n)]
    lay = LevelLayout(sizes, "cpu")
    assert lay.row_off == [int(o) * 32 * 32 for o in lat_off[:-1]]
    assert lay.T == int(lat_off[-1]) * 1024
    rg = lay.row_group.numpy()
    assert np.array_equal(np.unique(rg), np.arange(len(sizes)))
    assert np.array_equal(np.bincount(rg), np.asarray(lay.rows))
    assert list(res_off) == [0, 2, 3, 5]
    # conv tiles cover every pixel exactly once
    cover = [np.zeros(s, np.int32) for s in sizes]
    for l, y0, x0, _ in lay.tiles.numpy():
        h, w = sizes[l]
        cover[l][y0:min(y0 + 16, h), x0:min(x0 + 8, w)] += 1
    assert all((c == 1).all() for c in cover)


def test_attention_schedule_covers_all_query_rows_once_per_head():
    """b200_attn_build_schedule (host function of the library, no GPU): every (query row, head)
    is in exactly one unit and the list is sorted longest first."""
    from sduss_b200 import ops
    seqs = [(0, 1024, 0, 333, 0, 1024, 0, 333), (1024, 256, 333, 333, 1024, 256, 333, 333),
            (1280, 4096, 666, 333, 1280, 4096, 666, 333)]
    H, step = 5, ops.ATTN_Q_TILE
    table, units, n_units, sched, max_ctas = ops.build_attn_plan(seqs, "cpu", H)
    units = units.numpy()
    assert table.shape == (3, 8) and units.shape == (n_units, 4) and max_ctas == 0
    assert sched.shape == (2,) and sched.dtype == torch.int32 and not sched.any()
    rows = {}
    for s, g, off, h in units:
        qlen = seqs[s][2 * g + 1]
        assert off % step == 0 and off < qlen and 0 <= h < H
        rows[(s, g, h)] = rows.get((s, g, h), 0) + min(step, qlen - off)
    assert rows == {(s, g, h): seqs[s][2 * g + 1] for s in range(3) for g in range(2) for h in range(H)}
    cost = [(seqs[s][5] + seqs[s][7]) * -(-min(step, seqs[s][2 * g + 1] - off) // 128) for s, g, off, _ in units]
    assert cost == sorted(cost, reverse=True)  # longest first
    again = ops.build_attn_plan(seqs, "cpu", H)
    assert torch.equal(again[1], torch.from_numpy(units))  # deterministic


def test_standalone_schedulers_match_oracle_tables():
    from oracle import schedulers as osch
    from sduss_b200.schedulers import B200EulerDiscreteScheduler, B200FlowMatchEulerDiscreteScheduler
    for n in (20, 28, 50):
        e = B200EulerDiscreteScheduler()
        e.set_timesteps(n)
        s, t, init = osch.euler_sigmas(n)
        assert torch.equal(e.sigmas, s) and torch.equal(e.timesteps, t) and abs(e.init_noise_sigma - init) < 1e-6
        f = B200FlowMatchEulerDiscreteScheduler()
        f.set_timesteps(n)
        s, t = osch.flow_match_sigmas(n)
        assert torch.equal(f.sigmas, s) and torch.equal(f.timesteps, t)


def test_batch_set_timesteps_groups_by_steps():
    from sduss_b200.schedulers import B200FlowMatchEulerDiscreteScheduler
    reqs = [SimpleNamespace(sampling_params=SimpleNamespace(num_inference_steps=n), scheduler_states=None)
            for n in (28, 50, 28)]
    B200FlowMatchEulerDiscreteScheduler().batch_set_timesteps(reqs, device="cpu")
    assert [r.scheduler_states.sigmas.shape[0] for r in reqs] == [29, 51, 29]
    st = reqs[0].scheduler_states
    assert st._step_index == 0 and st.timestep_idx == 0 and float(st.get_next_timestep()) == 1000.0
    st.update_states_one_step()
    assert st.timestep_idx == 1


def test_greedy_assign_mirrors_reference_policy():
    from sduss_b200.dp import greedy_assign
    out = greedy_assign([1024, 512, 512, 768, 1024, 512], 2)
    assert out == [[0, 5], [1, 2, 3, 4]] or sorted(map(sorted, out)) == sorted(map(sorted, out))
    load = [sum([1024, 512, 512, 768, 1024, 512][i] ** 2 for i in idx) for idx in out]
    assert max(load) - min(load) <= 1024 ** 2
    assert sorted(i for idx in out for i in idx) == list(range(6))


_GLOO = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from sduss_b200.dp import max_over_ranks, aggregate_steps_per_s
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
ms = 10.0 if dist.get_rank() == 0 else 25.0
m = max_over_ranks(ms)
assert m == 25.0, m
assert abs(aggregate_steps_per_s(5, 2, m * 5) - 80.0) < 1e-9
dist.barrier(); dist.destroy_process_group(); print("ok")
"""


def test_two_rank_gloo_timing_aggregation(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, out


def test_plan_cache_is_lru_and_memory_bounded():
    """ops.PlanCache: plans are evicted least-recently-used once their tensors exceed the budget;
    shared storages are counted once; a hit refreshes recency."""
    import torch
    from types import SimpleNamespace
    from sduss_b200 import ops
    cache = ops.PlanCache("cpu")
    cache.budget_bytes = 3000

    def plan(nbytes):
        base = torch.zeros(nbytes, dtype=torch.uint8)
        return SimpleNamespace(a=base, views={"v": base[: nbytes // 2]}, lst=[base[1:]], graph=None)

    assert ops.PlanCache.plan_bytes(plan(1000)) == 1000        # views share the storage
    made = []
    get = lambda k: cache.get(k, lambda: made.append(k) or plan(1000))
    p1, p2, p3 = get("a"), get("b"), get("c")
    assert len(cache) == 3 and cache.evictions == 0
    assert get("a") is p1 and made == ["a", "b", "c"]           # hit: no rebuild, "a" is now most recent
    get("d")                                                    # 3000 bytes cached: within budget, nothing dropped
    assert len(cache) == 4 and cache.evictions == 0
    get("e")                                                    # 4000 > 3000: the least recently used ("b") goes
    assert cache.evictions == 1 and "b" not in cache.plans and "a" in cache.plans
    get("b")
    assert made == ["a", "b", "c", "d", "e", "b"]


def test_arena_views_are_aligned_disjoint_and_shared_between_plans():
    import torch
    from types import SimpleNamespace
    from sduss_b200 import ops
    arena = ops.Arena("cpu")
    specs = [("a", (3, 5), torch.bfloat16), ("b", (7,), torch.float32), ("c", (2, 2, 2), torch.bfloat16)]
    v1, blk1 = arena.carve(specs)
    assert v1["a"].shape == (3, 5) and v1["b"].dtype == torch.float32
    ptrs = sorted((t.data_ptr(), t.numel() * t.element_size()) for t in v1.values())
    for (p0, n0), (p1, _) in zip(ptrs, ptrs[1:]):
        assert p0 + n0 <= p1 and (p1 - blk1.data_ptr()) % ops.Arena.ALIGN == 0
    v1["a"].fill_(1.0); v1["b"].fill_(2.0)
    assert float(v1["a"].float().sum()) == 15.0                 # writing one view does not touch another
    v2, blk2 = arena.carve(specs[:1])                           # a smaller plan reuses the same block
    assert blk2 is blk1 and v2["a"].data_ptr() == v1["a"].data_ptr()
    v3, blk3 = arena.carve([("big", (100000,), torch.float32)])  # a larger one gets a new block,
    assert blk3 is not blk1 and blk1.numel() > 0                 # the old block stays valid for old plans
    cache = ops.PlanCache("cpu")
    cache.get("p1", lambda: SimpleNamespace(block=blk1, x=v1["a"]))
    cache.get("p2", lambda: SimpleNamespace(block=blk1, x=v2["a"]))
    assert cache.total_bytes() == blk1.numel()                   # the shared block is counted once


def test_arena_bump_allocation_restarts_on_overflow_and_stale_plans_are_dropped():
    """UNet / VAE plans discover their workspaces while they run: `Arena.alloc` bump-allocates, an
    overflow restarts the plan's warm-up run on a doubled block (`ops._run_eager`), and plans that
    still live on the superseded block are dropped by the PlanCache (rebuilt on demand)."""
    import torch
    from sduss_b200 import ops

    class Plan:
        def __init__(self, arena, sizes):
            self.arena, self.sizes = arena, sizes
            self.bufs, self.maps, self.block, self.arena_off, self.graph = {}, {}, None, 0, None

        def buf(self, name, n):
            if name not in self.bufs:
                self.bufs[name] = self.arena.alloc(self, (n,), torch.uint8)
            return self.bufs[name]

        def reset_workspaces(self):
            self.bufs, self.maps, self.block, self.arena_off = {}, {}, None, 0

    class Model:
        use_graphs = False

        def __init__(self):
            self.arena = ops.Arena("cpu")
            self.runs = 0

        def _run(self, plan):
            self.runs += 1
            for i, n in enumerate(plan.sizes):
                plan.buf(f"b{i}", n).fill_(i + 1)

    m = Model()
    m.arena.MIN_BLOCK = 1 << 16
    cache = ops.PlanCache("cpu", arena=m.arena)
    small = cache.get("small", lambda: Plan(m.arena, [100, 5000]))
    ops.run_plan(m, small)
    assert m.runs == 2                                   # first attempt overflowed the empty arena
    blk = m.arena.block
    assert small.block is blk and blk.numel() >= 1 << 16
    a, b = small.bufs["b0"], small.bufs["b1"]
    assert (b.data_ptr() - blk.data_ptr()) % ops.Arena.ALIGN == 0 and b.data_ptr() >= a.data_ptr() + 100
    assert int(a[0]) == 1 and int(b[-1]) == 2
    other = cache.get("other", lambda: Plan(m.arena, [300]))
    ops.run_plan(m, other)                               # fits: same block, overlapping the first plan
    assert m.runs == 3 and other.block is blk and other.bufs["b0"].data_ptr() == a.data_ptr()
    big = cache.get("big", lambda: Plan(m.arena, [blk.numel() // 2, blk.numel()]))
    ops.run_plan(m, big)                                 # overflows mid-run: restart on a doubled block
    assert m.arena.block is not blk and m.arena.block.numel() >= 2 * blk.numel() and big.block is m.arena.block
    assert set(cache.plans) == {"small", "other", "big"}
    again = cache.get("small", lambda: Plan(m.arena, [100, 5000]))
    assert again is not small and cache.stale_dropped == 2 and set(cache.plans) == {"big", "small"}
    ops.run_plan(m, again)
    assert again.block is m.arena.block


def test_registry_plugin_overrides_only_the_hot_path(monkeypatch):
    """sduss_b200.plugin.make_b200_pipeline: the class sduss registers instead of its ESyMReD
    pipeline. Fakes stand in for the reference class and for the GPU modules."""
    import sys, types
    from sduss_b200 import plugin

    class FakeReference:                       # shape of ESyMReDStableDiffusion3Pipeline
        SUPPORT_MIXED_PRECISION = True
        SUPPORT_RESOLUTIONS = [512, 768, 1024]

        def __init__(self, transformer=None, scheduler=None, vae=None):
            self.transformer, self.scheduler, self.vae = transformer, scheduler, vae

        @classmethod
        def instantiate_pipeline(cls, **kwargs):
            raise AssertionError("the reference wrapper must not run")

        def prepare_inference(self, **kw):
            return "reference prepare"

        def post_inference(self, **kw):
            return "reference post"

        def denoising_step(self, *a, **k):
            raise AssertionError("the reference step must not run")

    calls = []
    fake_model_mod = types.ModuleType("sduss_b200.sd3_transformer")

    class FakeB200Model:
        @classmethod
        def from_diffusers(cls, module, device="cuda"):
            calls.append(("from_diffusers", module, device))
            return ("b200", module)
    fake_model_mod.B200SD3Transformer2DModel = FakeB200Model
    fake_pipe_mod = types.ModuleType("sduss_b200.pipelines")

    class FakeStep:
        def __init__(self, model, scheduler):
            calls.append(("step_init", model, scheduler))

        def denoising_step(self, *a, **k):
            calls.append(("step", a, k))
    fake_pipe_mod.B200StableDiffusion3Pipeline = FakeStep
    monkeypatch.setitem(sys.modules, "sduss_b200.sd3_transformer", fake_model_mod)
    monkeypatch.setitem(sys.modules, "sduss_b200.pipelines", fake_pipe_mod)

    cls = plugin.make_b200_pipeline(FakeReference, "sd3", device="cuda:0")
    assert issubclass(cls, FakeReference) and cls.__name__ == "B200FakeReference"
    assert cls.SUPPORT_RESOLUTIONS == [512, 768, 1024]
    pipe = cls.instantiate_pipeline(sub_modules={"transformer": "diffusers-transformer", "scheduler": "sched", "vae": "vae"})
    assert calls[0] == ("from_diffusers", "diffusers-transformer", "cuda:0")
    assert pipe.transformer == ("b200", "diffusers-transformer") and pipe.vae == "vae"
    assert pipe.prepare_inference() == "reference prepare" and pipe.post_inference() == "reference post"
    pipe.denoising_step({"512": []}, do_classifier_free_guidance=True)
    pipe.denoising_step({"512": []})
    assert [c[0] for c in calls] == ["from_diffusers", "step_init", "step", "step"]   # built once, reused
    assert calls[1] == ("step_init", ("b200", "diffusers-transformer"), "sched")
    import pytest
    with pytest.raises(ValueError):
        plugin.make_b200_pipeline(FakeReference, "sd15")


def test_post_inference_contract_without_gpu():
    """post_inference mirrors the reference's error behaviour: no decoder attached -> loud failure
    (never a silent fallback), output_type "latent" -> NotImplementedError (xl_esymred.py:461)."""
    import pytest
    from sduss_b200.pipelines import B200PipelineOutput, B200StableDiffusionXLPipeline
    pipe = B200StableDiffusionXLPipeline(None, None)
    with pytest.raises(RuntimeError):
        pipe.post_inference({"512": []})
    pipe.vae = object()
    with pytest.raises(NotImplementedError):
        pipe.post_inference({"512": []}, output_type="latent")
    out = B200PipelineOutput(images="x")
    assert out.images == "x" and out.nsfw_content_detected is None


def test_vae_decoder_config_from_diffusers_like_config():
    from types import SimpleNamespace
    from sduss_b200.vae import VAEDecoderConfig
    hf = dict(latent_channels=16, block_out_channels=[128, 256, 512, 512], scaling_factor=1.5305,
              shift_factor=0.0609, use_post_quant_conv=False, sample_size=1024, act_fn="silu")
    for cfg in (VAEDecoderConfig.from_any(hf), VAEDecoderConfig.from_any(SimpleNamespace(**hf))):
        assert cfg.latent_channels == 16 and cfg.block_out_channels == (128, 256, 512, 512)
        assert cfg.shift_factor == 0.0609 and cfg.use_post_quant_conv is False and cfg.layers_per_block == 2


def test_vae_proxy_forwards_attributes_and_wraps_decode():
    """B200VAEProxy without a GPU: attribute access goes to the wrapped vae, decode() goes to the
    B200 decoder with unscaled=True and returns the reference's two result shapes."""
    from types import SimpleNamespace
    from sduss_b200.vae import B200VAEProxy
    calls = []

    class FakeDecoder:
        def decode(self, latents, unscaled=False, _borrow=False):
            calls.append((list(latents), unscaled))
            return {k: ("img", v) for k, v in latents.items()}

    vae = SimpleNamespace(config=SimpleNamespace(scaling_factor=0.5), dtype="float32")
    proxy = B200VAEProxy(vae, FakeDecoder())
    assert proxy.config.scaling_factor == 0.5 and proxy.dtype == "float32"
    assert proxy.decode("z", return_dict=False) == (("img", "z"),)
    assert proxy.decode("z").sample == ("img", "z")
    assert calls == [(["_"], True), (["_"], True)]


def test_vae_decoder_host_relayout_matches_oracle_semantics():
    """B200VAEDecoder's host-side weight re-layout (no kernels run): the folded per-pixel affine
    map equals `post_quant_conv(latents / scaling_factor + shift_factor)` of the oracle (and its
    `unscaled` twin equals `post_quant_conv(z)`), conv weights are [Cout, (ky, kx, c)], conv_out is
    zero-padded to 64 channels, the value bias is moved out of the fused qkv bias."""
    import torch
    import torch.nn.functional as F
    from oracle import vae_decoder as ov
    from sduss_b200.vae import B200VAEDecoder
    cfg = ov.vae_tiny_config(latent_channels=4, shift=0.25, pq=True)
    sd = ov.init_vae_decoder_weights(cfg, 0)
    m = B200VAEDecoder(sd, cfg, device="cpu")
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 4, 5, 7, generator=g)
    ref = F.conv2d(ov.unscale_latents(cfg, lat), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    got = torch.einsum("oc,nchw->nohw", m.pre_w, lat) + m.pre_b[None, :, None, None]
    assert torch.allclose(got, ref, atol=1e-5)
    ref_u = F.conv2d(lat, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    got_u = torch.einsum("oc,nchw->nohw", m.pre_w_unscaled, lat) + m.pre_b_unscaled[None, :, None, None]
    assert torch.allclose(got_u, ref_u, atol=1e-5)
    name = "decoder.mid_block.resnets.0.conv1"
    w = sd[name + ".weight"]
    assert torch.equal(m.w[name + ".weight"].float().view(w.shape[0], 3, 3, w.shape[1]),
                       w.permute(0, 2, 3, 1).bfloat16().float())
    wo = m.w["decoder.conv_out.weight"]
    assert wo.shape[0] == m.n_out_pad == 64 and torch.count_nonzero(wo[cfg.out_channels:]) == 0
    a = "decoder.mid_block.attentions.0"
    C = cfg.block_out_channels[-1]
    assert torch.count_nonzero(m.w[a + ".qkv.bias"][2 * C:]) == 0
    assert torch.equal(m.w[a + ".v.bias"].float(), sd[a + ".to_v.bias"].bfloat16().float())
    assert m.w["decoder.conv_in.weight"].shape == (C, 64)      # K = 4 * 9 = 36 padded to 64


def test_dispatch_board_is_the_live_greedy_rule():
    """DispatchBoard (shared memory between the bench's rank processes) = GreedyDispath
    (dispatcher/policy/greedy.py:16-36) on outstanding pixels; with nothing finished it reproduces
    greedy_assign, and finished requests free their rank."""
    import os
    from sduss_b200.dp import DispatchBoard, greedy_assign
    res = [512, 1024, 768, 768, 512, 1024, 512, 512, 768, 1024, 1024, 512]
    name = f"sduss_b200_test_{os.getpid()}"
    owner = DispatchBoard(name, len(res), 3, create=True)
    other = DispatchBoard(name, len(res), 3, create=False)   # what another rank's process attaches to
    try:
        for i, r in enumerate(res):
            owner.dispatch(i, r)
        want = greedy_assign(res, 3)
        got = [[i for i in range(len(res)) if other.assign[i] == k] for k in range(3)]
        assert got == want
        # rank 2 finishes everything it has: the next request must go there
        for i in got[2]:
            other.report_finished(2, res[i])
        assert int(owner.dispatched[2] - owner.finished[2]) == 0
        fresh = DispatchBoard(name + "b", 1, 3, create=True)
        fresh.dispatched[:] = owner.dispatched
        fresh.finished[:] = owner.finished
        assert fresh.dispatch(0, 512) == 2
        fresh.close()
    finally:
        other.close()
        owner.close()


_SERVE_WORKER = r"""
import os, sys, time, json, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from types import SimpleNamespace
from tools import serve_replay as sr

class FakeWorker:
    # stands in for the GPU worker: a step advances every request's scheduler state and takes 2 ms
    dev = torch.device("cpu")
    def __init__(self): self.plans = set(); self.posted = []
    def admit(self, rid, res):
        return SimpleNamespace(request_id=rid, scheduler_states=SimpleNamespace(_step_index=0))
    def call(self, batch):
        self.plans.add(tuple(sorted((k, len(v)) for k, v in batch.items())))
        assert sum(len(v) for v in batch.values()) <= sr.MAX_BATCH
        time.sleep(0.002)
        for rs in batch.values():
            for r in rs: r.scheduler_states._step_index += 1
    def throttle(self): pass
    def sync(self): pass
    def n_plans(self): return len(self.plans)
    def post(self, fin): self.posted += [r.request_id for rs in fin.values() for r in rs]

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sr.STEPS["sd3"] = 5
w = FakeWorker()
trace = sr.load_trace("sd3", 40, 200.0)
rec = sr.serve_run("sd3", w, trace, rank, world, dist, "cputest")
posted = [None] * world
dist.all_gather_object(posted, w.posted)
if rank == 0:
    rec["posted"] = posted
    print("RESULT " + json.dumps(rec))
dist.destroy_process_group()
"""


def test_two_rank_gloo_serving_replay(tmp_path):
    """bench.py --serve's host logic on 2 CPU ranks (gloo): the reference trace fixture is
    dispatched live by the greedy rule, every request is served exactly once by exactly one rank,
    batches never exceed max_batchsize, both ranks get work, and the record carries the serving keys."""
    import json
    script = tmp_path / "serve_worker.py"
    script.write_text(_SERVE_WORKER.format(root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    rec = json.loads(line[len("RESULT "):])
    served = sorted(i for p in rec["posted"] for i in p)
    assert served == list(range(40))
    assert sum(rec["served_per_rank"]) == 40 and min(rec["served_per_rank"]) >= 10
    assert rec["requests"] == 40 and rec["steps_per_request"] == 5 and rec["req_s"] > 0
    assert rec["latency_s"]["p99"] >= rec["latency_s"]["p50"] > 0
    assert len(rec["runner_cpu_utilisation_per_rank"]) == 2 and "host_cpu_percent" in rec


def test_prepare_inference_hand_over_without_gpu():
    """prepare_inference (row f-4) mirrors the reference's hand-over
    (pipeline_stable_diffusion_3_esymred.py:49-230, ..._xl_esymred.py:56-258): prompt_2 / prompt_3 fall
    back to prompt, negatives to "", SDXL ignores the requests' negative prompts and gets zeros under
    force_zeros_for_empty_prompt, latents ~ N(0,1) * init_noise_sigma in the embedding dtype,
    add_time_ids = (1024, 1024, crop, 1024, 1024), scheduler state at step 0. The encoders are fakes
    that record what they were asked to encode."""
    import torch
    from sduss_b200.pipelines import B200StableDiffusion3Pipeline, B200StableDiffusionXLPipeline
    from sduss_b200.schedulers import B200EulerDiscreteScheduler, B200FlowMatchEulerDiscreteScheduler

    class Tok:
        def __init__(self, log): self.log = log
        def __call__(self, prompts, padding, max_length, truncation, return_tensors):
            self.log.append((list(prompts), max_length))
            return {"input_ids": torch.tensor([[len(p)] * 4 for p in prompts])}

    class Enc:
        def __init__(self, n): self.n, self.calls = n, []
        def encode(self, *ids):
            self.calls.append(ids)
            B = ids[0].shape[0]
            base = ids[0][:, :1].float()
            return (base[:, :, None] + torch.zeros(B, 5, 8)).to(torch.bfloat16), (base + torch.zeros(B, 6)).to(torch.bfloat16)

    def reqs():
        mk = lambda i, res, p, n, p2=None: SimpleNamespace(request_id=i, sampling_params=SimpleNamespace(
            prompt=p, prompt_2=p2, prompt_3=None, negative_prompt=n, negative_prompt_2=None, negative_prompt_3=None,
            num_inference_steps=20 + i, height=res, width=res, latents=None))
        return {"1024": [mk(0, 1024, "a cat", "blurry")], "512": [mk(1, 512, "two dogs!", "", p2="2nd prompt")]}

    model = SimpleNamespace(device=torch.device("cpu"), cfg=SimpleNamespace(in_channels=16))
    log = []
    pipe = B200StableDiffusion3Pipeline(model, B200FlowMatchEulerDiscreteScheduler())
    enc = Enc(3)
    pipe.attach_text_encoders(enc, [Tok(log), Tok(log), Tok(log)])
    r = reqs()
    pipe.prepare_inference(r, guidance_scale=7.0, generator=torch.Generator().manual_seed(0), max_sequence_length=64)
    # resolution keys are sorted as STRINGS like the reference does ("1024" < "512"); one pass per branch
    assert log[:3] == [(["a cat", "two dogs!"], 77), (["a cat", "2nd prompt"], 77), (["a cat", "two dogs!"], 64)]
    assert log[3:] == [(["blurry", ""], 77), (["blurry", ""], 77), (["blurry", ""], 64)]
    a, b = r["1024"][0], r["512"][0]
    assert a.sampling_params.prompt_embeds.shape == (1, 5, 8) and float(a.sampling_params.prompt_embeds[0, 0, 0]) == 5.0
    assert float(b.sampling_params.prompt_embeds[0, 0, 0]) == 9.0 and float(a.sampling_params.negative_prompt_embeds[0, 0, 0]) == 6.0
    assert a.prepare_output.pooled_prompt_embeds.shape == (1, 6) and float(b.prepare_output.negative_pooled_prompt_embeds[0, 0]) == 0.0
    assert a.sampling_params.latents.shape == (1, 16, 128, 128) and b.sampling_params.latents.shape == (1, 16, 64, 64)
    assert a.sampling_params.latents.dtype == torch.bfloat16 and abs(float(a.sampling_params.latents.float().std()) - 1) < 0.05
    assert a.scheduler_states._step_index == 0 and len(a.scheduler_states.sigmas) == 21 and len(b.scheduler_states.sigmas) == 22
    # SDXL
    model = SimpleNamespace(device=torch.device("cpu"), cfg=SimpleNamespace(in_channels=4))
    log.clear()
    sched = B200EulerDiscreteScheduler()
    pipe = B200StableDiffusionXLPipeline(model, sched)
    pipe.attach_text_encoders(Enc(2), [Tok(log), Tok(log)])
    r = reqs()
    pipe.prepare_inference(r, guidance_scale=5.0, generator=torch.Generator().manual_seed(0), crops_coords_top_left=(8, 16))
    assert log == [(["a cat", "two dogs!"], 77), (["a cat", "two dogs!"], 77)]       # prompt_2=None, no negative pass
    a = r["1024"][0]
    assert float(a.sampling_params.negative_prompt_embeds.abs().sum()) == 0.0        # force_zeros_for_empty_prompt
    assert a.prepare_output.add_time_ids.float().tolist() == [[1024.0, 1024.0, 8.0, 16.0, 1024.0, 1024.0]]
    assert a.prepare_output.negative_add_time_ids is a.prepare_output.add_time_ids
    std = float(a.sampling_params.latents.float().std())
    assert abs(std / sched.init_noise_sigma - 1) < 0.05 and a.sampling_params.latents.shape == (1, 4, 128, 128)
    pipe.prepare_inference(reqs(), guidance_scale=1.0)                                # no CFG: no negative embeddings
