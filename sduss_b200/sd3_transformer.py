"""B200-native drop-in for sduss' PatchSD3Transformer2DModel
(reference: sduss/model_executor/modules/SD3Transformer.py:25-262).

Same forward signature and return type (dict resolution -> [n_r, 16, h, w] in, tuple(dict) out,
conditioning rows ordered resolution by resolution), but no 256-px chunk machinery: all image
tokens of all latents live in ONE packed buffer [sum S_i, 1536] bf16 (row for row identical to
the reference's split_sample_sd3 chunk stack), context tokens in [L*333, 1536], and every op
is a single launch of a hand-written sm_100a kernel through the C ABI (sduss_b200.ops):
tcgen05 GEMMs with fused bias / q-k RMSNorm / GELU / gate+residual epilogues, one packed
varlen joint-attention launch, warp-per-row LayerNorm+AdaLN modulation. AdaLN modulation
vectors for all 24 blocks come from one up-front GEMM per step (they depend only on temb and
are computed per latent, not per chunk as the reference does).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops


def _G(*a, **k):
    """ops.gemm against WEIGHTS (never written on the stream): W tiles may be fetched ahead of the wait
    on the previous kernel."""
    return ops.gemm(*a, w_static=True, **k)


@dataclass
class SD3Config:
    num_layers: int = 24
    num_attention_heads: int = 24
    attention_head_dim: int = 64
    in_channels: int = 16
    out_channels: int = 16
    patch_size: int = 2
    pos_embed_max_size: int = 384
    sample_size: int = 128
    joint_attention_dim: int = 4096
    caption_projection_dim: int = 1536
    pooled_projection_dim: int = 2048
    dual_attention_layers: List[int] = field(default_factory=lambda: list(range(13)))

    @property
    def inner_dim(self):
        return self.num_attention_heads * self.attention_head_dim

    @classmethod
    def from_any(cls, cfg):
        """Accepts this class, a diffusers FrozenDict/config object or a plain dict."""
        if isinstance(cfg, cls):
            return cls(**{k: getattr(cfg, k) for k in cls.__dataclass_fields__})
        get = cfg.get if hasattr(cfg, "get") else lambda k, d=None: getattr(cfg, k, d)
        kw = {}
        for k in cls.__dataclass_fields__:
            v = get(k, None)
            if v is not None:
                kw[k] = list(v) if k == "dual_attention_layers" else v
        return cls(**kw)


class _Plan:
    """Everything that depends only on the batch composition ((resolution, count) tuple):
    descriptor tables, attention work lists, gathered positional rows and workspaces."""

    def __init__(self, model: "B200SD3Transformer2DModel", comp, ctx_len: int):
        cfg, dev = model.cfg, model.device
        D, p = cfg.inner_dim, cfg.patch_size
        self.comp = comp
        self.ctx_len = ctx_len
        lat = []  # (res_key, index in res, h_tok, w_tok)
        for res, n, h, w in comp:
            for i in range(n):
                lat.append((res, i, h // p, w // p))
        self.L = L = len(lat)
        S = [ht * wt for _, _, ht, wt in lat]
        off = np.concatenate([[0], np.cumsum(S)]).astype(np.int64)
        self.T, self.Tc = int(off[-1]), L * ctx_len
        self.max_tokens = max(S)
        bf = dict(device=dev, dtype=torch.bfloat16)
        # static input / output staging per resolution (stable pointers -> graph-capturable)
        # (one flat allocation each, per-resolution views: the fused CFG + scheduler kernel
        # addresses predictions by element offset from flat_out)
        numel_in = [n * cfg.in_channels * h * w for _, n, h, w in comp]
        numel_out = [n * cfg.out_channels * h * w for _, n, h, w in comp]
        # Per-step workspaces, carved from the model's shared arena (see ops.Arena): every one of
        # them is fully rewritten by each forward before it is read.
        T, Tc = self.T, self.Tc
        b16, f32 = torch.bfloat16, torch.float32
        specs = [("flat_in", (sum(numel_in),), b16), ("flat_out", (sum(numel_out),), b16),
                 ("tokens", (T, cfg.in_channels * p * p), b16), ("x", (T, D), b16), ("xa", (T, D), b16),
                 ("xn", (T, D), b16),
                 ("xn2", (T, D), b16), ("c", (Tc, D), b16), ("cn", (Tc, D), b16), ("qkv", (T, 3 * D), b16),
                 ("qkv_c", (Tc, 3 * D), b16), ("att", (T, D), b16), ("att_c", (Tc, D), b16),
                 ("ff", (T, 4 * D), b16), ("ff_c", (Tc, 4 * D), b16),
                 ("out_tok", (T, p * p * cfg.out_channels), b16), ("mod", (L, model.mod_cols), b16),
                 ("t32", (L,), f32), ("tsin", (L, 256), b16), ("e1", (L, D), b16), ("e2", (L, D), b16),
                 ("e3", (L, D), b16), ("temb", (L, D), b16), ("pooled", (L, cfg.pooled_projection_dim), b16),
                 ("ehs", (Tc, cfg.joint_attention_dim), b16)]
        views, self.block = model.arena.carve(specs)
        for name, t in views.items():
            setattr(self, name, t)
        self.stage_in, self.stage_out, self.in_elem_off, self.out_elem_off = {}, {}, {}, {}
        oi = oo = 0
        for (res, n, h, w), ni, no in zip(comp, numel_in, numel_out):
            self.stage_in[res] = self.flat_in[oi:oi + ni].view(n, cfg.in_channels, h, w)
            self.stage_out[res] = self.flat_out[oo:oo + no].view(n, cfg.out_channels, h, w)
            self.in_elem_off[res], self.out_elem_off[res] = oi, oo
            oi, oo = oi + ni, oo + no
        in_ptr, out_ptr, desc = [], [], []
        for l, (res, i, ht, wt) in enumerate(lat):
            in_ptr.append(self.stage_in[res][i].data_ptr())
            out_ptr.append(self.stage_out[res][i].data_ptr())
            desc.append((int(off[l]), wt, ht, 0))
        self.in_ptr = torch.tensor(in_ptr, dtype=torch.int64).to(dev)
        self.out_ptr = torch.tensor(out_ptr, dtype=torch.int64).to(dev)
        self.desc = torch.tensor(desc, dtype=torch.int32).to(dev)
        self.row_group = torch.from_numpy(np.repeat(np.arange(L, dtype=np.int32), S)).to(dev)
        self.row_group_ctx = torch.from_numpy(
            np.repeat(np.arange(L, dtype=np.int32), ctx_len)).to(dev)
        # positional rows: cropped sincos table per latent, gathered once per composition
        pos = []
        M = cfg.pos_embed_max_size
        for _, _, ht, wt in lat:
            top, left = (M - ht) // 2, (M - wt) // 2
            rows = (torch.arange(top, top + ht)[:, None] * M + torch.arange(left, left + wt)[None]).reshape(-1)
            pos.append(rows)
        self.pos_rows = model.pos_table[torch.cat(pos).to(dev)].contiguous()
        # attention work lists
        joint = [(int(off[l]), S[l], l * ctx_len, ctx_len, int(off[l]), S[l], l * ctx_len, ctx_len)
                 for l in range(L)]
        self.joint_plan = ops.build_attn_plan(joint, dev, model.cfg.num_attention_heads)
        selfp = [(int(off[l]), S[l], 0, 0, int(off[l]), S[l], 0, 0) for l in range(L)]
        self.self_plan = ops.build_attn_plan(selfp, dev, model.cfg.num_attention_heads)
        # attention sources (pointers are static)
        self.src_img = ops.attn_source(q=self.qkv, q_col=0, k=self.qkv, k_col=D, v=self.qkv,
                                       v_col=2 * D, out=self.att)
        self.src_ctx = ops.attn_source(q=self.qkv_c, q_col=0, k=self.qkv_c, k_col=D, v=self.qkv_c,
                                       v_col=2 * D, out=self.att_c)
        self.graph = None
        self.graph_launches = 0
        self.cache = None          # _PatchCache, allocated when the plan first runs with the patch cache
        self.cached_state = None   # graph / use counters of the cached way of running this plan
        self.S = S


class _CachedState:
    graph = None
    graph_launches = 0
    uses = 0
    warm = False


class _PatchCache:
    """What a plan keeps ACROSS steps when the patch cache is on: per block the
    block output `xout`, the block input of the previous step `xprev` (MSE feature), the fused q|k|v
    projections `qkv` (`qkv2` for the second attention of dual blocks) -- rows of clean patches are
    simply not rewritten, so they still hold what the patch had when it was last computed -- plus
    the per-patch skip counters, the masks and the per-latent validity flags the step fills in.
    Everything is tied to the plan's row layout: a request keeps its cache while it stays in the same
    slot of the same batch composition from one step to the next (tags), otherwise its patches are
    recomputed once and the cache restarts (zero-copy; the reference keys Python dicts of tensors by
    request id and patch and re-stacks them every block, cache_manager.py:58-99)."""

    PATCH = 256

    def __init__(self, model, pl):
        dev, D = model.device, model.cfg.inner_dim
        assert all(s % self.PATCH == 0 for s in pl.S), "patch cache needs latents of whole 256-token patches"
        n = pl.T // self.PATCH
        self.n_patches = n
        self.patch_latent = torch.from_numpy(
            np.repeat(np.arange(pl.L, dtype=np.int32), [s // self.PATCH for s in pl.S])).to(dev)
        nb = len(model.blocks)
        bf = dict(device=dev, dtype=torch.bfloat16)
        self.xout = [torch.empty((pl.T, D), **bf) for _ in range(nb)]
        self.xprev = [torch.empty((pl.T, D), **bf) for _ in range(nb)]
        self.qkv = [torch.zeros((pl.T, 3 * D), **bf) for _ in range(nb)]  # K / V rows must stay finite
        self.qkv2 = [torch.zeros((pl.T, 3 * D), **bf) if b["dual"] else None for b in model.blocks]
        self.skipped = torch.zeros((nb, n), device=dev, dtype=torch.int32)
        self.mask = torch.ones((nb, n), device=dev, dtype=torch.int32)
        self.mse = torch.zeros((nb, n), device=dev, dtype=torch.float32)
        self.valid = torch.zeros((pl.L,), device=dev, dtype=torch.float32)
        self.ws = ops.patch_mask_workspace(n, dev)
        self.tags = [None] * pl.L     # (request id, CFG branch, step index the kept data is good for)
        self.src = [ops.attn_source(q=q, q_col=0, k=q, k_col=D, v=q, v_col=2 * D, out=pl.att) for q in self.qkv]
        self.src2 = [None if q is None else ops.attn_source(q=q, q_col=0, k=q, k_col=D, v=q, v_col=2 * D, out=pl.att)
                     for q in self.qkv2]

    def bytes(self):
        return sum(t.numel() * t.element_size() for grp in (self.xout, self.xprev, self.qkv, self.qkv2)
                   for t in grp if t is not None)


class B200SD3Transformer2DModel(torch.nn.Module):
    """Built from a diffusers-style SD3Transformer2DModel state dict; weights are re-laid out
    once (fused q|k|v matrices, all AdaLN linears concatenated) and kept in bf16 on the GPU."""

    SUPPORT_RESOLUTIONS = [256, 512, 768, 1024]

    def __init__(self, state_dict: Dict[str, torch.Tensor], config, device="cuda"):
        super().__init__()
        self.cfg = cfg = SD3Config.from_any(config)
        self.config = config
        self.device = torch.device(device)
        self.dtype = torch.bfloat16
        assert cfg.attention_head_dim == 64, "attention kernel is specialised for head_dim 64"
        D = cfg.inner_dim
        sd = state_dict

        def w(name):
            return sd[name].to(device=self.device, dtype=torch.bfloat16).contiguous()

        def cat(names):
            return torch.cat([sd[n] for n in names], 0).to(device=self.device, dtype=torch.bfloat16).contiguous()

        self.pe_w = w("pos_embed.proj.weight").reshape(D, -1).contiguous()
        self.pe_b = w("pos_embed.proj.bias")
        self.pos_table = sd["pos_embed.pos_embed"].reshape(-1, D).to(self.device, torch.bfloat16)
        self.te = {k: w(f"time_text_embed.{k}") for k in (
            "timestep_embedder.linear_1.weight", "timestep_embedder.linear_1.bias",
            "timestep_embedder.linear_2.weight", "timestep_embedder.linear_2.bias",
            "text_embedder.linear_1.weight", "text_embedder.linear_1.bias",
            "text_embedder.linear_2.weight", "text_embedder.linear_2.bias")}
        self.ctx_w, self.ctx_b = w("context_embedder.weight"), w("context_embedder.bias")
        self.blocks = []
        mod_w, mod_b, col = [], [], 0
        for i in range(cfg.num_layers):
            b = f"transformer_blocks.{i}"
            dual = i in cfg.dual_attention_layers
            last = i == cfg.num_layers - 1
            blk = {"dual": dual, "last": last}
            blk["mod"] = col
            mod_w.append(sd[b + ".norm1.linear.weight"]); mod_b.append(sd[b + ".norm1.linear.bias"])
            col += (9 if dual else 6) * D
            blk["cmod"] = col
            mod_w.append(sd[b + ".norm1_context.linear.weight"]); mod_b.append(sd[b + ".norm1_context.linear.bias"])
            col += (2 if last else 6) * D
            blk["qkv_w"] = cat([f"{b}.attn.to_{n}.weight" for n in "qkv"])
            blk["qkv_b"] = cat([f"{b}.attn.to_{n}.bias" for n in "qkv"])
            blk["aqkv_w"] = cat([f"{b}.attn.add_{n}_proj.weight" for n in "qkv"])
            blk["aqkv_b"] = cat([f"{b}.attn.add_{n}_proj.bias" for n in "qkv"])
            for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
                blk[n] = w(f"{b}.attn.{n}.weight")
            blk["out_w"], blk["out_b"] = w(b + ".attn.to_out.0.weight"), w(b + ".attn.to_out.0.bias")
            if not last:
                blk["aout_w"], blk["aout_b"] = w(b + ".attn.to_add_out.weight"), w(b + ".attn.to_add_out.bias")
                blk["ffc1_w"], blk["ffc1_b"] = w(b + ".ff_context.net.0.proj.weight"), w(b + ".ff_context.net.0.proj.bias")
                blk["ffc2_w"], blk["ffc2_b"] = w(b + ".ff_context.net.2.weight"), w(b + ".ff_context.net.2.bias")
            if dual:
                blk["qkv2_w"] = cat([f"{b}.attn2.to_{n}.weight" for n in "qkv"])
                blk["qkv2_b"] = cat([f"{b}.attn2.to_{n}.bias" for n in "qkv"])
                blk["norm_q2"], blk["norm_k2"] = w(b + ".attn2.norm_q.weight"), w(b + ".attn2.norm_k.weight")
                blk["out2_w"], blk["out2_b"] = w(b + ".attn2.to_out.0.weight"), w(b + ".attn2.to_out.0.bias")
            blk["ff1_w"], blk["ff1_b"] = w(b + ".ff.net.0.proj.weight"), w(b + ".ff.net.0.proj.bias")
            blk["ff2_w"], blk["ff2_b"] = w(b + ".ff.net.2.weight"), w(b + ".ff.net.2.bias")
            # q and k are RMS-normalised per head and scaled by learned weights: |q| <= 8 max|w_q|,
            # |k| <= 8 max|w_k| (head_dim 64), so |logit * scale * log2 e| <= 64 max|w_q| max|w_k| / 8 * 1.4427
            # (+1 % for the bf16 rounding of q and k). When that is <= 64 the attention kernel may run its
            # softmax without a reference maximum (B200AttnExtra.bounded_logits); checked per block and per
            # attention from the weights actually loaded, never assumed.
            def _bounded(qs, ks):
                wq = max(float(blk[n].float().abs().max()) for n in qs)
                wk = max(float(blk[n].float().abs().max()) for n in ks)
                return cfg.attention_head_dim * wq * wk / math.sqrt(cfg.attention_head_dim) * 1.4427 * 1.01 <= 64.0
            blk["bounded"] = _bounded(("norm_q", "norm_added_q"), ("norm_k", "norm_added_k"))
            blk["bounded2"] = dual and _bounded(("norm_q2",), ("norm_k2",))
            self.blocks.append(blk)
        self.out_mod = col
        mod_w.append(sd["norm_out.linear.weight"]); mod_b.append(sd["norm_out.linear.bias"])
        col += 2 * D
        self.mod_cols = col
        self.mod_w = torch.cat(mod_w, 0).to(self.device, torch.bfloat16).contiguous()
        self.mod_b = torch.cat(mod_b, 0).to(self.device, torch.bfloat16).contiguous()
        self.proj_w, self.proj_b = w("proj_out.weight"), w("proj_out.bias")
        self.arena = ops.Arena(self.device)
        self._plans = ops.PlanCache(self.device, arena=self.arena)
        self.use_graphs = ops.graphs_enabled()

    @classmethod
    def from_diffusers(cls, model, device="cuda"):
        """`model` is the diffusers SD3Transformer2DModel that sduss' loader hands to
        instantiate_pipeline (pipeline_stable_diffusion_3_esymred.py:24-36)."""
        return cls(model.state_dict(), model.config, device=device)

    # ------------------------------------------------------------------
    def _side_stream(self):
        st = getattr(self, "_side", None)
        if st is None:
            st = self._side = torch.cuda.Stream(device=self.device)
        return st

    def _plan(self, hidden_states, ctx_len) -> _Plan:
        comp = tuple((res, t.shape[0], t.shape[-2], t.shape[-1])
                     for res, t in hidden_states.items() if t is not None and t.shape[0] > 0)
        return self.plan_for(comp, ctx_len)

    def plan_for(self, comp, ctx_len) -> _Plan:
        """comp: ((resolution key, latents, h, w), ...) in ascending resolution order."""
        return self._plans.get((comp, ctx_len), lambda: _Plan(self, comp, ctx_len))

    # ---- patch cache
    def enable_patch_cache(self, forest, refresh: int = 2, max_cached_plans: int = 3):
        """forest: ops.DeviceForest deciding, from [block, timestep, MSE of the block input against the
        previous step], which 256-token patches a block recomputes (reference: the cuML RandomForest
        of ESYMRED_TRANSFORMER_PATH, cache_manager.py:37-44,161-191; refresh = 2 forced recompute
        after two skips, :183). Only `max_cached_plans` batch compositions keep their (multi-GB)
        cache buffers at a time; pass forest=None to switch the cache off."""
        import collections
        self.patch_forest, self.patch_refresh = forest, refresh
        self._cache_lru, self._cache_max = collections.OrderedDict(), max_cached_plans
        for pl in self._plans.values():
            # captured graphs of the cached forward hold the previous forest's device pointers
            pl.cached_state = _CachedState() if (forest is not None and pl.cache is not None) else None
            if forest is None:
                pl.cache = None
            elif pl.cache is not None:
                self._cache_lru[id(pl)] = pl

    def patch_cache_enabled(self):
        return getattr(self, "patch_forest", None) is not None

    def cache_for(self, pl):
        """The plan's cache buffers (allocated on first use; least recently used plans lose theirs)."""
        key = id(pl)
        if pl.cache is None:
            while len(self._cache_lru) >= self._cache_max:
                _, old = self._cache_lru.popitem(last=False)
                old.cached_state = None   # its graph points into the buffers: drop it first
                old.cache = None
            pl.cache, pl.cached_state = _PatchCache(self, pl), _CachedState()
        self._cache_lru[key] = pl
        self._cache_lru.move_to_end(key)
        return pl.cache

    # ---- per-request conditioning, computed once per request (sduss_b200.pipelines caches it)
    cond_kind = "sd3"

    def project_context(self, ehs: torch.Tensor) -> torch.Tensor:
        """context_embedder on raw prompt embeddings [n, ctx, 4096] -> [n, ctx, 1536]
        (SD3Transformer.py:115). It does not depend on the timestep, so a request's projected
        context is computed on first sight and re-used by all its steps."""
        n, ctx, d = ehs.shape
        out = torch.empty((n * ctx, self.cfg.inner_dim), device=self.device, dtype=torch.bfloat16)
        ops.gemm(ehs.reshape(n * ctx, d), self.ctx_w, out, bias=self.ctx_b)
        return out.view(n, ctx, -1)

    @torch.no_grad()
    def forward(self, hidden_states: Dict[str, torch.Tensor], encoder_hidden_states=None,
                pooled_projections=None, timestep=None, block_controlnet_hidden_states=None,
                joint_attention_kwargs=None, return_dict: bool = False, skip_layers=None,
                patch_size: Optional[int] = None, is_sliced: bool = True, save_index: int = 0,
                input_indices: Optional[dict] = None, _borrow: bool = False):
        assert block_controlnet_hidden_states is None and skip_layers is None
        assert not joint_attention_kwargs, "joint_attention_kwargs (LoRA scale / IP-adapter) unsupported"
        plan = self._plan(hidden_states, encoder_hidden_states.shape[1])

        def load_inputs(pl):
            for res, _, _, _ in pl.comp:
                pl.stage_in[res].copy_(hidden_states[res])
            pl.ehs.copy_(encoder_hidden_states.reshape(pl.Tc, -1))
            pl.pooled.copy_(pooled_projections)
            pl.t32.copy_(timestep.reshape(-1))
            ops.gemm(pl.ehs, self.ctx_w, pl.c, bias=self.ctx_b)  # context_embedder

        ops.run_plan(self, plan, load_inputs)
        out = plan.stage_out if _borrow else {k: v.clone() for k, v in plan.stage_out.items()}
        return (out,)

    def _run(self, pl: _Plan):
        cfg = self.cfg
        D, H, p = cfg.inner_dim, cfg.num_attention_heads, cfg.patch_size
        G = _G
        te = self.te
        scale = 1.0 / math.sqrt(cfg.attention_head_dim)
        # conditioning: temb = MLP(sinusoid(t)) + MLP(pooled); all AdaLN vectors in one GEMM
        ops.timestep_embedding(pl.t32, 256, out=pl.tsin)
        G(pl.tsin, te["timestep_embedder.linear_1.weight"], pl.e1, bias=te["timestep_embedder.linear_1.bias"])
        ops.silu(pl.e1, pl.e1)
        G(pl.e1, te["timestep_embedder.linear_2.weight"], pl.e2, bias=te["timestep_embedder.linear_2.bias"])
        G(pl.pooled, te["text_embedder.linear_1.weight"], pl.e3, bias=te["text_embedder.linear_1.bias"])
        ops.silu(pl.e3, pl.e3)
        G(pl.e3, te["text_embedder.linear_2.weight"], pl.temb, bias=te["text_embedder.linear_2.bias"],
          epi=ops.EPI_GATE_RESID, resid=pl.e2)
        ops.silu(pl.temb, pl.e1)
        G(pl.e1, self.mod_w, pl.mod, bias=self.mod_b)
        # streams (pl.c already holds the projected context: forward() / the step's prologue)
        ops.sd3_patchify(pl.in_ptr, pl.desc, pl.L, pl.max_tokens, cfg.in_channels, p, pl.tokens)
        G(pl.tokens, self.pe_w, pl.x, bias=self.pe_b, epi=ops.EPI_GATE_RESID, resid=pl.pos_rows)
        mod = pl.mod
        # Two-stream schedule: the context stream (1998 rows: its GEMMs fill only part of a wave)
        # runs on a side CUDA stream next to the image stream and the two meet at the joint
        # attention of every block. Inside a captured CUDA graph these become parallel branches.
        main = torch.cuda.current_stream()
        # two_streams = False serialises the context branch behind the image branch (used by
        # bench.py's per-kernel timing pass, where concurrent kernels would blur the durations)
        side = self._side_stream() if getattr(self, "two_streams", True) else main
        ev_main, ev_side = torch.cuda.Event(), torch.cuda.Event()
        ev_main.record(main)
        side.wait_event(ev_main)
        for blk in self.blocks:
            m, cm = blk["mod"], blk["cmod"]
            dual, last = blk["dual"], blk["last"]
            with torch.cuda.stream(side):
                if last:  # AdaLayerNormContinuous: (scale, shift)
                    ops.layernorm_mod(pl.c, pl.cn, eps=1e-6, mod=mod, row_group=pl.row_group_ctx,
                                      shift_col=cm + D, scale_col=cm)
                else:
                    ops.layernorm_mod(pl.c, pl.cn, eps=1e-6, mod=mod, row_group=pl.row_group_ctx,
                                      shift_col=cm, scale_col=cm + D)
                G(pl.cn, blk["aqkv_w"], pl.qkv_c, bias=blk["aqkv_b"], epi=ops.EPI_QK_RMSNORM,
                  rms_wq=blk["norm_added_q"], rms_wk=blk["norm_added_k"], rms_q_cols=D, rms_k_cols=D)
                ev_side.record(side)
            ops.layernorm_mod(pl.x, pl.xn, eps=1e-6, mod=mod, row_group=pl.row_group,
                              shift_col=m, scale_col=m + D,
                              y2=pl.xn2 if dual else None, shift2_col=m + 6 * D, scale2_col=m + 7 * D)
            G(pl.xn, blk["qkv_w"], pl.qkv, bias=blk["qkv_b"], epi=ops.EPI_QK_RMSNORM,
              rms_wq=blk["norm_q"], rms_wk=blk["norm_k"], rms_q_cols=D, rms_k_cols=D)
            main.wait_event(ev_side)
            ops.attn_varlen(pl.src_img, pl.src_ctx, *pl.joint_plan, scale, bounded=blk["bounded"])
            ev_main.record(main)
            if not last:
                with torch.cuda.stream(side):
                    side.wait_event(ev_main)
                    G(pl.att_c, blk["aout_w"], pl.c, bias=blk["aout_b"], epi=ops.EPI_GATE_RESID,
                      resid=pl.c, gate=mod[:, cm + 2 * D:cm + 3 * D], row_group=pl.row_group_ctx)
                    ops.layernorm_mod(pl.c, pl.cn, eps=1e-6, mod=mod, row_group=pl.row_group_ctx,
                                      shift_col=cm + 3 * D, scale_col=cm + 4 * D)
                    G(pl.cn, blk["ffc1_w"], pl.ff_c, bias=blk["ffc1_b"], epi=ops.EPI_GELU_TANH)
                    G(pl.ff_c, blk["ffc2_w"], pl.c, bias=blk["ffc2_b"], epi=ops.EPI_GATE_RESID,
                      resid=pl.c, gate=mod[:, cm + 5 * D:cm + 6 * D], row_group=pl.row_group_ctx)
            G(pl.att, blk["out_w"], pl.x, bias=blk["out_b"], epi=ops.EPI_GATE_RESID, resid=pl.x,
              gate=mod[:, m + 2 * D:m + 3 * D], row_group=pl.row_group)
            if dual:
                G(pl.xn2, blk["qkv2_w"], pl.qkv, bias=blk["qkv2_b"], epi=ops.EPI_QK_RMSNORM,
                  rms_wq=blk["norm_q2"], rms_wk=blk["norm_k2"], rms_q_cols=D, rms_k_cols=D)
                ops.attn_varlen(pl.src_img, None, *pl.self_plan, scale, bounded=blk["bounded2"])
                G(pl.att, blk["out2_w"], pl.x, bias=blk["out2_b"], epi=ops.EPI_GATE_RESID,
                  resid=pl.x, gate=mod[:, m + 8 * D:m + 9 * D], row_group=pl.row_group)
            ops.layernorm_mod(pl.x, pl.xn, eps=1e-6, mod=mod, row_group=pl.row_group,
                              shift_col=m + 3 * D, scale_col=m + 4 * D)
            G(pl.xn, blk["ff1_w"], pl.ff, bias=blk["ff1_b"], epi=ops.EPI_GELU_TANH)
            G(pl.ff, blk["ff2_w"], pl.x, bias=blk["ff2_b"], epi=ops.EPI_GATE_RESID, resid=pl.x,
              gate=mod[:, m + 5 * D:m + 6 * D], row_group=pl.row_group)
        ev_side.record(side)
        main.wait_event(ev_side)  # rejoin (the last block has no context output, but keep the graph closed)
        om = self.out_mod
        ops.layernorm_mod(pl.x, pl.xn, eps=1e-6, mod=mod, row_group=pl.row_group,
                          shift_col=om + D, scale_col=om)
        G(pl.xn, self.proj_w, pl.out_tok, bias=self.proj_b)
        ops.sd3_unpatchify(pl.out_tok, pl.desc, pl.L, pl.max_tokens, cfg.out_channels, p, pl.out_ptr)

    def _run_cached(self, pl: _Plan):
        """The forward with the patch cache (policy in oracle/patch_cache.py).
        Same launches as _run plus one decision kernel per block; every image-stream kernel takes the
        block's patch mask and skips the tiles of clean patches, whose block output / keys / values
        stay in the per-block buffers of pl.cache. The residual stream therefore moves through
        per-block buffers (block i reads xout[i-1], writes xout[i]) instead of being updated in
        place. The context stream (333 tokens per latent) is always recomputed."""
        cfg, cb = self.cfg, pl.cache
        D, H, p = cfg.inner_dim, cfg.num_attention_heads, cfg.patch_size
        G = _G
        te = self.te
        scale = 1.0 / math.sqrt(cfg.attention_head_dim)
        ops.timestep_embedding(pl.t32, 256, out=pl.tsin)
        G(pl.tsin, te["timestep_embedder.linear_1.weight"], pl.e1, bias=te["timestep_embedder.linear_1.bias"])
        ops.silu(pl.e1, pl.e1)
        G(pl.e1, te["timestep_embedder.linear_2.weight"], pl.e2, bias=te["timestep_embedder.linear_2.bias"])
        G(pl.pooled, te["text_embedder.linear_1.weight"], pl.e3, bias=te["text_embedder.linear_1.bias"])
        ops.silu(pl.e3, pl.e3)
        G(pl.e3, te["text_embedder.linear_2.weight"], pl.temb, bias=te["text_embedder.linear_2.bias"],
          epi=ops.EPI_GATE_RESID, resid=pl.e2)
        ops.silu(pl.temb, pl.e1)
        G(pl.e1, self.mod_w, pl.mod, bias=self.mod_b)
        ops.sd3_patchify(pl.in_ptr, pl.desc, pl.L, pl.max_tokens, cfg.in_channels, p, pl.tokens)
        G(pl.tokens, self.pe_w, pl.x, bias=self.pe_b, epi=ops.EPI_GATE_RESID, resid=pl.pos_rows)
        mod = pl.mod
        main = torch.cuda.current_stream()
        side = self._side_stream() if getattr(self, "two_streams", True) else main
        ev_main, ev_side = torch.cuda.Event(), torch.cuda.Event()
        ev_main.record(main)
        side.wait_event(ev_main)
        xin = pl.x
        for i, blk in enumerate(self.blocks):
            m, cm = blk["mod"], blk["cmod"]
            dual, last = blk["dual"], blk["last"]
            mk = cb.mask[i]
            with torch.cuda.stream(side):
                if last:
                    ops.layernorm_mod(pl.c, pl.cn, eps=1e-6, mod=mod, row_group=pl.row_group_ctx,
                                      shift_col=cm + D, scale_col=cm)
                else:
                    ops.layernorm_mod(pl.c, pl.cn, eps=1e-6, mod=mod, row_group=pl.row_group_ctx,
                                      shift_col=cm, scale_col=cm + D)
                G(pl.cn, blk["aqkv_w"], pl.qkv_c, bias=blk["aqkv_b"], epi=ops.EPI_QK_RMSNORM,
                  rms_wq=blk["norm_added_q"], rms_wk=blk["norm_added_k"], rms_q_cols=D, rms_k_cols=D)
                ev_side.record(side)
            # the decision: MSE of the block input against the previous step, forest, refresh rule
            ops.patch_mask(xin, cb.xprev[i], cb.patch_latent, pl.t32, cb.valid, cb.skipped[i], mk,
                           self.patch_forest, i, self.patch_refresh, cb.ws, mse=cb.mse[i])
            ops.layernorm_mod(xin, pl.xn, eps=1e-6, mod=mod, row_group=pl.row_group,
                              shift_col=m, scale_col=m + D,
                              y2=pl.xn2 if dual else None, shift2_col=m + 6 * D, scale2_col=m + 7 * D,
                              row_mask=mk)
            G(pl.xn, blk["qkv_w"], cb.qkv[i], bias=blk["qkv_b"], epi=ops.EPI_QK_RMSNORM,
              rms_wq=blk["norm_q"], rms_wk=blk["norm_k"], rms_q_cols=D, rms_k_cols=D, row_mask=mk)
            main.wait_event(ev_side)
            ops.attn_varlen(cb.src[i], pl.src_ctx, *pl.joint_plan, scale, q_mask=mk, bounded=blk["bounded"])
            ev_main.record(main)
            if not last:
                with torch.cuda.stream(side):
                    side.wait_event(ev_main)
                    G(pl.att_c, blk["aout_w"], pl.c, bias=blk["aout_b"], epi=ops.EPI_GATE_RESID,
                      resid=pl.c, gate=mod[:, cm + 2 * D:cm + 3 * D], row_group=pl.row_group_ctx)
                    ops.layernorm_mod(pl.c, pl.cn, eps=1e-6, mod=mod, row_group=pl.row_group_ctx,
                                      shift_col=cm + 3 * D, scale_col=cm + 4 * D)
                    G(pl.cn, blk["ffc1_w"], pl.ff_c, bias=blk["ffc1_b"], epi=ops.EPI_GELU_TANH)
                    G(pl.ff_c, blk["ffc2_w"], pl.c, bias=blk["ffc2_b"], epi=ops.EPI_GATE_RESID,
                      resid=pl.c, gate=mod[:, cm + 5 * D:cm + 6 * D], row_group=pl.row_group_ctx)
            G(pl.att, blk["out_w"], pl.xa, bias=blk["out_b"], epi=ops.EPI_GATE_RESID, resid=xin,
              gate=mod[:, m + 2 * D:m + 3 * D], row_group=pl.row_group, row_mask=mk)
            if dual:
                G(pl.xn2, blk["qkv2_w"], cb.qkv2[i], bias=blk["qkv2_b"], epi=ops.EPI_QK_RMSNORM,
                  rms_wq=blk["norm_q2"], rms_wk=blk["norm_k2"], rms_q_cols=D, rms_k_cols=D, row_mask=mk)
                ops.attn_varlen(cb.src2[i], None, *pl.self_plan, scale, q_mask=mk, bounded=blk["bounded2"])
                G(pl.att, blk["out2_w"], pl.xa, bias=blk["out2_b"], epi=ops.EPI_GATE_RESID,
                  resid=pl.xa, gate=mod[:, m + 8 * D:m + 9 * D], row_group=pl.row_group, row_mask=mk)
            ops.layernorm_mod(pl.xa, pl.xn, eps=1e-6, mod=mod, row_group=pl.row_group,
                              shift_col=m + 3 * D, scale_col=m + 4 * D, row_mask=mk)
            G(pl.xn, blk["ff1_w"], pl.ff, bias=blk["ff1_b"], epi=ops.EPI_GELU_TANH, row_mask=mk)
            G(pl.ff, blk["ff2_w"], cb.xout[i], bias=blk["ff2_b"], epi=ops.EPI_GATE_RESID, resid=pl.xa,
              gate=mod[:, m + 5 * D:m + 6 * D], row_group=pl.row_group, row_mask=mk)
            xin = cb.xout[i]
        ev_side.record(side)
        main.wait_event(ev_side)
        om = self.out_mod
        ops.layernorm_mod(xin, pl.xn, eps=1e-6, mod=mod, row_group=pl.row_group,
                          shift_col=om + D, scale_col=om)
        G(pl.xn, self.proj_w, pl.out_tok, bias=self.proj_b)
        ops.sd3_unpatchify(pl.out_tok, pl.desc, pl.L, pl.max_tokens, cfg.out_channels, p, pl.out_ptr)
