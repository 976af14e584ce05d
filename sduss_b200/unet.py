"""B200-native drop-in for sduss' PatchUNet (SDXL-base UNet)
(reference: sduss/model_executor/modules/unet.py:27-536, modules/resnet.py:380-460,
modules/transformer.py:25-290, modules/attention.py:52-232, modules/unet_2d_blocks.py).

Same forward contract (dict resolution -> [n_r, 4, h, w] in / tuple(dict) out, conditioning
rows ordered resolution by resolution) but instead of cutting every image into 256-px patches
with halos, each UNet level keeps ONE packed channels-last buffer [sum_i h_i*w_i, C] bf16:
  * 3x3 convs (stride 1 / 2) are implicit GEMMs on tcgen05 whose TMA box loads zero-fill the
    image border (no halo exchange, exact corners -- reference deviation D2 does not exist here);
  * GroupNorm uses exact whole-latent statistics (reference deviation D1 does not exist here);
  * the pixel rows ARE the attention tokens: self / cross attention run as one packed varlen
    launch per layer; text K/V of all 70 cross-attention layers come from ONE GEMM per step on
    [L*77, 2048] (the reference recomputes them per patch);
  * time-embedding projections of all resnets come from ONE GEMM per step.
"""
import os
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import ops
from .layout import LevelLayout


def _G(*a, **k):
    """ops.gemm against WEIGHTS (never written on the stream): W tiles may be fetched ahead of the wait
    on the previous kernel."""
    return ops.gemm(*a, w_static=True, **k)


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    down_has_attn: Tuple[bool, ...] = (False, True, True)
    num_heads: Tuple[int, ...] = (5, 10, 20)
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    pooled_dim: int = 1280
    norm_num_groups: int = 32
    norm_eps: float = 1e-5

    @classmethod
    def from_any(cls, cfg):
        if isinstance(cfg, cls):
            return cfg
        get = cfg.get if hasattr(cfg, "get") else lambda k, d=None: getattr(cfg, k, d)
        kw = {}
        for k in ("in_channels", "out_channels", "layers_per_block", "cross_attention_dim",
                  "addition_time_embed_dim", "norm_num_groups", "norm_eps", "pooled_dim"):
            if get(k, None) is not None:
                kw[k] = get(k)
        if get("block_out_channels", None) is not None:
            kw["block_out_channels"] = tuple(get("block_out_channels"))
        if get("transformer_layers_per_block", None) is not None:
            t = get("transformer_layers_per_block")
            kw["transformer_layers_per_block"] = tuple(t) if not isinstance(t, int) else (t,) * 3
        for src, dst in (("attention_head_dim", "num_heads"), ("num_heads", "num_heads")):
            if get(src, None) is not None and not isinstance(get(src), int):
                kw[dst] = tuple(get(src))
        if get("down_has_attn", None) is not None:
            kw["down_has_attn"] = tuple(get("down_has_attn"))
        elif get("down_block_types", None) is not None:
            kw["down_has_attn"] = tuple("CrossAttn" in t for t in get("down_block_types"))
        if get("projection_class_embeddings_input_dim", None) is not None and "pooled_dim" not in kw:
            kw["pooled_dim"] = get("projection_class_embeddings_input_dim") - 6 * kw.get(
                "addition_time_embed_dim", 256)
        return cls(**kw)


class _Plan:
    def __init__(self, model: "B200UNet", comp, ctx_len):
        dev = model.device
        cfg = model.cfg
        self.comp, self.ctx_len = comp, ctx_len
        sizes0 = [(h, w) for _, n, h, w in comp for _ in range(n)]
        self.L = L = len(sizes0)
        nlev = len(cfg.block_out_channels)
        self.levels = [LevelLayout([(h >> l, w >> l) for h, w in sizes0], dev) for l in range(nlev)]
        bf = dict(device=dev, dtype=torch.bfloat16)
        numel_in = [n * cfg.in_channels * h * w for _, n, h, w in comp]
        numel_out = [n * cfg.out_channels * h * w for _, n, h, w in comp]
        self.flat_in = torch.empty((sum(numel_in),), **bf)
        self.flat_out = torch.empty((sum(numel_out),), **bf)
        self.stage_in, self.stage_out, self.in_elem_off, self.out_elem_off = {}, {}, {}, {}
        oi = oo = 0
        for (res, n, h, w), ni, no in zip(comp, numel_in, numel_out):
            self.stage_in[res] = self.flat_in[oi:oi + ni].view(n, cfg.in_channels, h, w)
            self.stage_out[res] = self.flat_out[oo:oo + no].view(n, cfg.out_channels, h, w)
            self.in_elem_off[res], self.out_elem_off[res] = oi, oo
            oi, oo = oi + ni, oo + no
        ip, op = [], []
        for res, n, _, _ in comp:
            for i in range(n):
                ip.append(self.stage_in[res][i].data_ptr())
                op.append(self.stage_out[res][i].data_ptr())
        self.in_ptr = torch.tensor(ip, dtype=torch.int64).to(dev)
        self.out_ptr = torch.tensor(op, dtype=torch.int64).to(dev)
        # attention plans per level
        self.self_plan, self.cross_plan = {}, {}
        for l, lay in enumerate(self.levels):
            selfp = [(lay.row_off[i], lay.rows[i], 0, 0, lay.row_off[i], lay.rows[i], 0, 0) for i in range(L)]
            crossp = [(lay.row_off[i], lay.rows[i], 0, 0, 0, 0, i * ctx_len, ctx_len) for i in range(L)]
            self.self_plan[l] = ops.build_attn_plan(selfp, dev, cfg.num_heads[l])
            self.cross_plan[l] = ops.build_attn_plan(crossp, dev, cfg.num_heads[l])
        # 77 text tokens: the short-key cross-attention kernel (no work list: it reads the plan's seq_table)
        self.cross_short = ctx_len <= ops.ATTN_CROSS_SHORT_KEYS and os.environ.get("SDUSS_B200_NO_XSHORT", "0") != "1"
        self.max_rows = {l: max(lay.rows) for l, lay in enumerate(self.levels)}
        self.gn_ws = ops.groupnorm_workspace(self.levels[0].T, L, dev)
        self.arena, self.block, self.arena_off = model.arena, None, 0
        self.bufs: Dict[str, torch.Tensor] = {}
        self.maps: Dict[tuple, torch.Tensor] = {}
        self.stats_tag, self.stats_gen = {}, {}   # conv-epilogue GroupNorm statistics (ops.conv_stats_buffer)
        self.attn_src: Dict[tuple, object] = {}
        self.t32 = torch.empty((L,), device=dev, dtype=torch.float32)
        self.ids32 = torch.empty((L * 6,), device=dev, dtype=torch.float32)
        self.ehs = torch.empty((L * ctx_len, cfg.cross_attention_dim), **bf)
        self.text_embeds = torch.empty((L, cfg.pooled_dim), **bf)
        self.device = dev
        self.graph = None
        self.graph_launches = 0
        self.cache = None          # _UNetPatchCache, allocated when the plan first runs with the patch cache
        self.cached_state = None   # graph / use counters of the cached way of running this plan

    def buf(self, name, rows, cols):
        t = self.bufs.get(name)
        if t is None:  # bump-allocated from the model's shared arena (ops.Arena.alloc)
            t = self.bufs[name] = self.arena.alloc(self, (rows, cols), torch.bfloat16)
        assert t.shape == (rows, cols), (name, t.shape, rows, cols)
        return t

    def fbuf(self, name, numel):
        t = self.bufs.get(name)
        if t is None:
            t = self.bufs[name] = self.arena.alloc(self, (numel,), torch.float32)
        return t

    def reset_workspaces(self):
        """Forget every arena view and everything that captured its address (ops._run_eager)."""
        self.bufs, self.maps, self.block, self.arena_off = {}, {}, None, 0
        self.stats_tag, self.stats_gen = {}, {}
        if hasattr(self, "attn_src"):
            self.attn_src = {}
        if getattr(self, "cache", None) is not None:
            self.cache.src = {}   # attention sources of the cached forward point at arena views too

    def conv_maps(self, x, cin, level, stride):
        key = (x.data_ptr(), x.stride(0), cin, level, stride)
        m = self.maps.get(key)
        if m is None:
            m = self.maps[key] = ops.conv3x3_encode_maps(x, cin, self.levels[level].desc_host, stride)
        return m


class _CachedState:
    graph = None
    graph_launches = 0
    uses = 0
    warm = False


class _UNetPatchCache:
    """What a plan keeps ACROSS steps when the patch cache is on (SDXL variant;
    reference: CacheManager.get_mask, cache_manager.py:101-159, called once per down / mid / up block,
    modules/unet_2d_blocks.py:40,102,180,250,345). Per UNet block: the block input of the previous step
    (and, for an up block, of every skip tensor it consumes) for the MSE features, the skip counters and
    the mask of its 256-row patches. Per convolution of the block (conv1, conv2, shortcut, sampler) its
    output, per Transformer2D its residual stream, its output and the fused q|k|v of every layer -- what
    lies in clean patches is not rewritten, so it holds what the patch had when it was last computed. These buffers are owned here, not carved
    from the model's arena (the arena is shared by all plans and overwritten by whichever ran last).
    Validity is per latent and by plan slot, exactly as for the SD3.5 cache (sd3_transformer._PatchCache)."""

    PATCH = 256

    def __init__(self, model, pl):
        self.device = model.device
        self.L = pl.L
        self.bufs: Dict[str, torch.Tensor] = {}
        self.blocks: Dict[str, object] = {}
        self.level_tabs: Dict[int, tuple] = {}
        self.src: Dict[tuple, object] = {}
        self.valid = torch.zeros((pl.L,), device=self.device, dtype=torch.float32)
        self.tags = [None] * pl.L     # (request id, CFG branch, step index the kept data is good for)

    def buf(self, name, rows, cols, dtype=torch.bfloat16):
        t = self.bufs.get(name)
        if t is None:   # zeros: kept K / V rows must be finite before their first computation
            t = self.bufs[name] = torch.zeros((rows, cols), device=self.device, dtype=dtype)
        assert t.shape == (rows, cols), (name, t.shape, rows, cols)
        return t

    def block(self, key, lay, level):
        """Decision state of one UNet block, or None when a latent of this level is not made of whole
        256-row patches (768^2 at the deepest level: 24 x 24 pixels) -- that block then runs uncached."""
        if key in self.blocks:
            return self.blocks[key]
        st = None
        if all(r % self.PATCH == 0 for r in lay.rows):
            n = lay.T // self.PATCH
            if level not in self.level_tabs:
                pat = np.repeat(np.arange(lay.L, dtype=np.int32), [r // self.PATCH for r in lay.rows])
                self.level_tabs[level] = (torch.from_numpy(pat).to(self.device),
                                          ops.patch_mask_workspace(n, self.device))
            st = SimpleNamespace()
            st.n = n
            st.patch_latent, st.ws = self.level_tabs[level]
            st.skipped = torch.zeros((n,), device=self.device, dtype=torch.int32)
            st.mask = torch.ones((n,), device=self.device, dtype=torch.int32)
            st.mse = torch.zeros((n,), device=self.device, dtype=torch.float32)
        self.blocks[key] = st
        return st

    def masks(self):
        """{block key: int32 mask tensor} of the blocks that decide (diagnostics / tests)."""
        return {k: st.mask for k, st in self.blocks.items() if st is not None}

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())



class B200UNet(torch.nn.Module):
    SUPPORT_RESOLUTIONS = [256, 512, 768, 1024]

    def __init__(self, state_dict: Dict[str, torch.Tensor], config, device="cuda"):
        super().__init__()
        self.cfg = cfg = UNetConfig.from_any(config)
        self.config = config
        self.device = torch.device(device)
        self.dtype = torch.bfloat16
        sd = state_dict
        dev = self.device
        self.w: Dict[str, torch.Tensor] = {}

        def put(name, t):
            self.w[name] = t.to(device=dev, dtype=torch.bfloat16).contiguous()

        ch = cfg.block_out_channels
        # conv_in as a GEMM on im2col rows: [Cout, Cin*9] padded to K=64
        k_in = cfg.in_channels * 9
        self.k_in_pad = ((k_in + 63) // 64) * 64
        wi = torch.zeros(ch[0], self.k_in_pad)
        wi[:, :k_in] = sd["conv_in.weight"].reshape(ch[0], k_in).float()
        put("conv_in.weight", wi)
        put("conv_in.bias", sd["conv_in.bias"])
        for n in ("time_embedding.linear_1", "time_embedding.linear_2", "add_embedding.linear_1",
                  "add_embedding.linear_2"):
            put(n + ".weight", sd[n + ".weight"])
            put(n + ".bias", sd[n + ".bias"])
        # conv_out padded to 8 output channels
        self.n_out_pad = 8
        wo = torch.zeros(self.n_out_pad, ch[0], 3, 3)
        wo[:cfg.out_channels] = sd["conv_out.weight"].float()
        put("conv_out.weight", wo.permute(0, 2, 3, 1).reshape(self.n_out_pad, 9 * ch[0]))
        bo = torch.zeros(self.n_out_pad)
        bo[:cfg.out_channels] = sd["conv_out.bias"].float()
        put("conv_out.bias", bo)
        put("conv_norm_out.weight", sd["conv_norm_out.weight"])
        put("conv_norm_out.bias", sd["conv_norm_out.bias"])

        temb_w, temb_b, self.temb_off, tcol = [], [], {}, 0
        kv_w, self.kv_off, kcol = [], {}, 0

        def resnet(name):
            nonlocal tcol
            for n in ("norm1", "norm2"):
                put(f"{name}.{n}.weight", sd[f"{name}.{n}.weight"])
                put(f"{name}.{n}.bias", sd[f"{name}.{n}.bias"])
            for n in ("conv1", "conv2"):
                wt = sd[f"{name}.{n}.weight"]
                put(f"{name}.{n}.weight", wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1))
                put(f"{name}.{n}.bias", sd[f"{name}.{n}.bias"])
            if f"{name}.conv_shortcut.weight" in sd:
                wt = sd[f"{name}.conv_shortcut.weight"]
                put(f"{name}.conv_shortcut.weight", wt.reshape(wt.shape[0], wt.shape[1]))
                put(f"{name}.conv_shortcut.bias", sd[f"{name}.conv_shortcut.bias"])
            tw = sd[f"{name}.time_emb_proj.weight"]
            temb_w.append(tw)
            temb_b.append(sd[f"{name}.time_emb_proj.bias"])
            self.temb_off[name] = (tcol, tw.shape[0])
            tcol += tw.shape[0]

        def transformer(name, layers):
            nonlocal kcol
            for n in ("norm.weight", "norm.bias", "proj_in.weight", "proj_in.bias", "proj_out.weight",
                      "proj_out.bias"):
                put(f"{name}.{n}", sd[f"{name}.{n}"])
            for j in range(layers):
                b = f"{name}.transformer_blocks.{j}"
                for n in ("norm1", "norm2", "norm3"):
                    put(f"{b}.{n}.weight", sd[f"{b}.{n}.weight"])
                    put(f"{b}.{n}.bias", sd[f"{b}.{n}.bias"])
                put(b + ".attn1.qkv.weight", torch.cat([sd[f"{b}.attn1.to_{n}.weight"] for n in "qkv"], 0))
                put(b + ".attn1.out.weight", sd[b + ".attn1.to_out.0.weight"])
                put(b + ".attn1.out.bias", sd[b + ".attn1.to_out.0.bias"])
                put(b + ".attn2.q.weight", sd[b + ".attn2.to_q.weight"])
                put(b + ".attn2.out.weight", sd[b + ".attn2.to_out.0.weight"])
                put(b + ".attn2.out.bias", sd[b + ".attn2.to_out.0.bias"])
                kvw = torch.cat([sd[b + ".attn2.to_k.weight"], sd[b + ".attn2.to_v.weight"]], 0)
                kv_w.append(kvw)
                self.kv_off[b] = kcol
                kcol += kvw.shape[0]
                # GEGLU: interleave [32 hidden | 32 gate] row blocks so both halves share a tile
                w1, b1 = sd[b + ".ff.net.0.proj.weight"], sd[b + ".ff.net.0.proj.bias"]
                Fh = w1.shape[0] // 2
                idx = torch.arange(2 * Fh).view(2, Fh // 32, 32).permute(1, 0, 2).reshape(-1)
                put(b + ".ff1.weight", w1[idx])
                put(b + ".ff1.bias", b1[idx])
                # The three LayerNorms of the block folded into the GEMMs that consume them
                # (ops.fold_layernorm): W' = W o gamma, colsum(W'), bias' = bias + beta W^T
                for ln, dst, wt, bias in ((".norm1", ".attn1.qkv_f", torch.cat([sd[f"{b}.attn1.to_{n}.weight"] for n in "qkv"], 0), None),
                                          (".norm2", ".attn2.q_f", sd[b + ".attn2.to_q.weight"], None),
                                          (".norm3", ".ff1_f", w1[idx], b1[idx])):
                    wq, cs, bq = ops.fold_layernorm(wt.to(dev), sd[b + ln + ".weight"].to(dev), sd[b + ln + ".bias"].to(dev),
                                                    None if bias is None else bias.to(dev))
                    self.w[b + dst + ".weight"], self.w[b + dst + ".colsum"], self.w[b + dst + ".bias"] = wq, cs, bq
                put(b + ".ff2.weight", sd[b + ".ff.net.2.weight"])
                put(b + ".ff2.bias", sd[b + ".ff.net.2.bias"])

        for i in range(len(ch)):
            for j in range(cfg.layers_per_block):
                resnet(f"down_blocks.{i}.resnets.{j}")
                if cfg.down_has_attn[i]:
                    transformer(f"down_blocks.{i}.attentions.{j}", cfg.transformer_layers_per_block[i])
            if i != len(ch) - 1:
                n = f"down_blocks.{i}.downsamplers.0.conv"
                wt = sd[n + ".weight"]
                put(n + ".weight", wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1))
                put(n + ".bias", sd[n + ".bias"])
        resnet("mid_block.resnets.0")
        transformer("mid_block.attentions.0", cfg.transformer_layers_per_block[-1])
        resnet("mid_block.resnets.1")
        rev_layers = list(reversed(cfg.transformer_layers_per_block))
        rev_attn = list(reversed(cfg.down_has_attn))
        for i in range(len(ch)):
            for j in range(cfg.layers_per_block + 1):
                resnet(f"up_blocks.{i}.resnets.{j}")
                if rev_attn[i]:
                    transformer(f"up_blocks.{i}.attentions.{j}", rev_layers[i])
            if i != len(ch) - 1:
                n = f"up_blocks.{i}.upsamplers.0.conv"
                wt = sd[n + ".weight"]
                put(n + ".weight", wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1))
                put(n + ".bias", sd[n + ".bias"])
        put("temb_all.weight", torch.cat(temb_w, 0))
        put("temb_all.bias", torch.cat(temb_b, 0))
        self.temb_cols = tcol
        put("kv_all.weight", torch.cat(kv_w, 0))
        self.kv_cols = kcol
        self.arena = ops.Arena(self.device)  # per-step workspaces of all plans overlap here
        self._plans = ops.PlanCache(self.device, arena=self.arena)
        self.use_graphs = ops.graphs_enabled()
        # GroupNorm statistics from the producing convolution's epilogue: built, parity-tested, OFF
        # (SDXL step: convolutions +0.26 ms, GroupNorms -0.11 ms)
        self.fuse_gn_stats = False
        # The 210 LayerNorms of the 70 BasicTransformerBlocks folded into the GEMMs that consume them
        # (row statistics + an epilogue correction instead of a read + write of the activation)
        self.fold_ln = os.environ.get("SDUSS_B200_NO_LN_FOLD", "0") != "1"

    @classmethod
    def from_diffusers(cls, model, device="cuda"):
        """`model`: the diffusers UNet2DConditionModel sduss hands to instantiate_pipeline
        (pipeline_stable_diffusion_xl_esymred.py:30-41)."""
        return cls(model.state_dict(), model.config, device=device)

    @property
    def add_embedding(self):  # attribute the SDXL pipeline reads (base_module.py:24-26)
        return self

    # ------------------------------------------------------------------ building blocks
    def _gn(self, pl, x, name, level, out, silu, eps=None):
        lay = pl.levels[level]
        kw = dict(groups=self.cfg.norm_num_groups, eps=self.cfg.norm_eps if eps is None else eps, silu=silu)
        stats = ops.fresh_conv_stats(pl, x)
        if stats is not None:  # the convolution that wrote x left its statistics: x is read once
            return ops.groupnorm_from_conv_stats(x, out, self.w[name + ".weight"], self.w[name + ".bias"],
                                                 lay.row_group, stats, lay.lat_tiles, lay.L, pl.gn_ws, **kw)
        ops.groupnorm_nhwc(x, out, self.w[name + ".weight"], self.w[name + ".bias"], lay.row_group,
                           lay.lat_chunks, lay.L, pl.gn_ws, **kw)
        return out

    def _conv(self, pl, x, cin, name, in_level, stride, out, resid=None, mask=None, mask_scale=0, **epi):
        out_level = in_level + (1 if stride == 2 else 0)
        lay = pl.levels[out_level]
        w = self.w[name + ".weight"]
        cout = w.shape[0]
        maps = pl.conv_maps(x, cin, in_level, stride)
        out_maps = pl.conv_maps(out, cout, out_level, 1)
        resid_maps = pl.conv_maps(resid, cout, out_level, 1) if resid is not None else None
        # statistics of the output for the GroupNorm that usually reads it next (resnet norm2, the
        # next resnet's norm1, a Transformer2D norm, conv_norm_out)
        st = ops.conv_stats_buffer(pl, out, out_level, lay.n_tiles, cout) \
            if (self.fuse_gn_stats and cout % self.cfg.norm_num_groups == 0) else None
        if mask is not None:
            st = None   # (a skipped pixel block would leave no statistics)
            epi = dict(epi, row_mask=mask, row_mask_scale=mask_scale)
        return ops.conv3x3(maps, lay.tiles, lay.n_tiles, lay.desc, cin, cout, stride, w, out,
                           out_maps=out_maps, resid_maps=resid_maps, bias=self.w[name + ".bias"],
                           stats_out=st, **epi)

    def _resnet(self, pl, x, name, level, temb_all, cb=None, mask=None):
        """cb / mask (patch cache, see _run): both GroupNorms run over the whole tensor; conv1, conv2 and
        the 1x1 shortcut skip what lies in clean patches and their outputs live in cb's own buffers
        (reference: modules/resnet.py:399-458, conv1_output / conv2_output caches)."""
        lay = pl.levels[level]
        T, cin = x.shape
        cout = self.w[name + ".conv1.weight"].shape[0]
        pbuf = cb.buf if cb is not None else pl.buf
        h = self._gn(pl, x, name + ".norm1", level, pl.buf(f"gn{level}_{cin}", T, cin), True)
        off, n = self.temb_off[name]
        h1 = self._conv(pl, h, cin, name + ".conv1", level, 1, pbuf(name + ".h1", T, cout), mask=mask,
                        epi=ops.EPI_ROWVEC, rowvec=temb_all[:, off:off + n], row_group=lay.row_group)
        h2 = self._gn(pl, h1, name + ".norm2", level, pl.buf(f"gn{level}_{cout}", T, cout), True)
        if name + ".conv_shortcut.weight" in self.w:
            s = ops.gemm(x, self.w[name + ".conv_shortcut.weight"], pbuf(name + ".sc", T, cout),
                         bias=self.w[name + ".conv_shortcut.bias"], row_mask=mask)
        else:
            s = x
        return self._conv(pl, h2, cout, name + ".conv2", level, 1, pbuf(name + ".out", T, cout), mask=mask,
                          epi=ops.EPI_GATE_RESID, resid=s)

    def _transformer(self, pl, x, name, level, heads, layers, kv_all, cb=None, mask=None):
        """cb / mask (patch cache, see _run): every token-wise kernel skips the 256-row tiles of clean
        patches; the residual stream, the output and the per-layer q|k|v live in cb's own buffers, so the
        skipped rows still hold the values of their last computation (clean patches keep their keys /
        values for the flagged patches' queries and their Transformer2D output for what follows)."""
        w = self.w
        T, C = x.shape
        rm = dict(row_mask=mask) if mask is not None else {}
        pbuf = (lambda nm, r, c: cb.buf(nm, r, c)) if cb is not None else pl.buf

        def G(*a, **k):
            return _G(*a, **rm, **k)
        fold = self.fold_ln and C % 64 == 0
        # Folded LayerNorms: every GEMM that writes h also leaves per-64-column partial sums of its
        # output rows (rowpart_out), and the GEMM that consumes LN(h) reduces them to (mean, rstd) at
        # the start of each tile and corrects its accumulators (ln_rowpart + ln_colsum): the 3
        # LayerNorms of a block cost no launch and no pass over h.
        rp = pl.fbuf(f"rowpart{level}", T * (C // 64) * 2) if fold else None
        n = self._gn(pl, x, name + ".norm", level, pl.buf(f"gn{level}_{C}", T, C), False, eps=1e-6)
        h = G(n, w[name + ".proj_in.weight"], pbuf(name + ".h", T, C), bias=w[name + ".proj_in.bias"],
              rowpart_out=rp)
        ln = pl.buf(f"ln{level}_{C}", T, C)
        qkv = pl.buf(f"qkv{level}_{C}", T, 3 * C)
        att = pl.buf(f"att{level}_{C}", T, C)
        q2 = pl.buf(f"q2_{level}_{C}", T, C)
        ff = pl.buf(f"ff{level}_{C}", T, 4 * C)
        key = (level, C, "self")
        if key not in pl.attn_src:
            pl.attn_src[key] = ops.attn_source(q=qkv, q_col=0, k=qkv, k_col=C, v=qkv, v_col=2 * C, out=att)
            pl.attn_src[(level, C, "q")] = ops.attn_source(q=q2, out=att)
        def ln_gemm(norm, wname, out, **kw):
            """out = epilogue(LayerNorm(h) W^T + b)."""
            if fold:
                return G(h, w[b + wname + "_f.weight"], out, bias=w[b + wname + "_f.bias"], ln_rowpart=rp,
                         ln_colsum=w[b + wname + "_f.colsum"], ln_eps=1e-5, **kw)
            ops.layernorm_mod(h, ln, eps=1e-5, gamma=w[b + norm + ".weight"], beta=w[b + norm + ".bias"])
            return G(ln, w[b + wname + ".weight"], out, bias=w.get(b + wname + ".bias"), **kw)

        for j in range(layers):
            b = f"{name}.transformer_blocks.{j}"
            src_self = pl.attn_src[key]
            if cb is not None:   # this layer's own q|k|v: the keys / values of clean patches persist
                qkv = cb.buf(b + ".qkv", T, 3 * C)
                src_self = cb.src.get(b)
                if src_self is None:
                    src_self = cb.src[b] = ops.attn_source(q=qkv, q_col=0, k=qkv, k_col=C, v=qkv, v_col=2 * C, out=att)
            ln_gemm(".norm1", ".attn1.qkv", qkv)
            ops.attn_varlen(src_self, None, *pl.self_plan[level], 0.125, q_mask=mask)
            G(att, w[b + ".attn1.out.weight"], h, bias=w[b + ".attn1.out.bias"], epi=ops.EPI_GATE_RESID, resid=h,
              rowpart_out=rp)
            ln_gemm(".norm2", ".attn2.q", q2)
            ko = self.kv_off[b]
            src_kv = pl.attn_src.get((b, "kv"))
            if src_kv is None:
                src_kv = pl.attn_src[(b, "kv")] = ops.attn_source(k=kv_all, k_col=ko, v=kv_all, v_col=ko + C)
            if pl.cross_short:
                ops.attn_cross_short(pl.attn_src[(level, C, "q")], src_kv, pl.cross_plan[level][0], pl.L, heads,
                                     pl.max_rows[level], pl.ctx_len, 0.125, q_mask=mask)
            else:
                ops.attn_varlen(pl.attn_src[(level, C, "q")], src_kv, *pl.cross_plan[level], 0.125, q_mask=mask)
            G(att, w[b + ".attn2.out.weight"], h, bias=w[b + ".attn2.out.bias"], epi=ops.EPI_GATE_RESID, resid=h,
              rowpart_out=rp)
            ln_gemm(".norm3", ".ff1", ff, epi=ops.EPI_GEGLU)
            G(ff, w[b + ".ff2.weight"], h, bias=w[b + ".ff2.bias"], epi=ops.EPI_GATE_RESID, resid=h, rowpart_out=rp)
        return G(h, w[name + ".proj_out.weight"], pbuf(name + ".out", T, C),
                 bias=w[name + ".proj_out.bias"], epi=ops.EPI_GATE_RESID, resid=x)

    # ------------------------------------------------------------------ forward
    def _plan(self, sample, ctx_len) -> _Plan:
        comp = tuple((res, t.shape[0], t.shape[-2], t.shape[-1])
                     for res, t in sample.items() if t is not None and t.shape[0] > 0)
        return self.plan_for(comp, ctx_len)

    def plan_for(self, comp, ctx_len) -> _Plan:
        """comp: ((resolution key, latents, h, w), ...) in ascending resolution order."""
        return self._plans.get((comp, ctx_len), lambda: _Plan(self, comp, ctx_len))

    # ---- per-request conditioning, computed once per request (sduss_b200.pipelines caches it)
    cond_kind = "sdxl"

    def project_context(self, ehs: torch.Tensor) -> torch.Tensor:
        """Text K / V of every cross-attention layer for raw prompt embeddings [n, 77, 2048] ->
        [n, 77, kv_cols] (attn2.to_k / to_v of all 70 BasicTransformerBlocks, modules/attention.py:
        73-96). They do not depend on the timestep nor on the image: once per request, not once per
        patch and step as the reference computes them."""
        n, ctx, d = ehs.shape
        out = torch.empty((n * ctx, self.kv_cols), device=self.device, dtype=torch.bfloat16)
        ops.gemm(ehs.reshape(n * ctx, d), self.w["kv_all.weight"], out)
        return out.view(n, ctx, -1)

    def kv_buffer(self, pl: _Plan) -> torch.Tensor:
        return pl.buf("kv_all", pl.L * pl.ctx_len, self.kv_cols)

    @torch.no_grad()
    def forward(self, sample: Dict[str, torch.Tensor], timestep, encoder_hidden_states,
                class_labels=None, timestep_cond=None, attention_mask=None,
                cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                down_intrablock_additional_residuals=None, encoder_attention_mask=None,
                return_dict: bool = True, record: bool = False, patch_size: Optional[int] = None,
                is_sliced: bool = True, save_index: int = 0, input_indices: Optional[dict] = None,
                _borrow: bool = False):
        assert (class_labels is None and timestep_cond is None and attention_mask is None
                and cross_attention_kwargs is None and down_block_additional_residuals is None
                and mid_block_additional_residual is None
                and down_intrablock_additional_residuals is None and encoder_attention_mask is None)
        pl = self._plan(sample, encoder_hidden_states.shape[1])

        def load_inputs(pl):
            for res, _, _, _ in pl.comp:
                pl.stage_in[res].copy_(sample[res])
            pl.ehs.copy_(encoder_hidden_states.reshape(pl.ehs.shape))
            pl.t32.copy_(timestep.reshape(-1))
            pl.ids32.copy_(added_cond_kwargs["time_ids"].reshape(-1))
            pl.text_embeds.copy_(added_cond_kwargs["text_embeds"])
            # text K/V of every cross-attention layer in one GEMM (once per latent, not per patch)
            ops.gemm(pl.ehs, self.w["kv_all.weight"], self.kv_buffer(pl))

        ops.run_plan(self, pl, load_inputs)
        out = pl.stage_out if _borrow else {k: v.clone() for k, v in pl.stage_out.items()}
        return (out,)

    # ---- patch cache (SDXL variant)
    def enable_patch_cache(self, forest, forest_up=None, refresh: int = 4, max_cached_plans: int = 3):
        """forest: ops.DeviceForest on [block, timestep, MSE of the block input against the previous
        step] for the down and mid blocks; forest_up (default: forest) for the up blocks, whose feature
        row continues with the MSE of each of the block's skip tensors (reference: the cuML forests of
        ESYMRED_DOWNSAMPLE_PATH / ESYMRED_UPSAMPLE_PATH, cache_manager.py:27-36,101-159; refresh = 4:
        forced recompute after four skips, :147). forest=None switches the cache off.
        What a clean patch skips: the block's Transformer2D modules (per 256-row patch of the packed pixel
        rows) and the convolutions of its resnets and samplers (per strip of 16 pixel rows); GroupNorms
        always run over the whole tensor, as in the reference."""
        import collections
        self.patch_forest, self.patch_forest_up, self.patch_refresh = forest, forest_up or forest, refresh
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
            pl.cache, pl.cached_state = _UNetPatchCache(self, pl), _CachedState()
        self._cache_lru[key] = pl
        self._cache_lru.move_to_end(key)
        return pl.cache

    def _decide(self, pl, cb, key, level, index, x, skips=()):
        """One get_mask call of the reference, on the device: the block's patch mask (or None when this
        level is not made of whole 256-row patches). An up block first takes the MSE of every skip
        tensor it is about to consume (in the order of the reference's res_hidden_states_tuple)."""
        st = cb.block(key, pl.levels[level], level)
        if st is None:
            return None
        extra = None
        if skips:
            extra = cb.buf(key + ".rmse", len(skips), st.n, torch.float32)
            for k, sk in enumerate(skips):
                ops.patch_mse(sk, cb.buf(f"{key}.rprev{k}", *sk.shape), st.patch_latent, pl.t32, cb.valid,
                              extra[k], st.ws)
        ops.patch_mask(x, cb.buf(key + ".xprev", *x.shape), st.patch_latent, pl.t32, cb.valid, st.skipped,
                       st.mask, self.patch_forest_up if skips else self.patch_forest, index,
                       self.patch_refresh, st.ws, mse=st.mse, extra_mse=extra)
        return st.mask

    def _run_cached(self, pl: _Plan):
        """The forward with the patch cache: _run plus one decision per UNet block (down blocks, mid, up
        blocks: the total_blocks numbering of modules/unet.py:369-503)."""
        return self._run(pl, pl.cache)

    def _run(self, pl: _Plan, cb=None):
        cfg, w, G = self.cfg, self.w, _G
        ch = cfg.block_out_channels
        L = pl.L
        Tdim = ch[0] * 4
        # ---- conditioning (A2): emb = time MLP + text_time MLP; temb projections of all resnets
        tsin = ops.timestep_embedding(pl.t32, ch[0], out=pl.buf("tsin", L, ch[0]))
        e1 = G(tsin, w["time_embedding.linear_1.weight"], pl.buf("e1", L, Tdim), bias=w["time_embedding.linear_1.bias"])
        ops.silu(e1, e1)
        emb_t = G(e1, w["time_embedding.linear_2.weight"], pl.buf("emb_t", L, Tdim), bias=w["time_embedding.linear_2.bias"])
        add_in = pl.buf("add_in", L, cfg.pooled_dim + 6 * cfg.addition_time_embed_dim)
        ids = ops.timestep_embedding(pl.ids32, cfg.addition_time_embed_dim,
                                     out=pl.buf("ids_sin", 6 * L, cfg.addition_time_embed_dim))
        ops.copy_cols(pl.text_embeds, add_in, cfg.pooled_dim)
        ops.copy_cols(ids.view(L, -1), add_in[:, cfg.pooled_dim:], 6 * cfg.addition_time_embed_dim)
        a1 = G(add_in, w["add_embedding.linear_1.weight"], pl.buf("a1", L, Tdim), bias=w["add_embedding.linear_1.bias"])
        ops.silu(a1, a1)
        emb = G(a1, w["add_embedding.linear_2.weight"], pl.buf("emb", L, Tdim),
                bias=w["add_embedding.linear_2.bias"], epi=ops.EPI_GATE_RESID, resid=emb_t)
        semb = ops.silu(emb, pl.buf("semb", L, Tdim))
        temb_all = G(semb, w["temb_all.weight"], pl.buf("temb_all", L, self.temb_cols), bias=w["temb_all.bias"])
        # ---- text K/V of every cross-attention layer: filled by forward() / the step's prologue
        kv_all = self.kv_buffer(pl)
        # ---- conv_in
        l0 = pl.levels[0]
        cols = pl.buf("im2col", l0.T, self.k_in_pad)
        ops.pack_im2col3x3(pl.in_ptr, l0.desc, L, l0.max_pixels, cfg.in_channels, cols)
        x = G(cols, w["conv_in.weight"], pl.buf("conv_in", l0.T, ch[0]), bias=w["conv_in.bias"])
        skips = [x]
        level = 0
        for i in range(len(ch)):
            mask = self._decide(pl, cb, f"down_blocks.{i}", level, i, x) if cb is not None else None
            for j in range(cfg.layers_per_block):
                x = self._resnet(pl, x, f"down_blocks.{i}.resnets.{j}", level, temb_all, cb, mask)
                if cfg.down_has_attn[i]:
                    x = self._transformer(pl, x, f"down_blocks.{i}.attentions.{j}", level,
                                          cfg.num_heads[i], cfg.transformer_layers_per_block[i], kv_all,
                                          cb, mask)
                skips.append(x)
            if i != len(ch) - 1:
                name = f"down_blocks.{i}.downsamplers.0.conv"
                x = self._conv(pl, x, x.shape[1], name, level, 2,
                               (cb.buf if cb is not None else pl.buf)(name, pl.levels[level + 1].T, x.shape[1]),
                               mask=mask, mask_scale=2, epi=ops.EPI_BIAS)
                level += 1
                skips.append(x)
        mask = self._decide(pl, cb, "mid_block", level, len(ch), x) if cb is not None else None
        x = self._resnet(pl, x, "mid_block.resnets.0", level, temb_all, cb, mask)
        x = self._transformer(pl, x, "mid_block.attentions.0", level, cfg.num_heads[-1],
                              cfg.transformer_layers_per_block[-1], kv_all, cb, mask)
        x = self._resnet(pl, x, "mid_block.resnets.1", level, temb_all, cb, mask)
        rev_layers = list(reversed(cfg.transformer_layers_per_block))
        rev_attn = list(reversed(cfg.down_has_attn))
        rev_heads = list(reversed(cfg.num_heads))
        for i in range(len(ch)):
            mask = self._decide(pl, cb, f"up_blocks.{i}", level, len(ch) + 1 + i, x,
                                skips[-(cfg.layers_per_block + 1):]) if cb is not None else None
            for j in range(cfg.layers_per_block + 1):
                skip = skips.pop()
                T, c1, c2 = x.shape[0], x.shape[1], skip.shape[1]
                cat = pl.buf(f"up_blocks.{i}.cat{j}", T, c1 + c2)
                ops.copy_cols(x, cat, c1)
                ops.copy_cols(skip, cat[:, c1:], c2)
                x = self._resnet(pl, cat, f"up_blocks.{i}.resnets.{j}", level, temb_all, cb, mask)
                if rev_attn[i]:
                    x = self._transformer(pl, x, f"up_blocks.{i}.attentions.{j}", level, rev_heads[i],
                                          rev_layers[i], kv_all, cb, mask)
            if i != len(ch) - 1:
                name = f"up_blocks.{i}.upsamplers.0.conv"
                lo, hi = pl.levels[level], pl.levels[level - 1]
                C = x.shape[1]
                up = pl.buf(name + ".up", hi.T, C)
                ops.upsample2x(x, lo.desc, hi.desc, L, hi.max_pixels, C, up)
                level -= 1
                x = self._conv(pl, up, C, name, level, 1,
                               (cb.buf if cb is not None else pl.buf)(name, hi.T, C),
                               mask=mask, mask_scale=-2, epi=ops.EPI_BIAS)
        h = self._gn(pl, x, "conv_norm_out", 0, pl.buf(f"gn0_{ch[0]}", l0.T, ch[0]), True)
        o = self._conv(pl, h, ch[0], "conv_out", 0, 1, pl.buf("conv_out", l0.T, self.n_out_pad), epi=ops.EPI_BIAS)
        ops.scatter_nchw(o, l0.desc, L, l0.max_pixels, cfg.out_channels, pl.out_ptr)
