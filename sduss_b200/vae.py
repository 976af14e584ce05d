"""VAE decode of a mixed-resolution batch of final latents on the B200 kernels (SURVEY.md §8 row
f-4, the stage right after the denoising path).

Replaces, behind the same call, what sduss' `post_inference` does per resolution
(pipeline_stable_diffusion_xl_esymred.py:406-462: `latents / scaling_factor`, fp32-upcast
`self.vae.decode`; pipeline_stable_diffusion_3_esymred.py:391-415: `latents / scaling_factor +
shift_factor`, `self.vae.decode`): here ALL requests of the batch, whatever their resolution, go
through ONE pass on the packed NHWC layout the UNet path uses (one buffer per decoder level:
latent res x1, x2, x4, x8), in bf16 with fp32 accumulation -- bf16 has fp32's exponent range, so
the fp16 overflow that forces the reference to upcast the SDXL VAE does not exist.

Kernels: `b200_latent_affine` (un-scaling + post_quant_conv), `b200_pack_im2col3x3` + GEMM
(conv_in), `b200_conv3x3_bf16` (resnets, upsampler convs, conv_out; residual add in the epilogue),
`b200_groupnorm_nhwc_bf16` (+SiLU), `b200_upsample2x_nhwc`, `b200_gemm_bf16` (1x1 shortcuts, the
attention linears and both attention contractions: head_dim = 512 does not fit the head_dim-64
packed attention kernel), `b200_softmax_rows`, `b200_scatter_nchw`. No PyTorch arithmetic.
"""
import math
from dataclasses import dataclass, fields
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .layout import LevelLayout


@dataclass
class VAEDecoderConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.13025
    shift_factor: Optional[float] = None
    use_post_quant_conv: bool = True

    @classmethod
    def from_any(cls, cfg):
        if isinstance(cfg, cls):
            return cfg
        get = cfg.get if hasattr(cfg, "get") else lambda k, d=None: getattr(cfg, k, d)
        kw = {}
        for f in fields(cls):
            v = get(f.name, None)
            if v is not None:
                kw[f.name] = tuple(v) if isinstance(v, list) else v
        return cls(**kw)


class _Plan:
    def __init__(self, model: "B200VAEDecoder", comp):
        dev, cfg = model.device, model.cfg
        self.comp = comp
        sizes0 = [(h, w) for _, n, h, w in comp for _ in range(n)]
        self.L = L = len(sizes0)
        nlev = len(cfg.block_out_channels)
        self.levels = [LevelLayout([(h << l, w << l) for h, w in sizes0], dev) for l in range(nlev)]
        bf = dict(device=dev, dtype=torch.bfloat16)
        C, Co = cfg.latent_channels, cfg.out_channels
        up = 1 << (nlev - 1)
        n_in = [n * C * h * w for _, n, h, w in comp]
        n_out = [n * Co * h * w * up * up for _, n, h, w in comp]
        self.flat_in = torch.empty((sum(n_in),), **bf)    # request latents, NCHW
        self.flat_z = torch.empty((sum(n_in),), **bf)     # un-scaled (+ post_quant_conv) latents
        self.flat_out = torch.empty((sum(n_out),), **bf)  # images, NCHW
        self.stage_in, self.stage_out = {}, {}
        ip, zp, op = [], [], []
        oi = oo = 0
        for (res, n, h, w), ni, no in zip(comp, n_in, n_out):
            self.stage_in[res] = self.flat_in[oi:oi + ni].view(n, C, h, w)
            z = self.flat_z[oi:oi + ni].view(n, C, h, w)
            self.stage_out[res] = self.flat_out[oo:oo + no].view(n, Co, h * up, w * up)
            for i in range(n):
                ip.append(self.stage_in[res][i].data_ptr())
                zp.append(z[i].data_ptr())
                op.append(self.stage_out[res][i].data_ptr())
            oi, oo = oi + ni, oo + no
        as_dev = lambda p: torch.tensor(p, dtype=torch.int64).to(dev)
        self.in_ptr, self.z_ptr, self.out_ptr = as_dev(ip), as_dev(zp), as_dev(op)
        self.gn_ws = ops.groupnorm_workspace(self.levels[-1].T, L, dev)
        self.arena, self.block, self.arena_off = model.arena, None, 0
        self.bufs: Dict[str, torch.Tensor] = {}
        self.maps: Dict[tuple, torch.Tensor] = {}
        self.stats_tag, self.stats_gen = {}, {}   # conv-epilogue GroupNorm statistics (ops.conv_stats_buffer)
        self.device = dev
        self.graph = None
        self.graph_launches = 0

    def buf(self, name, rows, cols):
        t = self.bufs.get(name)
        if t is None:  # bump-allocated from the model's shared arena (ops.Arena.alloc)
            t = self.bufs[name] = self.arena.alloc(self, (rows, cols), torch.bfloat16)
        assert t.shape == (rows, cols), (name, t.shape, rows, cols)
        return t

    def ws(self, name, numel, dtype):
        t = self.bufs.get(name)
        if t is None:
            t = self.bufs[name] = self.arena.alloc(self, (numel,), dtype)
        return t

    def reset_workspaces(self):
        """Forget every arena view and everything that captured its address (ops._run_eager)."""
        self.bufs, self.maps, self.block, self.arena_off = {}, {}, None, 0
        self.stats_tag, self.stats_gen = {}, {}
        if hasattr(self, "attn_src"):
            self.attn_src = {}

    def conv_maps(self, x, cin, level):
        key = (x.data_ptr(), x.stride(0), cin, level)
        m = self.maps.get(key)
        if m is None:
            m = self.maps[key] = ops.conv3x3_encode_maps(x, cin, self.levels[level].desc_host, 1)
        return m


class B200VAEDecoder(torch.nn.Module):
    """`decode(latents_by_resolution)`: {res: [n, C, h, w]} scheduler-space latents of the finished
    requests -> {res: [n, 3, 8h, 8w]} images in [-1, 1] (what `vae.decode(...)[0]` returns in the
    reference, for every resolution at once)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], config, device="cuda"):
        super().__init__()
        self.cfg = cfg = VAEDecoderConfig.from_any(config)
        self.config = config
        self.device = dev = torch.device(device)
        self.dtype = torch.bfloat16
        # the decoder is small (50 M parameters): re-laid out on the host in fp32, whatever device /
        # dtype the state dict arrives in (encoder and quant_conv entries are ignored)
        sd = {k: v.detach().float().cpu() for k, v in state_dict.items()
              if k.startswith(("decoder.", "post_quant_conv."))}
        self.w: Dict[str, torch.Tensor] = {}

        def put(name, t):
            self.w[name] = t.to(device=dev, dtype=torch.bfloat16).contiguous()

        def conv3(name, pad_out=None):
            wt, b = sd[name + ".weight"].float(), sd[name + ".bias"].float()
            if pad_out is not None and wt.shape[0] < pad_out:
                wt = torch.cat([wt, torch.zeros(pad_out - wt.shape[0], *wt.shape[1:])])
                b = torch.cat([b, torch.zeros(pad_out - b.shape[0])])
            put(name + ".weight", wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1))
            put(name + ".bias", b)

        def norm(name):
            put(name + ".weight", sd[name + ".weight"])
            put(name + ".bias", sd[name + ".bias"])

        def resnet(name):
            norm(name + ".norm1"), norm(name + ".norm2")
            conv3(name + ".conv1"), conv3(name + ".conv2")
            if name + ".conv_shortcut.weight" in sd:
                wt = sd[name + ".conv_shortcut.weight"]
                put(name + ".conv_shortcut.weight", wt.reshape(wt.shape[0], wt.shape[1]))
                put(name + ".conv_shortcut.bias", sd[name + ".conv_shortcut.bias"])

        C = cfg.latent_channels
        ch = list(reversed(cfg.block_out_channels))
        self.ch = ch
        # latents / scaling_factor (+ shift_factor), then the 1x1 post_quant_conv: one affine map
        A = torch.eye(C, dtype=torch.float64) / cfg.scaling_factor
        b = torch.full((C,), float(cfg.shift_factor or 0.0), dtype=torch.float64)
        if cfg.use_post_quant_conv:
            Wpq = sd["post_quant_conv.weight"].double().reshape(C, C)
            A, b = Wpq @ A, Wpq @ b + sd["post_quant_conv.bias"].double()
        self.pre_w = A.float().contiguous().to(dev)
        self.pre_b = b.float().contiguous().to(dev)
        # the same map for latents the caller has already un-scaled (the reference's post_inference
        # does `latents / scaling_factor (+ shift_factor)` itself before `vae.decode`): B200VAEProxy
        Au, bu = torch.eye(C, dtype=torch.float64), torch.zeros(C, dtype=torch.float64)
        if cfg.use_post_quant_conv:
            Au, bu = Wpq, sd["post_quant_conv.bias"].double()
        self.pre_w_unscaled = Au.float().contiguous().to(dev)
        self.pre_b_unscaled = bu.float().contiguous().to(dev)
        # conv_in as a GEMM on im2col rows: [Cout, Cin*9] padded to a multiple of 64
        k_in = C * 9
        self.k_in_pad = ((k_in + 63) // 64) * 64
        wi = torch.zeros(ch[0], self.k_in_pad)
        wi[:, :k_in] = sd["decoder.conv_in.weight"].reshape(ch[0], k_in).float()
        put("decoder.conv_in.weight", wi)
        put("decoder.conv_in.bias", sd["decoder.conv_in.bias"])
        resnet("decoder.mid_block.resnets.0")
        a = "decoder.mid_block.attentions.0"
        norm(a + ".group_norm")
        put(a + ".qkv.weight", torch.cat([sd[f"{a}.to_{n}.weight"] for n in "qkv"], 0))
        # the value bias is added after P V (rows of P sum to one): qkv bias = [bq | bk | 0]
        put(a + ".qkv.bias", torch.cat([sd[a + ".to_q.bias"], sd[a + ".to_k.bias"],
                                        torch.zeros_like(sd[a + ".to_v.bias"])]))
        put(a + ".v.bias", sd[a + ".to_v.bias"])
        put(a + ".out.weight", sd[a + ".to_out.0.weight"])
        put(a + ".out.bias", sd[a + ".to_out.0.bias"])
        resnet("decoder.mid_block.resnets.1")
        for i in range(len(ch)):
            for j in range(cfg.layers_per_block + 1):
                resnet(f"decoder.up_blocks.{i}.resnets.{j}")
            if i != len(ch) - 1:
                conv3(f"decoder.up_blocks.{i}.upsamplers.0.conv")
        norm("decoder.conv_norm_out")
        # conv_out (3 channels) padded to 64: the epilogue's TMA store then writes whole 128-byte
        # pixel rows; with 8 channels (16-byte rows) this one layer took 1.13 ms instead of 0.48 ms
        # on 1.3 M pixels (profiles/r01_vae_decode.txt)
        self.n_out_pad = 64
        conv3("decoder.conv_out", pad_out=self.n_out_pad)
        self.arena = ops.Arena(self.device)  # per-step workspaces of all plans overlap here
        self._plans = ops.PlanCache(self.device, arena=self.arena)
        self.use_graphs = ops.graphs_enabled()
        # GroupNorm statistics from the producing convolution's epilogue (one read of the tensor).
        # Built, parity-tested and OFF: measured on the decoder (profiles/r02_gn_conv_stats.txt) the
        # column sums cost the convolutions +1.06 ms while the GroupNorms cannot save more than their
        # statistics pass (a third of 3.8 ms): no net gain. DESIGN.md section 4.7.
        self.fuse_gn_stats = False

    @classmethod
    def from_diffusers(cls, vae, device="cuda"):
        """`vae`: the diffusers AutoencoderKL the sduss pipelines hold as `self.vae`."""
        return cls(vae.state_dict(), vae.config, device=device)

    # ------------------------------------------------------------------ building blocks
    def _gn(self, pl, x, name, level, out, silu):
        lay = pl.levels[level]
        kw = dict(groups=self.cfg.norm_num_groups, eps=self.cfg.norm_eps, silu=silu)
        stats = ops.fresh_conv_stats(pl, x)
        if stats is not None:  # the convolution that wrote x left its statistics: x is read once
            return ops.groupnorm_from_conv_stats(x, out, self.w[name + ".weight"], self.w[name + ".bias"],
                                                 lay.row_group, stats, lay.lat_tiles, lay.L, pl.gn_ws, **kw)
        ops.groupnorm_nhwc(x, out, self.w[name + ".weight"], self.w[name + ".bias"], lay.row_group,
                           lay.lat_chunks, lay.L, pl.gn_ws, **kw)
        return out

    def _conv(self, pl, x, cin, name, level, out, resid=None, epi=ops.EPI_BIAS, stats=True):
        lay = pl.levels[level]
        w = self.w[name + ".weight"]
        cout = w.shape[0]
        st = ops.conv_stats_buffer(pl, out, level, lay.n_tiles, cout) \
            if (stats and self.fuse_gn_stats and cout % self.cfg.norm_num_groups == 0) else None
        return ops.conv3x3(pl.conv_maps(x, cin, level), lay.tiles, lay.n_tiles, lay.desc, cin, cout, 1,
                           w, out, out_maps=pl.conv_maps(out, cout, level),
                           resid_maps=pl.conv_maps(resid, cout, level) if resid is not None else None,
                           bias=self.w[name + ".bias"], epi=epi, stats_out=st)

    def _resnet(self, pl, x, name, level, out_name):
        """ResnetBlock2D without time embedding (diffusers Decoder): x + conv2(act(conv1(act(x))))."""
        T, cin = x.shape
        cout = self.w[name + ".conv1.weight"].shape[0]
        h = self._gn(pl, x, name + ".norm1", level, pl.buf(f"gn{level}_{cin}", T, cin), True)
        h1 = self._conv(pl, h, cin, name + ".conv1", level, pl.buf(f"h1_{level}_{cout}", T, cout))
        h2 = self._gn(pl, h1, name + ".norm2", level, pl.buf(f"gn{level}_{cout}", T, cout), True)
        if name + ".conv_shortcut.weight" in self.w:
            s = ops.gemm(x, self.w[name + ".conv_shortcut.weight"], pl.buf(f"sc_{level}_{cout}", T, cout),
                         bias=self.w[name + ".conv_shortcut.bias"])
        else:
            s = x
        return self._conv(pl, h2, cout, name + ".conv2", level, pl.buf(out_name, T, cout),
                          resid=s, epi=ops.EPI_GATE_RESID)

    def _attention(self, pl, x, name, out_name):
        """diffusers Attention(heads=1, dim_head=C, group_norm, residual_connection): per latent
        S = q k^T (GEMM, fp32 logits) -> row softmax -> P v (GEMM against v^T) on the level-0 rows."""
        lay = pl.levels[0]
        T, C = x.shape
        G = ops.gemm
        t = self._gn(pl, x, name + ".group_norm", 0, pl.buf(f"gn0_{C}", T, C), False)
        qkv = G(t, self.w[name + ".qkv.weight"], pl.buf("attn_qkv", T, 3 * C), bias=self.w[name + ".qkv.bias"])
        o = pl.buf("attn_o", T, C)
        vt = pl.buf("attn_vt", C, lay.max_pixels)
        wv = self.w[name + ".qkv.weight"][2 * C:]
        nmax = lay.max_pixels
        attn_s = pl.ws("attn_s", nmax * nmax, torch.float32)
        attn_p = pl.ws("attn_p", nmax * nmax, torch.bfloat16)
        for i in range(lay.L):
            r0, n = lay.row_off[i], lay.rows[i]
            s = attn_s[:n * n].view(n, n)
            p = attn_p[:n * n].view(n, n)
            G(qkv[r0:r0 + n, :C], qkv[r0:r0 + n, C:2 * C], s)            # fp32 logits
            ops.softmax_rows(s, p, 1.0 / math.sqrt(C))
            G(wv, t[r0:r0 + n], vt[:, :n])                                 # v^T = Wv t^T
            G(p, vt[:, :n], o[r0:r0 + n], bias=self.w[name + ".v.bias"])  # P v + bv
        return G(o, self.w[name + ".out.weight"], pl.buf(out_name, T, C), bias=self.w[name + ".out.bias"],
                 epi=ops.EPI_GATE_RESID, resid=x)

    # ------------------------------------------------------------------ forward
    def _plan(self, latents, unscaled=False) -> _Plan:
        comp = tuple((res, t.shape[0], t.shape[-2], t.shape[-1])
                     for res, t in latents.items() if t is not None and t.shape[0] > 0)

        def make():
            pl = _Plan(self, comp)
            pl.unscaled = unscaled
            return pl
        return self._plans.get((comp, unscaled), make)

    @torch.no_grad()
    def decode(self, latents: Dict[str, torch.Tensor], _borrow: bool = False,
               unscaled: bool = False) -> Dict[str, torch.Tensor]:
        """unscaled=True: `latents` are already `latents / scaling_factor (+ shift_factor)` (what the
        reference hands to `vae.decode`); default: scheduler-space latents, un-scaled here."""
        pl = self._plan(latents, unscaled)
        for res, _, _, _ in pl.comp:
            pl.stage_in[res].copy_(latents[res])
        ops.run_plan(self, pl)
        return pl.stage_out if _borrow else {k: v.clone() for k, v in pl.stage_out.items()}

    def _run(self, pl: _Plan):
        cfg, w, ch = self.cfg, self.w, self.ch
        L, l0 = pl.L, pl.levels[0]
        C = cfg.latent_channels
        pre_w, pre_b = ((self.pre_w_unscaled, self.pre_b_unscaled) if getattr(pl, "unscaled", False)
                        else (self.pre_w, self.pre_b))
        ops.latent_affine(pl.in_ptr, pl.z_ptr, l0.desc, L, l0.max_pixels, C, C, pre_w, pre_b)
        cols = pl.buf("im2col", l0.T, self.k_in_pad)
        ops.pack_im2col3x3(pl.z_ptr, l0.desc, L, l0.max_pixels, C, cols)
        x = ops.gemm(cols, w["decoder.conv_in.weight"], pl.buf("act0_a", l0.T, ch[0]),
                     bias=w["decoder.conv_in.bias"])
        x = self._resnet(pl, x, "decoder.mid_block.resnets.0", 0, "act0_b")
        x = self._attention(pl, x, "decoder.mid_block.attentions.0", "act0_a")
        x = self._resnet(pl, x, "decoder.mid_block.resnets.1", 0, "act0_b")
        level, flip = 0, 0
        for i, c in enumerate(ch):
            for j in range(cfg.layers_per_block + 1):
                # outputs alternate between two buffers per (level, width); x is never its own output
                x = self._resnet(pl, x, f"decoder.up_blocks.{i}.resnets.{j}", level, f"act{level}_{c}_{flip}")
                flip ^= 1
            if i != len(ch) - 1:
                name = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                lo, hi = pl.levels[level], pl.levels[level + 1]
                up = pl.buf(f"up{level + 1}_{c}", hi.T, c)
                ops.upsample2x(x, lo.desc, hi.desc, L, hi.max_pixels, c, up)
                level += 1
                x = self._conv(pl, up, c, name, level, pl.buf(f"act{level}_{c}_{flip}", hi.T, c))
                flip ^= 1
        lt = pl.levels[level]
        h = self._gn(pl, x, "decoder.conv_norm_out", level, pl.buf(f"gn{level}_{ch[-1]}", lt.T, ch[-1]), True)
        o = self._conv(pl, h, ch[-1], "decoder.conv_out", level, pl.buf("conv_out", lt.T, self.n_out_pad), stats=False)
        ops.scatter_nchw(o, lt.desc, L, lt.max_pixels, cfg.out_channels, pl.out_ptr)


def postprocess(image: torch.Tensor) -> torch.Tensor:
    """VaeImageProcessor.postprocess up to the float image (denormalize, NHWC); the PIL conversion
    and the mp.Queue hand-over stay with the reference's runner (runner/wrappers.py:58-66)."""
    return (image.float() / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1)


class B200VAEProxy:
    """Stands in for the diffusers AutoencoderKL a sduss pipeline holds as `self.vae`, so that the
    reference's own `post_inference` (pipeline_stable_diffusion_xl_esymred.py:406-462,
    pipeline_stable_diffusion_3_esymred.py:391-415) runs unmodified: `decode(z)` takes the latents
    the reference has already un-scaled and returns `(image,)` / an object with `.sample`; every
    other attribute (`config`, `dtype`, `to`, ...) is the wrapped module's."""

    def __init__(self, vae, decoder: "B200VAEDecoder" = None, device="cuda"):
        self.__dict__["_vae"] = vae
        self.__dict__["_decoder"] = decoder if decoder is not None else B200VAEDecoder.from_diffusers(vae, device)

    def __getattr__(self, name):
        return getattr(self.__dict__["_vae"], name)

    def decode(self, z, return_dict: bool = True, generator=None):
        image = self.__dict__["_decoder"].decode({"_": z}, unscaled=True)["_"]
        if not return_dict:
            return (image,)
        from types import SimpleNamespace
        return SimpleNamespace(sample=image)
