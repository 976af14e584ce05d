"""SDXL-path kernels on the packed NHWC layout vs PyTorch fp32 references / reference fixtures."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _pack(lats):  # list of [C,h,w] -> [sum hw, C]
    return torch.cat([t.permute(1, 2, 0).reshape(-1, t.shape[0]) for t in lats])


def _unpack(x, sizes):
    out, o = [], 0
    for h, w in sizes:
        out.append(x[o:o + h * w].reshape(h, w, -1).permute(2, 0, 1))
        o += h * w
    return out


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


@pytest.mark.parametrize("sizes,cin,cout,stride", [
    ([(32, 32), (64, 64)], 64, 128, 1),
    ([(8, 8), (16, 16), (24, 24)], 128, 64, 1),
    ([(64, 64), (128, 128)], 320, 320, 1),
    ([(32, 32), (48, 48)], 640, 1280, 1),
    ([(32, 32), (64, 64), (96, 96)], 64, 64, 2),
    ([(16, 16), (48, 48)], 320, 320, 2),
    ([(32, 32)], 320, 8, 1),
    # many output tiles per CTA with a short K loop and an odd number of 64-column chunks per
    # tile (Cout = 320 = 256 + 64; Cout = 8): regression for the epilogue staging-buffer reuse
    ([(128, 128), (128, 128), (64, 64)], 64, 320, 1),
    ([(128, 128), (128, 128)], 64, 8, 1),
    # enough 128-wide work items for the two-M-tiles-per-CTA variant (VAE decoder levels), with an
    # odd number of M tiles (the last pair is half empty), stride 1 and 2
    ([(512, 320), (16, 8)], 64, 128, 1),
    ([(1024, 640), (32, 16)], 64, 64, 2),
])
def test_conv3x3(cuda, sizes, cin, cout, stride):
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    lats = [_rand((cin, h, w), 10 + i).bfloat16().float() for i, (h, w) in enumerate(sizes)]
    wgt = _rand((cout, cin, 3, 3), 1, 1.0 / (3 * cin ** 0.5)).bfloat16().float()
    bias = _rand((cout,), 2).bfloat16().float()
    lin = LevelLayout(sizes, cuda)
    osz = [(h // stride, w // stride) for h, w in sizes]
    lout = LevelLayout(osz, cuda)
    x = _pack(lats).cuda().bfloat16().contiguous()
    wt = wgt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).cuda().bfloat16().contiguous()
    maps = ops.conv3x3_encode_maps(x, cin, lin.desc_host, stride)
    out = torch.zeros(lout.T, cout, device=cuda, dtype=torch.bfloat16)
    resid = _rand((lout.T, cout), 3).cuda().bfloat16()
    omaps = ops.conv3x3_encode_maps(out, cout, lout.desc_host, 1)
    rmaps = ops.conv3x3_encode_maps(resid, cout, lout.desc_host, 1)
    ops.conv3x3(maps, lout.tiles, lout.n_tiles, lout.desc, cin, cout, stride, wt, out, out_maps=omaps,
                resid_maps=rmaps, epi=ops.EPI_GATE_RESID, bias=bias.cuda().bfloat16())
    ref = [F.conv2d(t[None], wgt, bias, stride=stride, padding=1)[0] for t in lats]
    ref = _pack(ref) + resid.float().cpu()
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 2e-2 * ref.abs().max().item(), err
    # per-request row vector epilogue (time embedding add)
    rv = _rand((len(sizes), cout), 4).cuda().bfloat16()
    ops.conv3x3(maps, lout.tiles, lout.n_tiles, lout.desc, cin, cout, stride, wt, out, out_maps=omaps,
                epi=ops.EPI_ROWVEC, bias=bias.cuda().bfloat16(), rowvec=rv, row_group=lout.row_group)
    ref2 = _pack([F.conv2d(t[None], wgt, bias, stride=stride, padding=1)[0] + rv[i].float().cpu()[:, None, None]
                  for i, t in enumerate(lats)])
    err = (out.float().cpu() - ref2).abs().max().item()
    assert err <= 2e-2 * ref2.abs().max().item(), err


@pytest.mark.parametrize("sizes,C,silu,eps", [([(32, 32), (64, 64)], 320, True, 1e-5),
                                               ([(8, 8), (16, 16), (24, 24)], 64, False, 1e-6),
                                               ([(16, 16), (32, 32)], 2560, True, 1e-5),
                                               ([(64, 64)], 960, True, 1e-5)])
def test_groupnorm(cuda, sizes, C, silu, eps):
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    lats = [(_rand((C, h, w), 20 + i) * (1 + i) + 0.5 * i).bfloat16().float() for i, (h, w) in enumerate(sizes)]
    gam, bet = (1 + 0.1 * _rand((C,), 5)).bfloat16().float(), _rand((C,), 6).bfloat16().float()
    lay = LevelLayout(sizes, cuda)
    x = _pack(lats).cuda().bfloat16().contiguous()
    y = torch.empty_like(x)
    ws = ops.groupnorm_workspace(lay.T, lay.L, cuda)
    ops.groupnorm_nhwc(x, y, gam.cuda().bfloat16(), bet.cuda().bfloat16(), lay.row_group,
                       lay.lat_chunks, lay.L, ws, eps=eps, silu=silu)
    ref = [F.group_norm(t[None], 32, gam, bet, eps)[0] for t in lats]
    if silu:
        ref = [F.silu(t) for t in ref]
    err = (y.float().cpu() - _pack(ref)).abs().max().item()
    assert err < 4e-2, err
    y2 = torch.empty_like(x)
    ops.groupnorm_nhwc(x, y2, gam.cuda().bfloat16(), bet.cuda().bfloat16(), lay.row_group,
                       lay.lat_chunks, lay.L, ws, eps=eps, silu=silu)
    assert torch.equal(y, y2)  # deterministic (no atomics)


@pytest.mark.parametrize("sizes,C", [([(32, 32), (64, 64)], 1280), ([(128, 128), (64, 64), (128, 128)], 320),
                                     ([(256, 256)], 128), ([(8, 8)], 64)])
def test_groupnorm_single_launch_equals_three_launches(cuda, sizes, C, monkeypatch):
    """The opt-in single-launch kernel (statistics -> grid barrier -> finalize -> grid barrier -> apply; persistent
    grid, several items per CTA on the larger tensors) runs the same arithmetic as the three-launch path:
    bit-identical, also when repeated on one workspace (the barrier epoch keeps counting)."""
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    lats = [(_rand((C, h, w), 70 + i) * (1 + i) - 0.3 * i).bfloat16().float() for i, (h, w) in enumerate(sizes)]
    gam, bet = (1 + 0.1 * _rand((C,), 5)).cuda().bfloat16(), _rand((C,), 6).cuda().bfloat16()
    lay = LevelLayout(sizes, cuda)
    x = _pack(lats).cuda().bfloat16().contiguous()
    ws = ops.groupnorm_workspace(lay.T, lay.L, cuda)
    outs = []
    for mode in ("1", "0", "1", "1"):
        monkeypatch.setenv("SDUSS_B200_GN_FUSED", mode)
        y = torch.empty_like(x)
        ops.groupnorm_nhwc(x, y, gam, bet, lay.row_group, lay.lat_chunks, lay.L, ws, silu=True)
        torch.cuda.synchronize()
        outs.append(y)
    for y in outs[1:]:
        assert torch.equal(outs[0], y)
    ref = _pack([F.silu(F.group_norm(t[None], 32, gam.float().cpu(), bet.float().cpu(), 1e-5)[0]) for t in lats])
    assert (outs[0].float().cpu() - ref).abs().max().item() < 4e-2


@pytest.mark.parametrize("C", [320, 1280])
def test_groupnorm_batch_invariant(cuda, C):
    """A latent's GroupNorm output must not depend on the latents packed around it (the launch
    geometry changes with the total row count; the summation order may not)."""
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    sizes = [(16, 16), (64, 64), (128, 128), (32, 32)]
    lats = [(_rand((C, h, w), 40 + i) * (1 + i)).bfloat16().float() for i, (h, w) in enumerate(sizes)]
    gam, bet = (1 + 0.1 * _rand((C,), 5)).cuda().bfloat16(), _rand((C,), 6).cuda().bfloat16()

    def run(idx):
        lay = LevelLayout([sizes[i] for i in idx], cuda)
        x = _pack([lats[i] for i in idx]).cuda().bfloat16().contiguous()
        y = torch.empty_like(x)
        ws = ops.groupnorm_workspace(lay.T, lay.L, cuda)
        ops.groupnorm_nhwc(x, y, gam, bet, lay.row_group, lay.lat_chunks, lay.L, ws, silu=True)
        torch.cuda.synchronize()
        return y, lay
    mixed, lay = run(range(len(sizes)))
    off = 0
    for i, (h, w) in enumerate(sizes):
        solo, _ = run([i])
        assert torch.equal(mixed[off:off + h * w], solo), i
        off += h * w


def test_pack_scatter_upsample_copy(cuda):
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    g = torch.Generator().manual_seed(0)
    sizes = [(64, 64), (64, 64), (128, 128)]
    lats = [torch.randint(-100, 100, (4, h, w), generator=g).float() for h, w in sizes]
    dl = [t.cuda().bfloat16().contiguous() for t in lats]
    lay = LevelLayout(sizes, cuda)
    ptr = torch.tensor([t.data_ptr() for t in dl], dtype=torch.int64).cuda()
    cols = torch.full((lay.T, 64), 7.0, device=cuda, dtype=torch.bfloat16)
    ops.pack_im2col3x3(ptr, lay.desc, lay.L, lay.max_pixels, 4, cols)
    ref = torch.cat([F.unfold(t[None], 3, padding=1)[0].t() for t in lats])  # [hw, C*9], (c,ky,kx)
    assert torch.equal(cols[:, :36].float().cpu(), ref)
    assert torch.count_nonzero(cols[:, 36:]) == 0
    # scatter NHWC -> NCHW
    x = torch.randint(-100, 100, (lay.T, 8), generator=g).float()
    outs = [torch.zeros_like(t) for t in dl]
    optr = torch.tensor([t.data_ptr() for t in outs], dtype=torch.int64).cuda()
    ops.scatter_nchw(x.cuda().bfloat16(), lay.desc, lay.L, lay.max_pixels, 4, optr)
    for o, r in zip(outs, _unpack(x[:, :4], sizes)):
        assert torch.equal(o.float().cpu(), r)
    # nearest upsample
    small = [(h // 2, w // 2) for h, w in sizes]
    ls = LevelLayout(small, cuda)
    xs = torch.randint(-100, 100, (ls.T, 64), generator=g).float()
    up = torch.empty(lay.T, 64, device=cuda, dtype=torch.bfloat16)
    ops.upsample2x(xs.cuda().bfloat16(), ls.desc, lay.desc, lay.L, lay.max_pixels, 64, up)
    ref = _pack([F.interpolate(t[None], scale_factor=2.0, mode="nearest")[0] for t in _unpack(xs, small)])
    assert torch.equal(up.float().cpu(), ref)
    # column-block copy
    dst = torch.zeros(lay.T, 192, device=cuda, dtype=torch.bfloat16)
    ops.copy_cols(up, dst[:, 128:], 64)
    assert torch.equal(dst[:, 128:], up) and torch.count_nonzero(dst[:, :128]) == 0


def test_reference_patch_format_bit_exact(cuda):
    """b200_split_patches / b200_concat_patches reproduce PatchUNet.split_sample /
    concat_sample (modules/unet.py:104-202) bit for bit on the reference-generated fixture."""
    from oracle import pack as opack
    from sduss_b200 import ops
    z = np.load(os.path.join(G, "pack_sdxl.npz"))
    for tag in ("a", "b"):
        res = sorted((k[len(tag) + 4:] for k in z.files if k.startswith(tag + "_in_")), key=int)
        # fixture values are int16-range; keep them bf16-exact by taking them modulo 256
        lats, ldesc, pdesc = [], [], []
        for r in res:
            arr = torch.from_numpy(z[f"{tag}_in_{r}"].astype(np.float32)) % 256
            for t in arr:
                lats.append(t.cuda().bfloat16().contiguous())
        ref_patches = torch.from_numpy(z[f"{tag}_patches"].astype(np.float32))
        lat_off = z[f"{tag}_latent_offset"]
        for l, t in enumerate(lats):
            ldesc.append((0, t.shape[1], t.shape[2], 0))
            ph = t.shape[1] // 32
            assert lat_off[l + 1] - lat_off[l] == ph * ph
            for h in range(ph):
                for w in range(ph):
                    pdesc.append((l, h, w, 0))
        P = len(pdesc)
        assert P == ref_patches.shape[0]
        ptr = torch.tensor([t.data_ptr() for t in lats], dtype=torch.int64).cuda()
        ld = torch.tensor(ldesc, dtype=torch.int32).cuda()
        pd = torch.tensor(pdesc, dtype=torch.int32).cuda()
        out = torch.empty(P, 4, 34, 34, device=cuda, dtype=torch.bfloat16)
        ops.split_patches(ptr, ld, pd, P, 4, 32, out)
        # zero halo stays zero under the modulo; interior values are taken modulo 256
        ref = torch.where(ref_patches == 0, ref_patches, ref_patches % 256)
        mask = torch.zeros(34, 34, dtype=torch.bool)
        assert torch.equal(out.float().cpu() % 256, ref % 256)
        inner = out[:, :, 1:-1, 1:-1].contiguous()
        outs = [torch.zeros_like(t) for t in lats]
        optr = torch.tensor([t.data_ptr() for t in outs], dtype=torch.int64).cuda()
        ops.concat_patches(inner, ld, pd, P, 4, 32, optr)
        for a, b in zip(outs, lats):
            assert torch.equal(a, b)  # concat(split(x)) == x


@pytest.mark.parametrize("sizes,cin,cout,stride,epi", [
    ([(32, 32), (64, 64)], 64, 128, 1, "bias"),                 # VAE-like: 4 channels per group
    ([(8, 8), (16, 16), (24, 24)], 128, 320, 1, "rowvec"),      # groups straddle 64-column chunks, partial tiles
    ([(16, 16), (48, 48)], 320, 640, 2, "bias"),                # stride 2
    ([(64, 64), (128, 128)], 64, 320, 1, "resid"),              # residual epilogue (staging tile is refilled)
    ([(512, 320), (16, 8)], 64, 128, 1, "resid"),               # two M tiles per CTA
])
def test_groupnorm_from_conv_epilogue_statistics(cuda, sizes, cin, cout, stride, epi):
    """The convolution leaves per-(tile, half, channel) sums of its STORED outputs; GroupNorm built on
    them (one read of the tensor) must agree with the three-launch GroupNorm on the same tensor and
    with torch, be deterministic, and not depend on the latents packed around a latent."""
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    osz = [(h // stride, w // stride) for h, w in sizes]

    def run(sel):
        lats = [_rand((cin, *sizes[i]), 10 + i).bfloat16().float() for i in sel]
        lin, lout = LevelLayout([sizes[i] for i in sel], cuda), LevelLayout([osz[i] for i in sel], cuda)
        x = _pack(lats).cuda().bfloat16().contiguous()
        wgt = _rand((cout, cin, 3, 3), 1, 1.0 / (3 * cin ** 0.5)).bfloat16()
        wt = wgt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).cuda().contiguous()
        bias = _rand((cout,), 2).cuda().bfloat16()
        out = torch.zeros(lout.T, cout, device=cuda, dtype=torch.bfloat16)
        stats = torch.zeros(lout.n_tiles * 2 * cout * 2, device=cuda)
        kw = dict(out_maps=ops.conv3x3_encode_maps(out, cout, lout.desc_host, 1), bias=bias, stats_out=stats)
        if epi == "resid":
            resid = _pack([_rand((cout, *osz[i]), 30 + i) for i in sel]).cuda().bfloat16().contiguous()
            kw.update(epi=ops.EPI_GATE_RESID, resid_maps=ops.conv3x3_encode_maps(resid, cout, lout.desc_host, 1))
        elif epi == "rowvec":
            kw.update(epi=ops.EPI_ROWVEC, rowvec=_rand((len(sizes), cout), 4).cuda().bfloat16()[sel].contiguous(),
                      row_group=lout.row_group)
        ops.conv3x3(ops.conv3x3_encode_maps(x, cin, lin.desc_host, stride), lout.tiles, lout.n_tiles, lout.desc,
                    cin, cout, stride, wt, out, **kw)
        gam, bet = (1 + 0.1 * _rand((cout,), 5)).cuda().bfloat16(), _rand((cout,), 6).cuda().bfloat16()
        ws = ops.groupnorm_workspace(lout.T, lout.L, cuda)
        y = torch.empty_like(out)
        ops.groupnorm_from_conv_stats(out, y, gam, bet, lout.row_group, stats, lout.lat_tiles, lout.L, ws, silu=True)
        y3 = torch.empty_like(out)
        ops.groupnorm_nhwc(out, y3, gam, bet, lout.row_group, lout.lat_chunks, lout.L, ws, silu=True)
        torch.cuda.synchronize()
        return out, y, y3, lout, gam, bet

    out, y, y3, lout, gam, bet = run(list(range(len(sizes))))
    # same statistics up to summation order: outputs within one bf16 step of the 3-launch kernel
    assert (y.float() - y3.float()).abs().max().item() <= 3e-2
    ref = _pack([F.silu(F.group_norm(t[None], 32, gam.float().cpu(), bet.float().cpu(), 1e-5)[0])
                 for t in _unpack(out.float().cpu(), lout.sizes)])
    assert (y.float().cpu() - ref).abs().max().item() < 4e-2
    out2, y2, _, _, _, _ = run(list(range(len(sizes))))
    assert torch.equal(out, out2) and torch.equal(y, y2)                       # deterministic
    last = len(sizes) - 1
    _, ys, _, ls, _, _ = run([last])                                            # the last latent alone
    assert torch.equal(ys, y[lout.row_off[last]:])                              # batch invariant
