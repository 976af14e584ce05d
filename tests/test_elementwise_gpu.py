"""HBM-bound kernels vs oracle / golden fixtures. Index and scheduler work is bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_layernorm_mod(cuda):
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(0)
    for T, D in ((100, 1536), (37, 640), (64, 1280), (9, 128)):
        x = (torch.randn(T, D, generator=g) * 3 + 1).cuda().bfloat16()
        mod = torch.randn(4, 6 * D, generator=g).cuda().bfloat16()
        grp = (torch.arange(T) % 4).int().cuda()
        y, y2 = torch.empty_like(x), torch.empty_like(x)
        ops.layernorm_mod(x, y, eps=1e-6, mod=mod, row_group=grp, shift_col=0, scale_col=D,
                          y2=y2, shift2_col=2 * D, scale2_col=3 * D)
        n = torch.nn.functional.layer_norm(x.float(), (D,), eps=1e-6)
        m = mod.float()[grp.long()]
        assert (y.float() - (n * (1 + m[:, D:2 * D]) + m[:, :D])).abs().max() < 0.06
        assert (y2.float() - (n * (1 + m[:, 3 * D:4 * D]) + m[:, 2 * D:3 * D])).abs().max() < 0.06
        gam, bet = torch.randn(D, generator=g).cuda().bfloat16(), torch.randn(D, generator=g).cuda().bfloat16()
        ops.layernorm_mod(x, y, eps=1e-5, gamma=gam, beta=bet)
        ref = torch.nn.functional.layer_norm(x.float(), (D,), gam.float(), bet.float(), 1e-5)
        assert (y.float() - ref).abs().max() < 0.06


def test_layernorm_row_result_independent_of_packing(cuda):
    """A row's output must not depend on which rows / requests surround it: the kernel stages the
    folded parameters of two requests per block in shared memory and folds them on the fly for
    rows of any other request (slow path) - both must round identically (bit-exact)."""
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(1)
    for T, D in ((512, 1536), (512, 640)):
        x = (torch.randn(T, D, generator=g) * 2 + 0.5).cuda().bfloat16()
        mod = torch.randn(4, 6 * D, generator=g).cuda().bfloat16()
        grp_sorted = (torch.arange(T) * 4 // T).int()
        kw = dict(eps=1e-6, mod=mod, shift_col=0, scale_col=D, shift2_col=2 * D, scale2_col=3 * D)
        y, y2 = torch.empty_like(x), torch.empty_like(x)
        ops.layernorm_mod(x, y, row_group=grp_sorted.cuda(), y2=y2, **kw)
        # interleave the requests row by row: every block now sees all four requests
        perm = torch.arange(T).view(4, T // 4).t().reshape(-1)
        xp = x[perm.cuda()].contiguous()
        yp, yp2 = torch.empty_like(x), torch.empty_like(x)
        ops.layernorm_mod(xp, yp, row_group=grp_sorted[perm].contiguous().cuda(), y2=yp2, **kw)
        assert torch.equal(yp, y[perm.cuda()]) and torch.equal(yp2, y2[perm.cuda()])
        # and a single request on its own
        rows = slice(T // 4, T // 2)
        ys = torch.empty_like(x[rows])
        ops.layernorm_mod(x[rows].contiguous(), ys, row_group=torch.ones(T // 4, dtype=torch.int32).cuda(),
                          eps=1e-6, mod=mod, shift_col=0, scale_col=D)
        assert torch.equal(ys, y[rows])


def test_timestep_embedding_and_silu(cuda):
    from oracle.sd3_mmdit import timestep_embedding
    from sduss_b200 import ops
    t = torch.tensor([999.0, 981.0, 1.0, 0.0, 512.5, 1024.0], device=cuda)
    for dim in (256, 320):
        out = ops.timestep_embedding(t, dim)
        ref = timestep_embedding(t.cpu(), dim)
        assert (out.float().cpu() - ref).abs().max() < 1e-2  # bf16 output rounding + fp32 sincos
    x = torch.randn(16, 1536, device=cuda).bfloat16()
    assert (ops.silu(x).float() - torch.nn.functional.silu(x.float())).abs().max() < 2e-2


def test_sd3_pack_scatter_bit_exact(cuda):
    """patchify rows == the reference's split_sample_sd3 row order (golden) and
    unpatchify(patchify(x)) == x; values are small integers, exact in bf16."""
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(0)
    C, p = 16, 2
    spec = [("512", 2, 64), ("768", 1, 96), ("1024", 1, 128)]
    lat = {r: torch.randint(-120, 120, (n, C, s, s), generator=g).float() for r, n, s in spec}
    dev_lat = {r: v.cuda().bfloat16() for r, v in lat.items()}
    ptrs, desc, off = [], [], 0
    for r, n, s in spec:
        for i in range(n):
            ptrs.append(dev_lat[r][i].data_ptr())
            desc.append((off, s // p, s // p, 0))
            off += (s // p) ** 2
    T = off
    tokens = torch.zeros(T, C * p * p, device=cuda, dtype=torch.bfloat16)
    in_ptr = torch.tensor(ptrs, dtype=torch.int64).cuda()
    d = torch.tensor(desc, dtype=torch.int32).cuda()
    ops.sd3_patchify(in_ptr, d, len(ptrs), (128 // p) ** 2, C, p, tokens)
    # oracle: conv im2col order (c, py, px), tokens row-major over (h, w), latents back to back
    ref = torch.cat([torch.nn.functional.unfold(v, kernel_size=p, stride=p).transpose(1, 2).reshape(-1, C * p * p)
                     for v in lat.values()])
    assert torch.equal(tokens.float().cpu(), ref)
    # reference chunking (split_sample_sd3) of these token rows is a pure reshape of the buffer
    from oracle import pack as opack
    per_res = {}
    o = 0
    for r, n, s in spec:
        S = (s // p) ** 2
        per_res[r] = ref[o:o + n * S].reshape(n, S, -1).numpy()
        o += n * S
    chunks, lat_off, _ = opack.split_sample_sd3(per_res)
    assert np.array_equal(chunks.reshape(T, -1), tokens.float().cpu().numpy())
    # scatter: unpatchify layout is (py, px, c)
    tok2 = torch.randint(-120, 120, (T, p * p * C), generator=g).float()
    outs = {r: torch.zeros_like(v) for r, v in dev_lat.items()}
    optr = torch.tensor([outs[r][i].data_ptr() for r, n, s in spec for i in range(n)], dtype=torch.int64).cuda()
    ops.sd3_unpatchify(tok2.cuda().bfloat16(), d, len(ptrs), (128 // p) ** 2, C, p, optr)
    o = 0
    for r, n, s in spec:
        h = s // p
        x = tok2[o:o + n * h * h].reshape(n, h, h, p, p, C)
        x = torch.einsum("nhwpqc->nchpwq", x).reshape(n, C, s, s)
        assert torch.equal(outs[r].float().cpu(), x)
        o += n * h * h


def _step_refs(ops, x, out, sig, eps_rows=None):
    """One LatentRef per row of x (requests back to back); eps_rows[r] = (uncond row, cond row)."""
    R, n = x.shape[0], x[0].numel()
    xs = x.element_size()
    rows = []
    for r in range(R):
        u, c = (-1, r) if eps_rows is None else eps_rows[r]
        rows.append((x.data_ptr() + r * n * xs, out.data_ptr() + r * n * xs, n,
                     u * n if u >= 0 else -1, c * n, float(sig[r][0]), float(sig[r][1])))
    return ops.latent_refs(rows)


@pytest.mark.parametrize("dt", ["bf16", "f32"])
def test_flow_match_step_matches_reference_golden(cuda, dt):
    """FlowMatchEulerDiscreteScheduler.batch_step run by the reference's own code
    (tools/make_golden.py): bit-exact, in bf16 and in fp32 (samples and model outputs)."""
    from oracle import schedulers as osch
    from sduss_b200 import ops
    z = np.load(os.path.join(G, "sched_flow_match.npz"))
    tdt = torch.bfloat16 if dt == "bf16" else torch.float32
    steps, idx = z[dt + "_steps"], z[dt + "_idx"]
    x = torch.from_numpy(z[dt + "_x"]).cuda().to(tdt)
    v = torch.from_numpy(z[dt + "_v"]).cuda().to(tdt)
    tabs = [osch.flow_match_sigmas(int(s))[0] for s in steps]
    sig = [(float(t[i]), float(t[i + 1])) for t, i in zip(tabs, idx)]
    out = torch.empty_like(x)
    ops.cfg_scheduler_step(v, _step_refs(ops, x, out, sig), tdt, 1.0, False, 0)
    assert torch.equal(out.float().cpu(), torch.from_numpy(z[dt + "_prev"]))


@pytest.mark.parametrize("dt", ["bf16", "f32"])
def test_euler_step_and_scale_match_reference_golden(cuda, dt):
    from oracle import schedulers as osch
    from sduss_b200 import ops
    z = np.load(os.path.join(G, "sched_euler.npz"))
    tdt = torch.bfloat16 if dt == "bf16" else torch.float32
    for mode, pt in ((1, "epsilon"), (2, "v_prediction")):
        tag = f"{pt}_{dt}"
        steps, idx = z[tag + "_steps"], z[tag + "_idx"]
        x = torch.from_numpy(z[tag + "_x"]).cuda().to(tdt)
        e = torch.from_numpy(z[tag + "_eps"]).cuda().to(tdt)
        R, n = x.shape[0], x[0].numel()
        tabs = [osch.euler_sigmas(int(s))[0] for s in steps]
        sig = [(float(t[i]), float(t[i + 1])) for t, i in zip(tabs, idx)]
        out = torch.empty_like(x)
        ops.cfg_scheduler_step(e, _step_refs(ops, x, out, sig), tdt, 1.0, False, mode)
        assert torch.equal(out.float().cpu(), torch.from_numpy(z[tag + "_prev"])), pt
        # scale_model_input on the CFG-duplicated batch [x, x]: one gather with a CFG duplicate.
        # The kernel's output is the model's bf16 input; the reference result is in the sample
        # dtype, so for fp32 samples the comparison is against its bf16 rounding.
        y = torch.empty((2 * R, n), device="cuda", dtype=torch.bfloat16)
        xs = x.element_size()
        refs = ops.latent_refs((x.data_ptr() + r * n * xs, 0, n, r * n, (R + r) * n, sig[r][0], 0.0)
                               for r in range(R))
        ops.gather_latents(refs, tdt, y.view(-1), scale_input=True)
        want = torch.from_numpy(z[tag + "_scaled"]).reshape(2 * R, n).to(torch.bfloat16)
        assert torch.equal(y.cpu(), want), pt


def test_scheduler_mixins_match_reference_golden(cuda):
    """The drop-in scheduler methods (same names / arguments / side effects as the reference's
    batch_step and batch_scale_model_input) on the same fixtures, through the public classes."""
    from types import SimpleNamespace
    from oracle import schedulers as osch
    from sduss_b200.schedulers import (B200EulerDiscreteScheduler, B200FlowMatchEulerDiscreteScheduler,
                                       SchedulerStates)
    z = np.load(os.path.join(G, "sched_euler.npz"))
    for pt in ("epsilon", "v_prediction"):
        for dt, tdt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
            tag = f"{pt}_{dt}"
            sch = B200EulerDiscreteScheduler(prediction_type=pt)
            reqs = []
            for n_steps, i in zip(z[tag + "_steps"], z[tag + "_idx"]):
                sig, ts, _ = osch.euler_sigmas(int(n_steps))
                st = SchedulerStates(sig, int(n_steps), ts)
                st._step_index = st.timestep_idx = int(i)
                reqs.append(SimpleNamespace(scheduler_states=st))
            x = torch.from_numpy(z[tag + "_x"]).cuda().to(tdt)
            e = torch.from_numpy(z[tag + "_eps"]).cuda().to(tdt)
            scaled = sch.batch_scale_model_input(reqs, torch.cat([x, x]), None)
            assert scaled.dtype == tdt
            assert torch.equal(scaled.to(torch.bfloat16).cpu(),
                               torch.from_numpy(z[tag + "_scaled"]).to(torch.bfloat16))
            prev = sch.batch_step(reqs, e, None, x, return_dict=False)
            assert prev.dtype == tdt and torch.equal(prev.float().cpu(), torch.from_numpy(z[tag + "_prev"]))
            assert [r.scheduler_states._step_index for r in reqs] == [int(i) + 1 for i in z[tag + "_idx"]]
            with pytest.raises(NotImplementedError):
                sch.batch_step(reqs, e, None, x, s_churn=0.1)
    z = np.load(os.path.join(G, "sched_flow_match.npz"))
    sch = B200FlowMatchEulerDiscreteScheduler()
    for dt, tdt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
        reqs = []
        for n_steps, i in zip(z[dt + "_steps"], z[dt + "_idx"]):
            sig, ts = osch.flow_match_sigmas(int(n_steps))
            st = SchedulerStates(sig, int(n_steps), ts)
            st._step_index = int(i)
            reqs.append(SimpleNamespace(scheduler_states=st))
        x = torch.from_numpy(z[dt + "_x"]).cuda().to(tdt)
        v = torch.from_numpy(z[dt + "_v"]).cuda().to(tdt)
        prev = sch.batch_step(reqs, v, x, None, return_dict=False)
        assert torch.equal(prev.float().cpu(), torch.from_numpy(z[dt + "_prev"]))


def test_cfg_combine_matches_torch_bf16(cuda):
    from oracle import schedulers as osch
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(3)
    R, n = 3, 16 * 64 * 64
    eps = torch.randn(2 * R, n, generator=g).cuda().bfloat16()
    x = torch.randn(R, n, generator=g).cuda().bfloat16()
    sig = [(1.0, 0.9), (0.5, 0.45), (0.1, 0.0)]
    out = torch.empty_like(x)
    refs = _step_refs(ops, x, out, sig, eps_rows=[(r, R + r) for r in range(R)])
    ops.cfg_scheduler_step(eps, refs, torch.bfloat16, 7.0, True, 0)
    comb = osch.cfg_combine(eps.cpu(), 7.0)  # bf16 tensor ops, as the reference pipeline does
    s = torch.tensor(sig)
    ref = osch.flow_match_batch_step(comb, x.cpu(), s[:, 0], s[:, 1])
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("ldt", [torch.float32, torch.float16])
def test_step_keeps_the_latent_dtype(cuda, ldt):
    """fp32 / fp16 latents with a bf16 model: the update runs in fp32 on the upcast sample exactly
    as the reference does and the result stays in the LATENT's dtype (the reference would cast it to
    the model dtype; keeping it loses nothing between steps). 40 requests: two launches of 32."""
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(5)
    R, n = 40, 1000
    eps = torch.randn(2 * R, n, generator=g).cuda().bfloat16()
    x = torch.randn(R, n, generator=g).cuda().to(ldt)
    sig = [(1.0 - 0.01 * r, 0.98 - 0.01 * r) for r in range(R)]
    out = torch.empty_like(x)
    refs = _step_refs(ops, x, out, sig, eps_rows=[(r, R + r) for r in range(R)])
    ops.cfg_scheduler_step(eps, refs, ldt, 5.0, True, 1)
    u, c = eps[:R], eps[R:]
    comb = (u + 5.0 * (c - u)).float()
    s = torch.tensor(sig, device="cuda")
    xf = x.float()
    x0 = xf - s[:, :1] * comb
    d = (xf - x0) / s[:, :1]
    want = (xf + d * (s[:, 1:] - s[:, :1])).to(ldt)
    assert out.dtype == ldt and torch.equal(out, want)


def test_write_f32_and_gather_rows(cuda):
    from sduss_b200 import ops
    vals = [float(i) * 0.37 - 3 for i in range(300)]  # > 256: two launches
    dst = torch.zeros(320, device="cuda")
    ops.write_f32(dst, vals)
    assert torch.equal(dst[:300].cpu(), torch.tensor(vals, dtype=torch.float32))
    assert dst[300:].abs().sum().item() == 0
    g = torch.Generator().manual_seed(1)
    for cols, dtype in ((1536 * 5, torch.bfloat16), (6, torch.float32), (130, torch.bfloat16)):
        srcs = [torch.randn(cols, generator=g).to(dtype).cuda() for _ in range(70)]  # > 64
        out = torch.zeros((70, cols + 8), device="cuda", dtype=dtype)
        ops.gather_rows(out, srcs)  # row stride > row bytes
        assert torch.equal(out[:, :cols], torch.stack(srcs))
        assert out[:, cols:].abs().sum().item() == 0
