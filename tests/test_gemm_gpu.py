"""tcgen05 GEMM + fused epilogues vs a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(torch.bfloat16)


def _close(out, ref, tol=2e-2):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (200, 136, 192),
                                   (1357, 1536, 1536), (4096, 4608, 1536), (333, 64, 1536),
                                   (14848, 1536, 1536)])
def test_gemm_bias(cuda, M, N, K):
    from sduss_b200 import ops
    a, w, b = _rand((M, K), cuda, seed=1), _rand((N, K), cuda, 0.05, seed=2), _rand((N,), cuda, seed=3)
    out = ops.gemm(a, w, bias=b)
    ref = a.float() @ w.float().t() + b.float()
    _close(out, ref)
    out32 = ops.gemm(a, w, out_fp32=True)
    _close(out32, a.float() @ w.float().t(), tol=2e-3)


@pytest.mark.parametrize("M,N,K", [(40960, 320, 64), (40960, 64, 64), (30000, 192, 128), (40960, 448, 64)])
def test_gemm_many_short_tiles_with_n_tail(cuda, M, N, K):
    """Dozens of output tiles per CTA, one or two K blocks each, and an odd number of 64-column
    chunks per tile (N tail / N < BN). Regression: the TMA-store staging buffers alternated per
    tile, so such a tile was followed by a chunk staged into the buffer whose store was still in
    flight (SDXL conv_in, M = 40960, N = 320, K = 64, came out corrupted and non-deterministic)."""
    from sduss_b200 import ops
    a, w, b = _rand((M, K), cuda, seed=1), _rand((N, K), cuda, 0.2, seed=2), _rand((N,), cuda, seed=3)
    ref = a.float() @ w.float().t() + b.float()
    outs = [ops.gemm(a, w, bias=b) for _ in range(4)]
    for o in outs:
        _close(o, ref)
        assert torch.equal(o, outs[0])
    resid = _rand((M, N), cuda, seed=6)
    outs = [ops.gemm(a, w, bias=b, epi=ops.EPI_GATE_RESID, resid=resid) for _ in range(3)]
    for o in outs:
        _close(o, ref + resid.float())
        assert torch.equal(o, outs[0])


def test_gemm_strided_a(cuda):
    from sduss_b200 import ops
    big = _rand((512, 1024), cuda, seed=4)
    a = big[:, 256:256 + 384]
    w = _rand((192, 384), cuda, 0.05, seed=5)
    _close(ops.gemm(a, w), a.float() @ w.float().t())


def test_gemm_gelu_tanh(cuda):
    from sduss_b200 import ops
    a, w, b = _rand((777, 256), cuda, seed=1), _rand((512, 256), cuda, 0.1, seed=2), _rand((512,), cuda, seed=3)
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_GELU_TANH)
    ref = torch.nn.functional.gelu(a.float() @ w.float().t() + b.float(), approximate="tanh")
    _close(out, ref)


def test_gemm_gate_resid(cuda):
    from sduss_b200 import ops
    M, N, K, G = 1000, 384, 256, 5
    a, w, b = _rand((M, K), cuda, seed=1), _rand((N, K), cuda, 0.1, seed=2), _rand((N,), cuda, seed=3)
    resid = _rand((M, N), cuda, seed=6)
    gate = _rand((G, 3 * N), cuda, seed=7)
    grp = (torch.arange(M, device=cuda) * G // M).to(torch.int32)
    ref = resid.float() + gate[:, N:2 * N].float()[grp.long()] * (a.float() @ w.float().t() + b.float())
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_GATE_RESID, resid=resid, gate=gate[:, N:2 * N],
                   row_group=grp)
    _close(out, ref)
    # in-place residual, no gate
    x = resid.clone()
    ops.gemm(a, w, out=x, bias=b, epi=ops.EPI_GATE_RESID, resid=x)
    _close(x, resid.float() + a.float() @ w.float().t() + b.float())


def test_gemm_qk_rmsnorm(cuda):
    from sduss_b200 import ops
    M, K, H = 700, 256, 4
    N = 3 * H * 64
    a, w, b = _rand((M, K), cuda, seed=1), _rand((N, K), cuda, 0.1, seed=2), _rand((N,), cuda, seed=3)
    wq, wk = _rand((64,), cuda, seed=8) + 1, _rand((64,), cuda, seed=9) + 1
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_QK_RMSNORM, rms_wq=wq, rms_wk=wk,
                   rms_q_cols=H * 64, rms_k_cols=H * 64, rms_eps=1e-6, q_scale=0.25)
    y = (a.float() @ w.float().t() + b.float()).view(M, 3, H, 64)

    def rms(t, wt):
        return t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * wt.float()
    ref = torch.stack([rms(y[:, 0], wq) * 0.25, rms(y[:, 1], wk), y[:, 2]], 1).view(M, N)
    _close(out, ref)


def test_gemm_geglu(cuda):
    from sduss_b200 import ops
    M, K, F = 520, 128, 256  # F = hidden width; linear produces 2F
    a = _rand((M, K), cuda, seed=1)
    w = _rand((2 * F, K), cuda, 0.1, seed=2)
    b = _rand((2 * F,), cuda, seed=3)
    # interleave rows [32 hidden | 32 gate]
    idx = torch.arange(2 * F, device=cuda).view(2, F // 32, 32).permute(1, 0, 2).reshape(-1)
    out = ops.gemm(a, w[idx].contiguous(), bias=b[idx].contiguous(), epi=ops.EPI_GEGLU)
    y = a.float() @ w.float().t() + b.float()
    ref = y[:, :F] * torch.nn.functional.gelu(y[:, F:])
    _close(out, ref)


def test_gemm_rowvec(cuda):
    from sduss_b200 import ops
    M, N, K, G = 640, 320, 192, 3
    a, w, b = _rand((M, K), cuda, seed=1), _rand((N, K), cuda, 0.1, seed=2), _rand((N,), cuda, seed=3)
    rv = _rand((G, N), cuda, seed=5)
    grp = (torch.arange(M, device=cuda) % G).to(torch.int32)
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_ROWVEC, rowvec=rv, row_group=grp)
    _close(out, a.float() @ w.float().t() + b.float() + rv.float()[grp.long()])


def test_gemm_rejects_bad_args(cuda):
    from sduss_b200 import ops
    from sduss_b200._lib import B200Error
    a, w = _rand((64, 36), cuda), _rand((64, 36), cuda)
    with pytest.raises(B200Error):
        ops.gemm(a, w)


@pytest.mark.parametrize("M,N,K", [(2560, 1280, 1280), (2560, 1280, 5120), (1998, 1536, 1536), (2560, 1208, 320)])
def test_gemm_one_round_shapes_use_192_wide_tiles(cuda, M, N, K):
    """Problems whose 128 x 256 tiling leaves >= 1/4 of the SMs idle run as 128 x 192 tiles (SDXL
    level 2, SD3 context rows): bias, in-place residual (TMA-staged residual tiles), N tails that
    end inside a 64-column chunk, determinism."""
    from sduss_b200 import ops
    a, w, b = _rand((M, K), cuda, seed=1), _rand((N, K), cuda, 0.03, seed=2), _rand((N,), cuda, seed=3)
    ref = a.float() @ w.float().t() + b.float()
    outs = [ops.gemm(a, w, bias=b) for _ in range(3)]
    for o in outs:
        _close(o, ref)
        assert torch.equal(o, outs[0])
    resid = _rand((M, N), cuda, seed=6)
    x = resid.clone()
    ops.gemm(a, w, out=x, bias=b, epi=ops.EPI_GATE_RESID, resid=x)
    _close(x, ref + resid.float())
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_GELU_TANH)
    _close(out, torch.nn.functional.gelu(ref, approximate="tanh"))


@pytest.mark.parametrize("M,N,K,epi", [(2560, 3840, 1280, "bias"), (300, 1280, 640, "bias"), (2560, 10240, 1280, "geglu"),
                                       (777, 256, 128, "resid")])
def test_layernorm_folded_into_the_gemm(cuda, M, N, K, epi):
    """LN(x) W^T + b computed as rstd * (x (W o gamma)^T - mean * colsum) + (b + beta W^T): row
    statistics (b200_row_stats_bf16) + an epilogue correction instead of a LayerNorm pass. Checked
    against fp32 LayerNorm -> linear in torch, with inputs that have a large row mean (the
    cancellation case) and per-row scales."""
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(M, K, generator=g) * (0.5 + torch.rand(M, 1, generator=g) * 3) + torch.randn(M, 1, generator=g) * 4)
    x = x.cuda().bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    gamma = (1 + 0.2 * torch.randn(K, generator=g)).cuda().bfloat16()
    beta = (0.3 * torch.randn(K, generator=g)).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda().bfloat16()
    stats = torch.empty(2 * M, device=cuda)
    ops.row_stats(x, stats, 1e-5)
    xf = x.float()
    st = stats.view(M, 2)
    assert torch.allclose(st[:, 0], xf.mean(1), atol=1e-3, rtol=1e-4)
    assert torch.allclose(st[:, 1], torch.rsqrt(xf.var(1, unbiased=False) + 1e-5), rtol=1e-3)
    wq, cs, bq = ops.fold_layernorm(w, gamma, beta, bias)
    ln = torch.nn.functional.layer_norm(xf, (K,), gamma.float(), beta.float(), 1e-5)
    ref = ln @ w.float().t() + bias.float()
    kw = {}
    if epi == "geglu":
        kw["epi"] = ops.EPI_GEGLU
        ref = ref.view(M, N // 64, 2, 32)
        ref = (ref[:, :, 0] * torch.nn.functional.gelu(ref[:, :, 1])).reshape(M, N // 2)
    elif epi == "resid":
        r = torch.randn(M, N, generator=g).cuda().bfloat16()
        kw.update(epi=ops.EPI_GATE_RESID, resid=r)
        ref = ref + r.float()
    out = ops.gemm(x, wq, bias=bq, ln_stats=stats, ln_colsum=cs, **kw)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2.5e-2 * ref.abs().max().item() + 1e-2, (err, ref.abs().max().item())
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
    assert cos > 0.9995, cos
    # the same with the statistics taken from partial sums a PRODUCER GEMM left behind (no statistics
    # kernel): x2 = x0 @ P^T + resid written with rowpart_out, then consumed with ln_rowpart
    if K % 64 == 0:
        x0 = torch.randn(M, 128, generator=g).cuda().bfloat16()
        pw = (torch.randn(K, 128, generator=g) / 11).cuda().bfloat16()
        rs = (torch.randn(M, K, generator=g) * 2 + 3).cuda().bfloat16()
        part = torch.zeros(M * (K // 64) * 2, device=cuda)
        x2 = ops.gemm(x0, pw, epi=ops.EPI_GATE_RESID, resid=rs, rowpart_out=part)
        pv = part.view(K // 64, M, 2).transpose(0, 1)         # chunk-major: [K/64][M] (sum, sum of squares)
        x2f = x2.float()
        assert torch.allclose(pv[:, :, 0].sum(1) / K, x2f.mean(1), atol=2e-2)          # sums of the pre-rounding values
        ref2 = torch.nn.functional.layer_norm(x2f, (K,), gamma.float(), beta.float(), 1e-5) @ w.float().t() + bias.float()
        if epi == "geglu":
            ref2 = ref2.view(M, N // 64, 2, 32)
            ref2 = (ref2[:, :, 0] * torch.nn.functional.gelu(ref2[:, :, 1])).reshape(M, N // 2)
        elif epi == "resid":
            ref2 = ref2 + kw["resid"].float()
        out3 = ops.gemm(x2, wq, bias=bq, ln_rowpart=part, ln_colsum=cs, ln_eps=1e-5, **kw)
        torch.cuda.synchronize()
        err3 = (out3.float() - ref2).abs().max().item()
        assert err3 <= 2.5e-2 * ref2.abs().max().item() + 1e-2, (err3, ref2.abs().max().item())
        assert torch.nn.functional.cosine_similarity(out3.float().flatten(), ref2.flatten(), dim=0).item() > 0.9995
    # same accuracy class as the unfolded path (LayerNorm kernel -> GEMM)
    y = ops.layernorm_mod(x, torch.empty_like(x), eps=1e-5, gamma=gamma, beta=beta)
    out2 = ops.gemm(y, w, bias=bias, **kw)
    err2 = (out2.float() - ref).abs().max().item()
    assert err <= 3 * err2 + 2e-2, (err, err2)
