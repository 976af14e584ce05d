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
