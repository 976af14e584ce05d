"""B200 SD3 MMDiT forward (C-ABI kernels) vs the fp32 CPU oracle on identical random-init
weights. Tolerance (bf16 kernels vs fp32 oracle): cosine >= 0.999 on the noise
prediction of every latent and max-abs error <= 6% of the output's max-abs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(cfg, spec, seed=1):
    g = torch.Generator().manual_seed(seed)
    hs = {r: torch.randn(n, cfg.in_channels, int(r) // 8, int(r) // 8, generator=g) for r, n in spec.items()}
    L = sum(spec.values())
    ehs = torch.randn(L, cfg.context_len, cfg.joint_attention_dim, generator=g)
    pooled = torch.randn(L, cfg.pooled_projection_dim, generator=g)
    t = torch.rand(L, generator=g) * 1000
    return hs, ehs, pooled, t


def _compare(out, ref):
    for r in ref:
        a, b = out[r].float().cpu(), ref[r]
        for i in range(a.shape[0]):
            cos = torch.nn.functional.cosine_similarity(a[i].flatten(), b[i].flatten(), dim=0).item()
            err = (a[i] - b[i]).abs().max().item() / b[i].abs().max().item()
            assert cos >= 0.999, (r, i, cos)
            assert err <= 0.06, (r, i, err)


def _build(cfg, seed=0):
    from oracle import sd3_mmdit as o3
    from sduss_b200.sd3_transformer import B200SD3Transformer2DModel
    sd = o3.init_sd3_weights(cfg, seed)
    # the oracle sees the same bf16-rounded weights the kernels use
    sd = {k: v.to(torch.bfloat16).float() for k, v in sd.items()}
    model = B200SD3Transformer2DModel(sd, cfg, device="cuda")
    return sd, model


@pytest.mark.parametrize("spec", [{"256": 2}, {"256": 1, "512": 2, "768": 1}])
def test_tiny_mmdit_matches_oracle(cuda, spec):
    from oracle import sd3_mmdit as o3
    cfg = o3.sd3_tiny_config()
    sd, model = _build(cfg)
    hs, ehs, pooled, t = _inputs(cfg, spec)
    q = lambda x: x.to(torch.bfloat16).float()
    ref = o3.sd3_forward(sd, cfg, {k: q(v) for k, v in hs.items()}, q(ehs), q(pooled), t)
    out = model({k: v.cuda().bfloat16() for k, v in hs.items()}, ehs.cuda().bfloat16(),
                pooled.cuda().bfloat16(), t.cuda(), is_sliced=True, patch_size=256)[0]
    assert list(out.keys()) == list(ref.keys())
    _compare(out, ref)


def test_batch_invariance(cuda):
    """Mixed-batch output of a latent == its single-latent output (what Mixfusion relies on)."""
    from oracle import sd3_mmdit as o3
    cfg = o3.sd3_tiny_config()
    _, model = _build(cfg)
    hs, ehs, pooled, t = _inputs(cfg, {"256": 2, "512": 1})
    dev = lambda x: x.cuda().bfloat16()
    full = model({k: dev(v) for k, v in hs.items()}, dev(ehs), dev(pooled), t.cuda())[0]
    solo = model({"512": dev(hs["512"])}, dev(ehs[2:3]), dev(pooled[2:3]), t[2:3].cuda())[0]
    assert torch.equal(full["512"], solo["512"])


def test_medium_width_one_layer_stack(cuda):
    """Full SD3.5-medium width (24 heads, 1536, 4096-d context, 333 tokens), 2 layers."""
    from oracle import sd3_mmdit as o3
    cfg = o3.sd35_medium_config()
    cfg.num_layers = 2
    cfg.dual_attention_layers = [0]
    sd, model = _build(cfg)
    hs, ehs, pooled, t = _inputs(cfg, {"256": 1, "512": 1})
    q = lambda x: x.to(torch.bfloat16).float()
    ref = o3.sd3_forward(sd, cfg, {k: q(v) for k, v in hs.items()}, q(ehs), q(pooled), t)
    out = model({k: v.cuda().bfloat16() for k, v in hs.items()}, ehs.cuda().bfloat16(),
                pooled.cuda().bfloat16(), t.cuda())[0]
    _compare(out, ref)
