"""Packed varlen tcgen05 attention vs PyTorch fp32 softmax attention, per sequence."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
H = 3


def _rand(shape, dev, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g).to(dev).to(torch.bfloat16)


def _ref(q, k, v, scale):
    # q: [Sq, H, 64] etc. fp32 reference
    q, k, v = (t.float().transpose(0, 1) for t in (q, k, v))
    p = torch.softmax(q @ k.transpose(1, 2) * scale, dim=-1)
    return (p @ v).transpose(0, 1)


def _check(out, ref):
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
    assert cos > 0.9995, cos


# max_ctas: None = one CTA per SM (these small cases then run one unit per CTA); 1 / 3 / 7 force the
# persistent path: several units per CTA, with idle query tiles, partial tiles and 1-step units
# in between full ones.
@pytest.mark.parametrize("max_ctas", [None, 1, 3, 7])
@pytest.mark.parametrize("img_lens,ctx_len", [((128,), 0), ((256, 1024), 0), ((1024, 256, 2304), 333),
                                              ((256,), 45), ((200, 77), 50), ((40, 700, 64, 1), 0)])
def test_joint_and_self(cuda, img_lens, ctx_len, max_ctas):
    from sduss_b200 import ops
    L = len(img_lens)
    Ta, Tb = sum(img_lens), L * ctx_len
    C = H * 64
    qkv_a = _rand((Ta, 3 * C), cuda, 1)
    out_a = torch.zeros(Ta, C, device=cuda, dtype=torch.bfloat16)
    if ctx_len:
        qkv_b = _rand((Tb, 3 * C), cuda, 2)
        out_b = torch.zeros(Tb, C, device=cuda, dtype=torch.bfloat16)
    seqs, ra = [], 0
    for i, s in enumerate(img_lens):
        seqs.append((ra, s, i * ctx_len, ctx_len, ra, s, i * ctx_len, ctx_len))
        ra += s
    plan = ops.build_attn_plan(seqs, cuda, H, max_ctas)
    sa = ops.attn_source(q=qkv_a, q_col=0, k=qkv_a, k_col=C, v=qkv_a, v_col=2 * C, out=out_a)
    sb = ops.attn_source(q=qkv_b, q_col=0, k=qkv_b, k_col=C, v=qkv_b, v_col=2 * C, out=out_b) if ctx_len else None
    scale = 1 / math.sqrt(64)
    ops.attn_varlen(sa, sb, *plan, scale)
    torch.cuda.synchronize()
    ra = 0
    for i, s in enumerate(img_lens):
        x = qkv_a[ra:ra + s].view(s, 3, H, 64)
        if ctx_len:
            y = qkv_b[i * ctx_len:(i + 1) * ctx_len].view(ctx_len, 3, H, 64)
            x = torch.cat([x, y], 0)
        ref = _ref(x[:, 0], x[:, 1], x[:, 2], scale).reshape(-1, C)
        _check(out_a[ra:ra + s], ref[:s])
        if ctx_len:
            _check(out_b[i * ctx_len:(i + 1) * ctx_len], ref[s:])
        ra += s


@pytest.mark.parametrize("max_ctas", [None, 4])
def test_cross(cuda, max_ctas):
    """SDXL cross attention: Q = image tokens (A), K/V = 77 text tokens (B)."""
    from sduss_b200 import ops
    img_lens, T = (1024, 4096, 256), 77
    C = H * 64
    q = _rand((sum(img_lens), C), cuda, 3)
    kv = _rand((len(img_lens) * T, 2 * C), cuda, 4)
    out = torch.zeros_like(q)
    seqs, ra = [], 0
    for i, s in enumerate(img_lens):
        seqs.append((ra, s, 0, 0, 0, 0, i * T, T))
        ra += s
    plan = ops.build_attn_plan(seqs, cuda, H, max_ctas)
    sa = ops.attn_source(q=q, out=out)
    sb = ops.attn_source(k=kv, k_col=0, v=kv, v_col=C)
    ops.attn_varlen(sa, sb, *plan, 0.125)
    torch.cuda.synchronize()
    ra = 0
    for i, s in enumerate(img_lens):
        kk = kv[i * T:(i + 1) * T].view(T, 2, H, 64)
        ref = _ref(q[ra:ra + s].view(s, H, 64), kk[:, 0], kk[:, 1], 0.125).reshape(s, C)
        _check(out[ra:ra + s], ref)
        ra += s


def _joint(cuda, qkv_a, qkv_b, img_lens, ctx_len, max_ctas=None):
    """Runs the packed joint attention for sequences laid out back to back; returns (out_a, out_b)."""
    from sduss_b200 import ops
    C = H * 64
    out_a = torch.zeros(qkv_a.shape[0], C, device=cuda, dtype=torch.bfloat16)
    out_b = torch.zeros(qkv_b.shape[0], C, device=cuda, dtype=torch.bfloat16)
    seqs, ra = [], 0
    for i, s in enumerate(img_lens):
        seqs.append((ra, s, i * ctx_len, ctx_len, ra, s, i * ctx_len, ctx_len))
        ra += s
    plan = ops.build_attn_plan(seqs, cuda, H, max_ctas)
    sa = ops.attn_source(q=qkv_a, q_col=0, k=qkv_a, k_col=C, v=qkv_a, v_col=2 * C, out=out_a)
    sb = ops.attn_source(q=qkv_b, q_col=0, k=qkv_b, k_col=C, v=qkv_b, v_col=2 * C, out=out_b)
    ops.attn_varlen(sa, sb, *plan, 0.125)
    torch.cuda.synchronize()
    return out_a, out_b


def test_deterministic_and_batch_invariant(cuda):
    """Same inputs -> same bits on every launch, and a sequence's output does not depend on the
    sequences packed around it (the property Mixfusion's mixed batches rely on)."""
    img_lens, ctx = (1024, 200, 2304, 256), 333
    C = H * 64
    qkv_a = _rand((sum(img_lens), 3 * C), cuda, 11)
    qkv_b = _rand((len(img_lens) * ctx, 3 * C), cuda, 12)
    oa, ob = _joint(cuda, qkv_a, qkv_b, img_lens, ctx)
    for _ in range(5):
        oa2, ob2 = _joint(cuda, qkv_a, qkv_b, img_lens, ctx)
        assert torch.equal(oa, oa2) and torch.equal(ob, ob2)
    ra = 0
    for i, s in enumerate(img_lens):
        a1, b1 = _joint(cuda, qkv_a[ra:ra + s].contiguous(), qkv_b[i * ctx:(i + 1) * ctx].contiguous(), (s,), ctx)
        assert torch.equal(a1, oa[ra:ra + s]), i
        assert torch.equal(b1, ob[i * ctx:(i + 1) * ctx]), i
        ra += s
    # ... nor on how the units are spread over the CTAs (a unit's arithmetic is self-contained)
    for m in (1, 5, 100000):
        oa3, ob3 = _joint(cuda, qkv_a, qkv_b, img_lens, ctx, max_ctas=m)
        assert torch.equal(oa, oa3) and torch.equal(ob, ob3), m


@pytest.mark.parametrize("T", [77, 13, 80])
def test_cross_short_keys(cuda, T):
    """b200_attn_cross_short_bf16 (SDXL cross attention, <= 80 keys per sequence) vs fp32 softmax
    attention and vs the persistent kernel on the same inputs; ragged query lengths (partial 128-row
    tiles, a partial 16-row warp tile), K / V and Q / out living at column offsets of wider buffers;
    the query-tile mask of the patch cache."""
    from sduss_b200 import ops
    img_lens = (1024, 4096, 256, 328)
    C = H * 64
    qbuf = _rand((sum(img_lens), C + 64), cuda, 3)            # q at column 64 of a wider buffer
    kv = _rand((len(img_lens) * T, 2 * C + 128), cuda, 4)     # k at column 128, v right after
    q = qbuf[:, 64:]
    seqs, ra = [], 0
    for i, s in enumerate(img_lens):
        seqs.append((ra, s, 0, 0, 0, 0, i * T, T))
        ra += s
    plan = ops.build_attn_plan(seqs, cuda, H)
    sb = ops.attn_source(k=kv, k_col=128, v=kv, v_col=128 + C)
    out = torch.zeros(sum(img_lens), C + 64, device=cuda, dtype=torch.bfloat16)
    sa = ops.attn_source(q=qbuf, q_col=64, out=out, o_col=64)
    ops.attn_cross_short(sa, sb, plan[0], len(img_lens), H, max(img_lens), T, 0.125)
    big = torch.zeros_like(out)
    ops.attn_varlen(ops.attn_source(q=qbuf, q_col=64, out=big, o_col=64), sb, *plan, 0.125)
    torch.cuda.synchronize()
    assert (out[:, :64] == 0).all()                            # nothing written outside the head columns
    ra = 0
    for i, s in enumerate(img_lens):
        kk = kv[i * T:(i + 1) * T, 128:].reshape(T, 2, H, 64)
        ref = _ref(q[ra:ra + s].reshape(s, H, 64), kk[:, 0], kk[:, 1], 0.125).reshape(s, C)
        _check(out[ra:ra + s, 64:], ref)
        ra += s
    d = (out.float() - big.float()).abs().max().item()
    assert d <= 2e-2 * big.float().abs().max().item(), d      # the two kernels agree to bf16 rounding
    # deterministic, and a sequence's result does not depend on its neighbours
    out2 = torch.zeros_like(out)
    ops.attn_cross_short(ops.attn_source(q=qbuf, q_col=64, out=out2, o_col=64), sb, plan[0], len(img_lens), H,
                         max(img_lens), T, 0.125)
    solo = torch.zeros_like(out)
    one = ops.build_attn_plan([seqs[1]], cuda, H)
    ops.attn_cross_short(ops.attn_source(q=qbuf, q_col=64, out=solo, o_col=64), sb, one[0], 1, H, img_lens[1], T, 0.125)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    assert torch.equal(solo[1024:1024 + 4096], out[1024:1024 + 4096])
    # query-tile mask (256-row chunks of the packed rows): clean chunks keep what the buffer held
    n_chunks = (sum(img_lens) + 255) // 256
    mask = (torch.arange(n_chunks) % 3 != 1).int().cuda()
    m = torch.full_like(out, 7.0)
    ops.attn_cross_short(ops.attn_source(q=qbuf, q_col=64, out=m, o_col=64), sb, plan[0], len(img_lens), H,
                         max(img_lens), T, 0.125, q_mask=mask)
    torch.cuda.synchronize()
    rows = torch.zeros(sum(img_lens), dtype=torch.bool, device=cuda)
    ra = 0
    for s in img_lens:                                         # a tile is looked up by its first row
        for t0 in range(0, s, 128):
            if mask[(ra + t0) >> 8] != 0:
                rows[ra + t0:ra + min(t0 + 128, s)] = True
        ra += s
    assert torch.equal(m[rows][:, 64:], out[rows][:, 64:]) and (m[~rows] == 7.0).all()


def test_joint_bounded_logits_matches_the_running_max_path(cuda):
    """B200AttnExtra.bounded_logits (SD3.5: q, k RMS-normalised, |logit * scale * log2 e| <= 64): the
    softmax without a reference maximum gives the same attention -- against fp32 and against the default
    instantiation on the same inputs, with rows whose logits are all far below zero (no maximum to
    subtract: their P values are tiny, their normalised output must still be exact to bf16)."""
    from sduss_b200 import ops
    C = H * 64
    img_lens, ctx = (1024, 2304, 333), 77
    g = torch.Generator().manual_seed(5)

    def unit(n, cols):  # rows of norm 8 per head, like an RMS-normalised head
        t = torch.randn(n, cols // 64, 64, generator=g)
        return (t / t.norm(dim=-1, keepdim=True) * 8).reshape(n, cols)
    def pack(n):
        q, k = unit(n, C) * 1.5, unit(n, C) * 1.5           # |logit| * 0.125 * 1.4427 <= 26
        return torch.cat([q, k, torch.randn(n, C, generator=g)], 1)
    qkv_a, qkv_b = pack(sum(img_lens)), pack(len(img_lens) * ctx)
    qkv_a[:64, :C] = -qkv_a[100:101, C:2 * C] * 1.0         # queries anti-aligned with one key: very negative row maxima elsewhere
    qkv_a, qkv_b = qkv_a.cuda().bfloat16(), qkv_b.cuda().bfloat16()
    seqs, ra = [], 0
    for i, s in enumerate(img_lens):
        seqs.append((ra, s, i * ctx, ctx, ra, s, i * ctx, ctx))
        ra += s
    plan = ops.build_attn_plan(seqs, cuda, H)
    outs = {}
    for bounded in (False, True):
        oa = torch.zeros(qkv_a.shape[0], C, device=cuda, dtype=torch.bfloat16)
        ob = torch.zeros(qkv_b.shape[0], C, device=cuda, dtype=torch.bfloat16)
        sa = ops.attn_source(q=qkv_a, q_col=0, k=qkv_a, k_col=C, v=qkv_a, v_col=2 * C, out=oa)
        sb = ops.attn_source(q=qkv_b, q_col=0, k=qkv_b, k_col=C, v=qkv_b, v_col=2 * C, out=ob)
        ops.attn_varlen(sa, sb, *plan, 0.125, bounded=bounded)
        torch.cuda.synchronize()
        outs[bounded] = (oa, ob)
    ra = 0
    for i, s in enumerate(img_lens):
        x = torch.cat([qkv_a[ra:ra + s], qkv_b[i * ctx:(i + 1) * ctx]], 0).view(s + ctx, 3, H, 64)
        ref = _ref(x[:, 0], x[:, 1], x[:, 2], 0.125).reshape(-1, C)
        for bounded in (False, True):
            _check(outs[bounded][0][ra:ra + s], ref[:s])
            _check(outs[bounded][1][i * ctx:(i + 1) * ctx], ref[s:])
        ra += s
    assert torch.isfinite(outs[True][0].float()).all() and torch.isfinite(outs[True][1].float()).all()
    d = (outs[True][0].float() - outs[False][0].float()).abs().max().item()
    assert d < 2e-2, d
