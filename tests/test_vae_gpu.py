"""VAE decode stage on the B200 kernels vs the fp32 CPU oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_softmax_rows(cuda):
    from sduss_b200 import ops
    g = torch.Generator().manual_seed(0)
    for rows, cols in ((64, 64), (37, 1024), (5, 16384), (8, 2308)):
        big = torch.randn(rows, cols + 4, generator=g) * 30
        s = big.cuda()[:, :cols]                       # row stride > cols
        p = torch.empty(rows, cols, device=cuda, dtype=torch.bfloat16)
        scale = 1.0 / math.sqrt(512)
        ops.softmax_rows(s, p, scale)
        ref = torch.softmax(big[:, :cols] * scale, dim=-1)
        assert (p.float().cpu() - ref).abs().max() < 4e-3 * ref.max().item() + 1e-6
        assert (p.float().sum(-1).cpu() - 1).abs().max() < 2e-2


@pytest.mark.parametrize("C,pq", [(4, True), (16, False)])
def test_latent_affine(cuda, C, pq):
    from sduss_b200 import ops
    from sduss_b200.layout import LevelLayout
    g = torch.Generator().manual_seed(1)
    sizes = [(8, 8), (16, 24), (8, 8)]
    lats = [torch.randn(C, h, w, generator=g).bfloat16() for h, w in sizes]
    W = torch.randn(C, C, generator=g) if pq else torch.eye(C) * 1.7
    b = torch.randn(C, generator=g)
    dl = [t.cuda().contiguous() for t in lats]
    outs = [torch.empty_like(t) for t in dl]
    lay = LevelLayout([(8, 8), (16, 24), (8, 8)], cuda)
    ptr = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64).cuda()
    ops.latent_affine(ptr(dl), ptr(outs), lay.desc, lay.L, lay.max_pixels, C, C, W.cuda(), b.cuda())
    for t, o in zip(lats, outs):
        ref = torch.einsum("oc,chw->ohw", W, t.float()) + b[:, None, None]
        assert (o.float().cpu() - ref).abs().max() <= 2 ** -7 * ref.abs().max() + 1e-6


def _decode_case(cuda, C, shift, pq, sizes):
    from oracle import vae_decoder as ov
    from sduss_b200.vae import B200VAEDecoder
    cfg = ov.vae_tiny_config(latent_channels=C, shift=shift, pq=pq)
    sd = {k: v.bfloat16().float() for k, v in ov.init_vae_decoder_weights(cfg, 0).items()}
    model = B200VAEDecoder(sd, cfg, device=cuda)
    g = torch.Generator().manual_seed(3)
    lat = {res: (torch.randn(n, C, h, w, generator=g) * 0.5).bfloat16() for res, (n, h, w) in sizes.items()}
    out = model.decode({k: v.cuda() for k, v in lat.items()})
    torch.cuda.synchronize()
    ref = ov.vae_decode(sd, cfg, {k: v.float() for k, v in lat.items()})
    return model, lat, out, ref


@pytest.mark.parametrize("C,shift,pq", [(4, None, True), (16, 0.0609, False)])
def test_vae_decode_matches_oracle(cuda, C, shift, pq):
    """Mixed-resolution batch (8x8, 16x16 and 16x8 latents) in one pass vs per-image fp32 oracle.
    Tolerance (bf16 activations, fp32 accumulation, ~25 layers): cosine >= 0.999 per image and
    max-abs <= 6 % of the output range -- the bar of the denoising-step parity tests."""
    sizes = {"64": (2, 8, 8), "128": (1, 16, 16), "96": (1, 16, 8)}
    model, lat, out, ref = _decode_case(cuda, C, shift, pq, sizes)
    for res in sizes:
        assert out[res].shape == ref[res].shape
        for i in range(ref[res].shape[0]):
            a, b = out[res][i].float().cpu().flatten(), ref[res][i].flatten()
            cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
            assert cos >= 0.999, (res, i, cos)
            assert (a - b).abs().max() <= 0.06 * (b.max() - b.min()), (res, i)


def test_vae_decode_batch_invariant_and_replay(cuda):
    """An image must not depend on what it is batched with (bit-exact), and the CUDA-graph replay
    of a plan must reproduce the eager first call."""
    sizes = {"64": (2, 8, 8), "128": (1, 16, 16)}
    model, lat, out, _ = _decode_case(cuda, 4, None, True, sizes)
    again = model.decode({k: v.cuda() for k, v in lat.items()})      # captured
    third = model.decode({k: v.cuda() for k, v in lat.items()})      # replayed
    for res in sizes:
        assert torch.equal(out[res], again[res]) and torch.equal(out[res], third[res])
    solo = model.decode({"128": lat["128"].cuda()})
    assert torch.equal(solo["128"], out["128"])
    solo = model.decode({"64": lat["64"][1:2].cuda()})
    assert torch.equal(solo["64"][0], out["64"][1])


def test_post_inference_sets_outputs(cuda):
    from types import SimpleNamespace as NS
    from oracle import vae_decoder as ov
    from sduss_b200.pipelines import B200StableDiffusionXLPipeline
    from sduss_b200.vae import B200VAEDecoder
    cfg = ov.vae_tiny_config()
    sd = ov.init_vae_decoder_weights(cfg, 0)
    pipe = B200StableDiffusionXLPipeline(None, None, vae=B200VAEDecoder(sd, cfg, device=cuda))
    g = torch.Generator().manual_seed(5)
    mk = lambda i, h: NS(request_id=i, output=None, sampling_params=NS(
        latents=(torch.randn(1, 4, h, h, generator=g) * 0.5).bfloat16().cuda()))
    reqs = {"128": [mk(0, 16)], "64": [mk(1, 8), mk(2, 8)]}
    pipe.post_inference(reqs, output_type="pt")
    ref = ov.vae_decode({k: v.bfloat16().float() for k, v in sd.items()}, cfg,
                        {r: torch.cat([q.sampling_params.latents.float().cpu() for q in rs]) for r, rs in reqs.items()})
    for res, rs in reqs.items():
        for i, r in enumerate(rs):
            img = r.output.images
            assert img.shape == (3, 8 * int(res) // 8, 8 * int(res) // 8) and 0 <= img.min() and img.max() <= 1
            assert (img.cpu() - ov.postprocess(ref[res][i])).abs().max() < 0.06


def test_vae_proxy_decodes_unscaled_latents_like_vae_decode(cuda):
    """B200VAEProxy.decode(z) == reference `vae.decode(z)` semantics: z is already un-scaled
    (xl_esymred.py:441 / 3_esymred.py:408 do that before calling decode)."""
    from types import SimpleNamespace
    from oracle import vae_decoder as ov
    from sduss_b200.vae import B200VAEDecoder, B200VAEProxy
    cfg = ov.vae_tiny_config(latent_channels=16, shift=0.0609, pq=True)
    sd = {k: v.bfloat16().float() for k, v in ov.init_vae_decoder_weights(cfg, 0).items()}
    fake_vae = SimpleNamespace(config=SimpleNamespace(scaling_factor=cfg.scaling_factor), marker=7)
    proxy = B200VAEProxy(fake_vae, B200VAEDecoder(sd, cfg, device=cuda))
    assert proxy.marker == 7 and proxy.config.scaling_factor == cfg.scaling_factor
    g = torch.Generator().manual_seed(9)
    lat = (torch.randn(2, 16, 16, 8, generator=g) * 0.5).bfloat16()
    z = ov.unscale_latents(cfg, lat.float()).bfloat16()          # what the reference passes in
    (img,) = proxy.decode(z.cuda(), return_dict=False)
    assert proxy.decode(z.cuda()).sample.shape == img.shape == (2, 3, 128, 64)
    ref = ov.decode_single(sd, cfg, z.float())
    for i in range(2):
        a, b = img[i].float().cpu().flatten(), ref[i].flatten()
        assert torch.nn.functional.cosine_similarity(a, b, dim=0).item() >= 0.999
        assert (a - b).abs().max() <= 0.06 * (b.max() - b.min())
