"""Pins the oracle against fixtures produced by the reference's own code
(tools/make_golden.py): pack / scatter indexing is bit-exact, scheduler updates match to fp32
rounding (tolerance stated per assertion)."""
import os

import numpy as np
import torch

from oracle import pack as opack
from oracle import schedulers as osch

G = os.path.join(os.path.dirname(__file__), "golden")


def _cases(z, tag):
    return sorted(k[len(tag) + 4:] for k in z.files if k.startswith(tag + "_in_"))


def test_split_concat_sdxl_bit_exact():
    z = np.load(os.path.join(G, "pack_sdxl.npz"))
    for tag in ("a", "b"):
        res = sorted(_cases(z, tag), key=int)
        samples = {r: z[f"{tag}_in_{r}"].astype(np.float32) for r in res}
        patches, padding, lat_off, res_off, pmap = opack.split_sample(samples)
        assert np.array_equal(patches.astype(np.int32), z[f"{tag}_patches"].astype(np.int32))
        assert np.array_equal(padding, z[f"{tag}_padding_idx"])
        assert np.array_equal(lat_off, z[f"{tag}_latent_offset"])
        assert np.array_equal(res_off, z[f"{tag}_resolution_offset"])
        assert np.array_equal(pmap, z[f"{tag}_patch_map"])
        back = opack.concat_sample_2d(patches[:, :, 1:-1, 1:-1], lat_off)
        for r in res:
            assert np.array_equal(back[r].astype(np.int32), z[f"{tag}_back_{r}"].astype(np.int32))
            assert np.array_equal(back[r], samples[r])  # scatter(pack(x)) == x


def test_split_concat_sd3_bit_exact():
    z = np.load(os.path.join(G, "pack_sd3.npz"))
    for tag in ("a", "b"):
        res = sorted(_cases(z, tag), key=int)
        samples = {r: z[f"{tag}_in_{r}"].astype(np.float32) for r in res}
        chunks, lat_off, res_off = opack.split_sample_sd3(samples)
        assert np.array_equal(chunks.astype(np.int32), z[f"{tag}_chunks"].astype(np.int32))
        assert np.array_equal(lat_off, z[f"{tag}_latent_offset"])
        assert np.array_equal(res_off, z[f"{tag}_resolution_offset"])
        # the reference's chunk stack, flattened, IS the packed token buffer [sum S_i, D]
        flat = np.concatenate([samples[r].reshape(-1, samples[r].shape[-1]) for r in res])
        assert np.array_equal(chunks.reshape(-1, chunks.shape[-1]), flat)
        back = opack.concat_sample_sd3(chunks, lat_off)
        for r in res:
            assert np.array_equal(back[r].astype(np.int32), z[f"{tag}_back_{r}"].astype(np.int32))


def test_euler_matches_reference():
    z = np.load(os.path.join(G, "sched_euler.npz"))
    for pt in ("epsilon", "v_prediction"):
        for dn, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
            tag = f"{pt}_{dn}"
            steps, idx = z[tag + "_steps"], z[tag + "_idx"]
            tabs = [osch.euler_sigmas(int(n))[0] for n in steps]
            sig = [t[i] for t, i in zip(tabs, idx)]
            sig_next = [t[i + 1] for t, i in zip(tabs, idx)]
            x = torch.from_numpy(z[tag + "_x"]).to(dt)
            eps = torch.from_numpy(z[tag + "_eps"]).to(dt)
            scaled = osch.batch_scale_model_input(torch.cat([x, x]), sig)
            prev = osch.euler_batch_step(eps, x, sig, sig_next, pt)
            # same op order as the reference -> identical results (0 tolerance)
            assert torch.equal(scaled.float(), torch.from_numpy(z[tag + "_scaled"]))
            assert torch.equal(prev.float(), torch.from_numpy(z[tag + "_prev"]))


def test_flow_match_matches_reference():
    z = np.load(os.path.join(G, "sched_flow_match.npz"))
    for dn, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        steps, idx = z[dn + "_steps"], z[dn + "_idx"]
        tabs = [osch.flow_match_sigmas(int(n))[0] for n in steps]
        sig = [t[i] for t, i in zip(tabs, idx)]
        sig_next = [t[i + 1] for t, i in zip(tabs, idx)]
        x = torch.from_numpy(z[dn + "_x"]).to(dt)
        v = torch.from_numpy(z[dn + "_v"]).to(dt)
        prev = osch.flow_match_batch_step(v, x, sig, sig_next)
        assert torch.equal(prev.float(), torch.from_numpy(z[dn + "_prev"]))


def test_sigma_tables_closed_form():
    s, t, init = osch.euler_sigmas(50)
    assert s.shape == (51,) and t.shape == (50,) and s[-1] == 0 and t[0] == 981 and t[-1] == 1
    assert abs(init - (s[0].item() ** 2 + 1) ** 0.5) < 1e-6 and 13.0 < s[0] < 13.3
    s, t = osch.flow_match_sigmas(28)
    assert s.shape == (29,) and abs(s[0].item() - 1.0) < 1e-6 and s[-1] == 0
    assert torch.allclose(t, s[:-1] * 1000)
    assert torch.all(s[:-1] > s[1:])
