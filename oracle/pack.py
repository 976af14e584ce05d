"""ORACLE (test infrastructure): the reference's pack / scatter indexing, restated with plain
Python loops + numpy (integer / index work; small cases only).

Pinned against the reference's own code by tests/golden/pack_*.npz (tools/make_golden.py runs
sduss/model_executor/modules/utils.py and the split/concat methods of modules/unet.py):
  split_sample      modules/unet.py:104-184 (twin: modules/utils.py:4-83)
  concat_sample     modules/unet.py:187-202          (2-D patch stitch, SDXL)
  split_sample_sd3  modules/utils.py:86-122          (256-token chunks, SD3)
  concat_sample     modules/utils.py:124-136         (token regroup, SD3)
"""
import math

import numpy as np


def split_tables(resolutions, counts, patch_size=256):
    """Integer tables of split_sample for `counts[i]` latents at `resolutions[i]` (dict order).
    Returns padding_idx [P,4] (top,left,bottom,right neighbour or -1), latent_offset [L+1],
    resolution_offset [R+1], patch_map [P] (1-based latent id)."""
    latent_offset, resolution_offset, patch_map, padding = [0], [0], [], []
    for res, n in zip(resolutions, counts):
        if n == 0:
            continue
        ph = res // patch_size
        for _ in range(n):
            latent_offset.append(latent_offset[-1] + ph * ph)
            for h in range(ph):
                for w in range(ph):
                    me = len(padding)
                    if ph == 1:
                        padding.append([-1, -1, -1, -1])
                    else:
                        top = -1 if h == 0 else me - ph
                        bottom = -1 if h == ph - 1 else me + ph
                        left = -1 if w == 0 else me - 1
                        right = -1 if w == ph - 1 else me + 1
                        padding.append([top, left, bottom, right])
                    patch_map.append(len(latent_offset) - 1)
        resolution_offset.append(len(latent_offset) - 1)
    return (np.asarray(padding, np.int32).reshape(-1, 4), np.asarray(latent_offset, np.int32),
            np.asarray(resolution_offset, np.int32), np.asarray(patch_map, np.int32))


def split_sample(samples, patch_size=256):
    """samples: dict res -> ndarray [n, C, h, w]. Returns haloed patches [P, C, ps+2, ps+2]
    (zero border at image edges) in the reference's order, plus the integer tables."""
    lp = patch_size // 8
    out = []
    for res, arr in samples.items():
        if arr is None or arr.shape[0] == 0:
            continue
        ph = int(res) // patch_size
        for s in arr:
            p = np.pad(s, ((0, 0), (1, 1), (1, 1)))
            for h in range(ph):
                for w in range(ph):
                    out.append(p[:, h * lp:(h + 1) * lp + 2, w * lp:(w + 1) * lp + 2])
    res_list = [int(r) for r, a in samples.items() if a is not None and a.shape[0]]
    counts = [a.shape[0] for a in samples.values() if a is not None and a.shape[0]]
    return (np.stack(out),) + split_tables(res_list, counts, patch_size)


def concat_sample_2d(patches, latent_offset, patch_size=256):
    """Inverse of split_sample on un-haloed patches [P, C, ps, ps] -> dict str(res) -> [n,C,h,w]."""
    out = {}
    for i in range(len(latent_offset) - 1):
        n = int(latent_offset[i + 1] - latent_offset[i])
        ph = int(math.sqrt(n))
        rows = []
        for h in range(ph):
            rows.append(np.concatenate(
                [patches[latent_offset[i] + h * ph + w] for w in range(ph)], axis=-1))
        out.setdefault(str(ph * patch_size), []).append(np.concatenate(rows, axis=-2))
    return {k: np.stack(v) for k, v in out.items()}


def split_sample_sd3(samples, patch_size=256):
    """samples: dict res -> [n, S, D] tokens. Every latent is cut into (res/ps)^2 consecutive
    chunks of S/(res/ps)^2 tokens (flat token order). Returns chunks [T, 256, D] and tables."""
    latent_offset, resolution_offset, chunks = [0], [0], []
    for res, arr in samples.items():
        if arr is None or arr.shape[0] == 0:
            continue
        n_chunks = (int(res) // patch_size) ** 2
        for s in arr:
            latent_offset.append(latent_offset[-1] + n_chunks)
            chunks.extend(np.split(s, n_chunks, axis=0))
        resolution_offset.append(len(latent_offset) - 1)
    return (np.stack(chunks), np.asarray(latent_offset, np.int32),
            np.asarray(resolution_offset, np.int32))


def concat_sample_sd3(chunks, latent_offset, patch_size=256):
    out = {}
    for i in range(len(latent_offset) - 1):
        n = int(latent_offset[i + 1] - latent_offset[i])
        size = int(math.sqrt(n)) * patch_size
        arr = chunks[latent_offset[i]:latent_offset[i + 1]].reshape(1, -1, chunks.shape[-1])
        out.setdefault(str(size), []).append(arr)
    return {k: np.concatenate(v, axis=0) for k, v in out.items()}
