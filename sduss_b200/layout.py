"""Packed-NHWC layout tables for a mixed-resolution batch at one UNet level."""
from typing import List, Tuple

import numpy as np
import torch


class LevelLayout:
    """Latents (h_i, w_i) packed back to back as pixel rows. Holds every int32 table the
    kernels need: lat_desc {row offset, H, W, 0}, row_group, 64-row GroupNorm chunks and the
    16x8 convolution tile list."""

    def __init__(self, sizes: List[Tuple[int, int]], device):
        self.sizes = sizes
        self.L = len(sizes)
        rows = [h * w for h, w in sizes]
        off = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
        self.row_off = [int(o) for o in off[:-1]]
        self.rows = rows
        self.T = int(off[-1])
        self.max_pixels = max(rows)
        desc = np.asarray([(self.row_off[i], h, w, 0) for i, (h, w) in enumerate(sizes)], np.int32)
        self.desc_host = desc
        self.desc = torch.from_numpy(desc).to(device)
        self.row_group = torch.from_numpy(np.repeat(np.arange(self.L, dtype=np.int32), rows)).to(device)
        assert all(r % 64 == 0 for r in rows), "GroupNorm chunks need pixel counts multiple of 64"
        chunks = np.asarray([(self.row_off[i] // 64, rows[i] // 64, 0, 0) for i in range(self.L)], np.int32)
        self.lat_chunks = torch.from_numpy(chunks).to(device)
        tiles, lat_tiles = [], []
        for i, (h, w) in enumerate(sizes):
            first = len(tiles)
            for y0 in range(0, h, 16):
                for x0 in range(0, w, 8):
                    tiles.append((i, y0, x0, 0))
            lat_tiles.append((first, len(tiles) - first, h * w, 0))
        self.n_tiles = len(tiles)
        self.tiles = torch.from_numpy(np.asarray(tiles, np.int32)).to(device)
        # per latent: {first conv tile, tiles, pixels}: where the convolution epilogue leaves the
        # per-tile GroupNorm partial sums of this latent (b200_groupnorm_from_conv_stats)
        self.lat_tiles = torch.from_numpy(np.asarray(lat_tiles, np.int32)).to(device)
