#!/usr/bin/env python
"""Useful fraction of the (row, offset) slots the tensor-core sparse conv multiplies, per encoder level,
for different tile orders (CPU, oracle rulebooks on synthetic SUN RGB-D scenes). Reproduces the numbers
quoted in DESIGN.md §4 / csrc/tilesort.cu.

  python scripts/tile_padding_stats.py [--scenes 4] [--tile 256] [--groups 1 2 4]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import geometry as G          # noqa: E402
from uni3detr_b200 import synth           # noqa: E402


def signature(act):
    """12-bit key of csrc/tilesort.cu: any neighbour per (kz,ky) line (9 bits) and per kx column (3 bits)."""
    key = np.zeros(act.shape[1], np.int64)
    for line in range(9):
        key |= act[3 * line:3 * line + 3].any(0).astype(np.int64) << line
    for c in range(3):
        key |= act[c::3].any(0).astype(np.int64) << (9 + c)
    return key


def efficiency(act, order, tile):
    n = act.shape[1]
    nt = (n + tile - 1) // tile
    pad = np.zeros((27, nt * tile), bool)
    pad[:, :n] = act[:, order]
    return act.sum() / (pad.reshape(27, nt, tile).any(2).sum() * tile)


def levels(n_scenes):
    cfg = synth.load_model_cfg("sunrgbd")
    vl, enc = cfg["pts_voxel_layer"], cfg["pts_middle_encoder"]
    scenes = [synth.make_scene("sunrgbd", i) for i in range(n_scenes)]
    _, _, coors, _ = G.voxelize_batch_hard(scenes, vl["point_cloud_range"], vl["voxel_size"], 5, 40000)
    c, d = coors, tuple(enc["sparse_shape"])
    strides, pads = enc.get("encoder_strides", (2, 2, 2, 1)), enc["encoder_paddings"]
    for i in range(4):
        yield f"subm stage {i}", G.subm_rulebook(c, d), c[:, 0]
        if i < 3:
            s = strides[i]
            s = tuple(s) if isinstance(s, (list, tuple)) else (s,) * 3
            p = pads[i][-1]
            p = tuple(p) if isinstance(p, (list, tuple)) else (p,) * 3
            oc, nb, od = G.down_rulebook(c, d, s, p)
            yield f"strided {i}->{i + 1}", nb, oc[:, 0]
            c, d = oc, od


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=4)
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--groups", type=int, nargs="*", default=[1, 2])
    a = ap.parse_args()
    print(f"{a.scenes} scenes, {a.tile}-row tiles: useful fraction of multiplied slots")
    for name, nbr, batch in levels(a.scenes):
        act = nbr >= 0
        n = act.shape[1]
        full = (act.astype(np.int64) << np.arange(27)[:, None]).sum(0)
        sig = signature(act)
        row = [f"{name:16s} rows {n:7d} pairs/row {act.sum() / n:5.1f}",
               f"natural {efficiency(act, np.arange(n), a.tile):.2f}",
               f"27-bit sort {efficiency(act, np.argsort(full, kind='stable'), a.tile):.2f}",
               f"signature {efficiency(act, np.argsort(sig, kind='stable'), a.tile):.2f}"]
        for g in a.groups:      # bucket inside groups of g scenes (keeps the gathers of a tile in g scenes)
            key = (batch.astype(np.int64) // g) * 4096 + sig
            row.append(f"sig/{g}-scene groups {efficiency(act, np.argsort(key, kind='stable'), a.tile):.2f}")
        print(" | ".join(row))


if __name__ == "__main__":
    main()
