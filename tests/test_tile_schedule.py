"""Tile scheduling rationale (csrc/tilesort.cu), checked on CPU with the oracle's rulebooks: bucketing the
rows of a level by the 12-bit neighbour signature must raise the useful fraction of the (row, offset) slots
the tensor-core conv multiplies, on every level of the SUN RGB-D encoder. Numbers quoted in DESIGN.md come
from scripts/tile_padding_stats.py (more scenes -> fuller buckets -> higher fractions)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("tile_padding_stats", os.path.join(ROOT, "scripts", "tile_padding_stats.py"))
stats = importlib.util.module_from_spec(spec)
spec.loader.exec_module(stats)


def test_signature_matches_the_device_formula():
    """scripts/ definition == the bit formula of tile_key_of_mask (csrc/tilesort.cu), on random masks."""
    rng = np.random.default_rng(0)
    act = rng.random((27, 500)) < 0.3
    mask = (act.astype(np.int64) << np.arange(27)[:, None]).sum(0)
    key = np.zeros_like(mask)
    for line in range(9):
        key |= (((mask >> (3 * line)) & 7) != 0).astype(np.int64) << line
    for c in range(3):
        key |= ((mask & (0x1249249 << c)) != 0).astype(np.int64) << (9 + c)
    np.testing.assert_array_equal(stats.signature(act), key)


def test_signature_buckets_cut_tile_padding_on_every_level():
    seen = 0
    for name, nbr, batch in stats.levels(1):
        act = nbr >= 0
        n = act.shape[1]
        nat = stats.efficiency(act, np.arange(n), 256)
        sig = stats.efficiency(act, np.argsort(stats.signature(act), kind="stable"), 256)
        assert sig > nat * 1.1, (name, nat, sig)
        if name == "subm stage 0":
            assert nat < 0.1 and sig > 0.25, (nat, sig)      # 16-channel layers: >= 3x fewer multiplied tiles
        if name == "subm stage 1":
            assert sig > nat * 1.5, (nat, sig)
        seen += 1
    assert seen == 7
