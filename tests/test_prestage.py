"""Input pre-stage (validated on hardware in round 2).

CPU part: the oracle restatement of LoadPointsFromFile / PointsRangeFilter / PointSample against
hand-computed cases and numpy, and the host-side parsing of the reference's `test_pipeline` dicts.
GPU part: csrc/points.cu vs the oracle (bit-exact rows / offsets, 1e-6 floor height)."""
import os

import numpy as np
import pytest
import torch

from oracle import pipeline as P

PCR = [-3.2, -0.2, -2.0, 3.2, 6.2, 0.56]
SUNRGBD_TEST_PIPELINE = [   # projects/configs/uni3detr/uni3detr_sunrgbd.py:175-191
    dict(type="LoadPointsFromFile", coord_type="DEPTH", shift_height=True, load_dim=6, use_dim=[0, 1, 2]),
    dict(type="PointsRangeFilter", point_cloud_range=PCR),
    dict(type="PointSample", num_points=100000),
    dict(type="DefaultFormatBundle3D", class_names=[], with_label=False),
    dict(type="Collect3D", keys=["points"])]


def test_oracle_load_points_known_answer():
    """101 points with z = 0..100: virtual index 100 * 0.0099 = 0.99 -> floor height 0.99 (hand computed)."""
    raw = np.zeros((101, 6), np.float32)
    raw[:, 0] = 1.0
    raw[:, 2] = np.arange(101)[::-1]          # order must not matter
    raw[:, 3:] = 7.0                           # colour columns: dropped by use_dim
    pts, floor = P.load_points(raw.reshape(-1), 6, [0, 1, 2], shift_height=True)
    assert pts.shape == (101, 4) and pts.dtype == np.float32
    np.testing.assert_allclose(floor, 0.99, rtol=0, atol=1e-6)
    np.testing.assert_allclose(pts[:, 3], raw[:, 2] - np.float32(0.99), rtol=0, atol=1e-6)
    np.testing.assert_array_equal(pts[:, :3], raw[:, :3])
    pts5, _ = P.load_points(np.arange(20, dtype=np.float32), 5, 5)          # nuScenes style: use_dim=5
    np.testing.assert_array_equal(pts5, np.arange(20, dtype=np.float32).reshape(4, 5))


def test_oracle_range_filter_is_strict_and_order_preserving():
    p = np.array([[0, 0, 0, 1], [3.2, 0, 0, 2], [-3.2, 0, 0, 3], [1, 6.2, 0, 4], [1, 1, 0.56, 5], [3.1, 6.1, 0.5, 6],
                  [1, 1, -2.0, 7], [-1, 2, -1, 8]], np.float32)
    np.testing.assert_array_equal(P.range_filter(p, PCR)[:, 3], [1, 6, 8])


def test_oracle_point_sample_semantics():
    rng = np.random.default_rng(0)
    c = P.sample_choices(50, 20, rng)
    assert len(c) == 20 and len(set(c.tolist())) == 20          # enough points: without replacement
    c = P.sample_choices(5, 20, rng)
    assert len(c) == 20 and set(c.tolist()) <= set(range(5))    # too few: with replacement
    pts = np.arange(15, dtype=np.float32).reshape(5, 3)
    np.testing.assert_array_equal(P.point_sample(pts, [4, 0, 4]), pts[[4, 0, 4]])


def test_oracle_point_sample_matches_the_reference_class():
    """Golden choices drawn by the reference's own PointSample class (first-party copy in uni3detr.py:50-111)
    under legacy numpy seeds == the oracle's restatement fed with the same legacy stream."""
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_point_sample.npz")))
    for ci in range(4):
        n, num, seed = int(g[f"c{ci}_n"]), int(g[f"c{ci}_num"]), int(g[f"c{ci}_seed"])
        got = P.sample_choices(n, num, np.random.RandomState(seed))
        np.testing.assert_array_equal(got, g[f"c{ci}_choices"])


def test_prestage_parses_the_reference_pipeline():
    from uni3detr_b200.prestage import PointsPreStage
    pre = PointsPreStage(SUNRGBD_TEST_PIPELINE)
    assert (pre.load_dim, pre.use_dim, pre.shift_height, pre.num_points) == (6, [0, 1, 2], True, 100000)
    assert pre.pc_range == PCR and pre.channels == 4
    with pytest.raises(NotImplementedError):
        PointsPreStage([dict(type="LoadPointsFromFile", load_dim=5, use_dim=5),
                        dict(type="LoadPointsFromMultiSweeps", sweeps_num=9)])


@pytest.mark.skipif(not os.path.isdir("/root/reference/projects/configs"), reason="reference tree absent")
def test_prestage_reads_the_shipped_configs():
    from uni3detr_b200.compat import Config
    from uni3detr_b200.prestage import PointsPreStage
    cfg = Config.fromfile("/root/reference/projects/configs/uni3detr/uni3detr_sunrgbd.py")
    pre = PointsPreStage(cfg.test_pipeline)
    assert pre.load_dim == 6 and pre.use_dim == [0, 1, 2] and pre.shift_height and pre.num_points == 100000
    assert pre.pc_range == PCR


def _float_key(f):
    u = np.float32(f).view(np.uint32)
    return np.uint32(~u) if (u & np.uint32(0x80000000)) else np.uint32(u | np.uint32(0x80000000))


def _key_float(k):
    k = np.uint32(k)
    return (np.uint32(k & np.uint32(0x7FFFFFFF)) if (k & np.uint32(0x80000000)) else np.uint32(~k)).view(np.float32)


def _radix_select_percentile(z):
    """Python model of k_points_stats (csrc/points.cu): 4 x 8-bit radix select of the lower order statistic
    on order-preserving keys, the next one from the duplicate count / the smallest larger key, lerp in double."""
    n = len(z)
    keys = np.array([_float_key(v) for v in z], dtype=np.uint64)
    q = 0.99 / 100.0
    vi = n * q + (1.0 + q * (-1.0)) - 1.0
    k_lo = max(0, min(int(np.floor(vi)), n - 1))
    gamma = vi - np.floor(vi)
    prefix, rank, less = 0, k_lo, 0
    for p in (3, 2, 1, 0):
        hi_mask = 0 if p == 3 else (0xFFFFFFFF << (8 * (p + 1))) & 0xFFFFFFFF
        sel = (keys & hi_mask) == prefix
        hist = np.bincount(((keys[sel] >> (8 * p)) & 255).astype(np.int64), minlength=256)
        acc, b = 0, 0
        while b < 256 and acc + hist[b] <= rank:
            acc += hist[b]
            b += 1
        b = min(b, 255)
        rank -= acc
        less += acc
        prefix |= b << (8 * p)
    eq = int((keys == prefix).sum())
    larger = keys[keys > prefix]
    k_hi = min(k_lo + 1, n - 1)
    a = _key_float(prefix)
    bb = a if (k_hi < less + eq or len(larger) == 0) else _key_float(int(larger.min()))
    return np.float32(float(a) + (float(bb) - float(a)) * gamma)


def test_radix_select_percentile_algorithm_matches_numpy():
    """The selection algorithm of the shift_height kernel, modelled in Python, vs np.percentile(z, 0.99):
    within 1 ulp (the two order statistics are exact; only the interpolation arithmetic differs)."""
    rng = np.random.default_rng(1)
    for t in range(60):
        n = int(rng.integers(1, 3000))
        z = (rng.normal(size=n) * rng.uniform(0.01, 10)).astype(np.float32)
        if n > 5 and t % 3 == 0:
            z[: n // 2] = z[0]                      # duplicated order statistics
        ref = np.float32(np.percentile(z, 0.99))
        got = _radix_select_percentile(z)
        assert abs(float(got) - float(ref)) <= 2 * float(np.spacing(np.float32(max(abs(ref), 1e-30)))), (n, got, ref)


def _raw_scene(n, seed):
    rng = np.random.default_rng(seed)
    raw = np.zeros((n, 6), np.float32)
    raw[:, :3] = rng.uniform([-4, -1, -2.5], [4, 7, 1.0], (n, 3))
    raw[:, 3:] = rng.random((n, 3))
    if n > 10:
        raw[3, 2] = raw[4, 2] = raw[:, 2].min()     # duplicated order statistics
    return raw


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [(20000, 1, 357), (101,), (2, 0, 5000)])
def test_gpu_points_prepare_vs_oracle(sizes):
    from uni3detr_b200 import ops
    raws = [_raw_scene(n, 10 + i) for i, n in enumerate(sizes)]
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device="cuda")
    raw = torch.from_numpy(np.concatenate(raws)).cuda()
    pts, out_off, floor = ops.points_prepare(raw, off, len(sizes), [0, 1, 2], True, PCR)
    o = out_off.cpu().numpy()
    for b, r in enumerate(raws):
        ref, ref_floor = P.load_points(r, 6, [0, 1, 2], True)
        ref = P.range_filter(ref, PCR)
        assert o[b + 1] - o[b] == len(ref)
        got = pts[o[b]:o[b + 1]].cpu().numpy()
        np.testing.assert_array_equal(got[:, :3], ref[:, :3])                       # byte copies, order kept
        if len(r):
            np.testing.assert_allclose(float(floor[b]), ref_floor, rtol=0, atol=2e-6)   # percentile: 1 ulp
        np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=0, atol=4e-6)


@pytest.mark.gpu
def test_gpu_points_gather():
    from uni3detr_b200 import ops
    pts = torch.randn(1000, 4, device="cuda")
    ch = torch.randint(0, 1000, (2500,), dtype=torch.int32, device="cuda")
    assert torch.equal(ops.points_gather(pts, ch), pts[ch.long()])
