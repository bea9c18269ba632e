"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (uni3detr_b200/).

CPU restatement (numpy) of the point-cloud half of the reference's test pipeline, the step before the
hot path: `LoadPointsFromFile`, `PointsRangeFilter`, `PointSample` as configured at
projects/configs/uni3detr/uni3detr_sunrgbd.py:175-191 (and uni3detr_scannet_large.py, which loads 6 dims
and uses 3).

PARITY STATUS: `LoadPointsFromFile` and `PointsRangeFilter` are **unpinned by the reference** (mmdet3d
v1.0.0rc5 classes, mmdet3d/datasets/pipelines/loading.py, transforms_3d.py: not vendored, not installable
here; restated from the published source; `np.percentile(z, 0.99)` behind `shift_height` is numpy's own
function and is called as such). `PointSample` is **pinned**: the reference carries a first-party copy of the
class (projects/mmdet3d_plugin/models/detectors/uni3detr.py:50-111) and tests/golden/make_golden_pipeline.py
stores its choices under fixed legacy numpy seeds (tests/test_prestage.py).
"""
import numpy as np

F32 = np.float32


def load_points(raw, load_dim, use_dim, shift_height=False):
    """LoadPointsFromFile.__call__: reshape(-1, load_dim)[:, use_dim]; shift_height inserts
    z - np.percentile(z, 0.99) after the xyz columns. `use_dim` int n means range(n)."""
    pts = np.asarray(raw, F32).reshape(-1, load_dim)
    if isinstance(use_dim, int):
        use_dim = list(range(use_dim))
    pts = pts[:, list(use_dim)]
    floor = F32(0)
    if shift_height:
        floor = F32(np.percentile(pts[:, 2], 0.99)) if len(pts) else F32(0)
        height = (pts[:, 2] - floor).astype(F32)
        pts = np.concatenate([pts[:, :3], height[:, None], pts[:, 3:]], 1)
    return pts.astype(F32), floor


def range_filter(points, pc_range):
    """PointsRangeFilter -> BasePoints.in_range_3d: strict inequalities on x, y, z; order kept."""
    p = np.asarray(points, F32)
    r = np.asarray(pc_range, F32)
    m = ((p[:, 0] > r[0]) & (p[:, 1] > r[1]) & (p[:, 2] > r[2]) &
         (p[:, 0] < r[3]) & (p[:, 1] < r[4]) & (p[:, 2] < r[5]))
    return p[m]


def sample_choices(n_points, num_samples, rng):
    """PointSample._points_random_sampling (uni3detr.py:67-111 / mmdet3d transforms_3d.py) with
    sample_range=None: replace only when there are fewer points than samples. `rng`: a numpy Generator or a
    legacy RandomState (the reference draws from the global legacy stream, np.random.choice)."""
    replace = n_points < num_samples
    return rng.choice(n_points, num_samples, replace=replace)


def point_sample(points, choices):
    return np.asarray(points, F32)[np.asarray(choices)]
