"""Golden vectors for the `box_merging` post-processing of the KITTI config
(projects/configs/uni3detr/uni3detr_kitti_3classes.py:115-117 ->
projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:881-892 ->
projects/mmdet3d_plugin/core/bbox/bbox_merging.py).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_postproc.py
Output (committed): tests/golden/golden_box_merging.npz

The reference's OWN file bbox_merging.py is executed unmodified from /root/reference: the score sort, the
greedy same-class merge with the per-coordinate median, the corner construction (a camera-frame routine
applied to LiDAR boxes - reproduced as is) and the overlap formula all run from the reference's code. Only
its un-installable import is stubbed: `shapely.geometry.Polygon` (area, intersection().area) [restated with
the oracle's float64 convex clipping - so the polygon arithmetic itself is NOT a pin]. `np.bool`, removed
from numpy >= 1.24 and used at bbox_merging.py:133, is aliased to `bool`.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import postproc as OP  # noqa: E402


class _Area:
    def __init__(self, a):
        self.area = a


class Polygon:                                    # [restated] shapely.geometry.Polygon for convex quads
    def __init__(self, pts):
        p = np.asarray(pts, np.float64)
        a = 0.5 * np.sum(p[:, 0] * np.roll(p[:, 1], -1) - p[:, 1] * np.roll(p[:, 0], -1))
        self.p = p if a >= 0 else p[::-1]         # counter-clockwise
        self.area = abs(a)

    def intersection(self, other):
        return _Area(OP.poly_clip_area(self.p, other.p))


def load_reference_module():
    if not hasattr(np, "bool"):
        np.bool = bool                            # bbox_merging.py:133 predates numpy 1.24
    geo = types.ModuleType("shapely.geometry")
    geo.Polygon = Polygon
    sh = types.ModuleType("shapely")
    sh.geometry = geo
    sys.modules.setdefault("shapely", sh)
    sys.modules.setdefault("shapely.geometry", geo)
    path = os.path.join(REF, "projects/mmdet3d_plugin/core/bbox/bbox_merging.py")
    spec = importlib.util.spec_from_file_location("ref_bbox_merging", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_case(n, n_cls, seed):
    """Clustered LiDAR-style boxes [x, y, z(bottom), dx, dy, dz, yaw]: many overlapping groups."""
    rng = np.random.default_rng(seed)
    centres = np.concatenate([rng.uniform(0, 60, (10, 1)), rng.uniform(-30, 30, (10, 1)), rng.uniform(-2, 0, (10, 1))], 1)
    which = rng.integers(0, 10, n)
    b = np.zeros((n, 7), np.float32)
    b[:, :3] = centres[which] + rng.normal(0, 0.4, (n, 3)) * [1, 1, 0.2]
    b[:, 3:6] = [3.9, 1.6, 1.5] + rng.normal(0, 0.15, (n, 3))
    b[:, 6] = rng.uniform(-np.pi, np.pi, 10)[which] + rng.normal(0, 0.1, n)
    scores = rng.random(n).astype(np.float32)
    labels = (which % n_cls).astype(np.int64)          # one class per cluster + a few strays
    stray = rng.random(n) < 0.15
    labels[stray] = rng.integers(0, n_cls, stray.sum())
    return labels, b, scores


def main():
    ref = load_reference_module()
    out = {}
    for ci, (n, n_cls, seed) in enumerate([(120, 3, 0), (300, 3, 1), (1, 3, 2), (40, 1, 3)]):
        labels, boxes, scores = make_case(n, n_cls, seed)
        cl, bx, sc, idx = ref.nms_boxes_3d_merge_only(
            labels.copy(), boxes.copy(), scores.copy(), overlapped_fn=ref.overlapped_boxes_3d_fast_poly,
            overlapped_thres=0.1, appr_factor=1e6, top_k=-1, attributes=np.arange(len(labels)))
        out[f"c{ci}_in_labels"], out[f"c{ci}_in_boxes"], out[f"c{ci}_in_scores"] = labels, boxes, scores
        out[f"c{ci}_labels"], out[f"c{ci}_boxes"], out[f"c{ci}_scores"] = cl, bx, sc
        out[f"c{ci}_idx"] = np.asarray(idx[0])
        print(f"case {ci}: {n} boxes -> {len(cl)} kept")
    np.savez_compressed(os.path.join(HERE, "golden_box_merging.npz"), **out)


if __name__ == "__main__":
    main()
