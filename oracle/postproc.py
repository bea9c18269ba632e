"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (uni3detr_b200/).

CPU restatement (numpy float64) of the post-processing that follows the forward hot path
(SURVEY.md §8f rank 1): Uni3DETRHead.get_bboxes,
projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:827-918 for
post_processing.type == 'nms' (per-class mmcv nms3d at :847-871, score_thr :895-908, num_thr :910-914).

PARITY STATUS: the first-party control flow (bottom-centre shift, per-class loop, class-major output
order, score_thr scalar / per-class list, num_thr) is **pinned**: tests/golden/make_golden_getbboxes.py runs
the reference's OWN get_bboxes + NMSFreeCoder.decode (golden_get_bboxes.npz) and
tests/test_oracle_golden.py::test_get_bboxes_nms_control_flow checks get_bboxes_nms against it.
mmcv.ops.nms3d itself (iou3d_nms3d_forward) is third-party and not vendored - **unpinned by the reference**
(the golden run uses this file's nms3d in its place). It is restated from its published algorithm: sort by
score (descending), rotated-rectangle BEV IoU over (x, y, dx, dy, heading), greedy suppression of boxes
with IoU > threshold. The IoU here is the exact polygon-intersection area (float64 clipping), checked
against closed-form cases in tests/test_oracle.py.
"""
import numpy as np


def rect_corners(b):
    """[x, y, z, dx, dy, dz, heading] -> (4,2) corners, counter-clockwise."""
    c, s = np.cos(b[6]), np.sin(b[6])
    hx, hy = 0.5 * b[3], 0.5 * b[4]
    loc = np.array([[hx, hy], [-hx, hy], [-hx, -hy], [hx, -hy]])
    rot = np.array([[c, -s], [s, c]])
    return loc @ rot.T + np.array([b[0], b[1]])


def poly_clip_area(pa, pb):
    """Area of convex polygon pa clipped by convex ccw polygon pb (Sutherland-Hodgman)."""
    poly = [tuple(p) for p in pa]
    for e in range(len(pb)):
        p0, p1 = pb[e], pb[(e + 1) % len(pb)]
        ex, ey = p1[0] - p0[0], p1[1] - p0[1]
        out = []
        for i in range(len(poly)):
            s, t = poly[i], poly[(i + 1) % len(poly)]
            ds = ex * (s[1] - p0[1]) - ey * (s[0] - p0[0])
            dt = ex * (t[1] - p0[1]) - ey * (t[0] - p0[0])
            if ds >= 0:
                out.append(s)
            if (ds >= 0) != (dt >= 0):
                u = ds / (ds - dt)
                out.append((s[0] + u * (t[0] - s[0]), s[1] + u * (t[1] - s[1])))
        poly = out
        if not poly:
            return 0.0
    a = 0.0
    for i in range(len(poly)):
        s, t = poly[i], poly[(i + 1) % len(poly)]
        a += s[0] * t[1] - s[1] * t[0]
    return abs(a) * 0.5


def bev_iou(a, b):
    """Rotated BEV IoU of two [x,y,z,dx,dy,dz,heading] boxes (mmcv iou3d: EPS = 1e-8)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    so = poly_clip_area(rect_corners(a), rect_corners(b))
    return so / max(a[3] * a[4] + b[3] * b[4] - so, 1e-8)


def nms3d(boxes, scores, thr):
    """mmcv.ops.nms3d: indices (into `boxes`) of the kept boxes, score-descending."""
    order = np.argsort(-np.asarray(scores), kind="stable")
    keep = []
    for i in order:
        if all(bev_iou(boxes[i], boxes[j]) <= thr for j in keep):
            keep.append(int(i))
    return np.asarray(keep, np.int64)


def get_bboxes_nms(decoded, num_classes, post_processing):
    """uni3detr_head.py:827-918 for one scene, post_processing.type == 'nms'.
    decoded: dict(bboxes (n,7|9), scores (n,), labels (n,)) from NMSFreeCoder.decode.
    Returns (bboxes, scores, labels) after the bottom-centre shift, per-class NMS (class-major
    order), score_thr and num_thr."""
    b = np.array(decoded["bboxes"], np.float64)
    s = np.asarray(decoded["scores"], np.float64)
    l = np.asarray(decoded["labels"], np.int64)
    b[:, 2] = b[:, 2] - b[:, 5] * 0.5                                        # :842
    ob, os_, ol = [], [], []
    for j in range(num_classes):                                             # :853
        ind = l == j
        if ind.sum() == 0:
            continue
        bj, sj = b[ind], s[ind]
        k = nms3d(bj[:, :7], sj, post_processing["nms_thr"])                 # :861
        ob.append(bj[k]); os_.append(sj[k]); ol.extend([j] * len(k))
    if not ob:
        return np.zeros((0, b.shape[1])), np.zeros(0), np.zeros(0, np.int64)
    b, s, l = np.concatenate(ob), np.concatenate(os_), np.asarray(ol, np.int64)
    if "score_thr" in post_processing:                                       # :895-908
        thr = post_processing["score_thr"]
        if isinstance(thr, (list, tuple)):
            ind = np.zeros(len(s), bool)
            for j in range(num_classes):
                ind |= (l == j) & (s > thr[j])
        else:
            ind = s > thr
        b, s, l = b[ind], s[ind], l[ind]
    if "num_thr" in post_processing:                                         # :910-914
        ind = np.argsort(-s, kind="stable")[: post_processing["num_thr"]]
        b, s, l = b[ind], s[ind], l[ind]
    return b, s, l
