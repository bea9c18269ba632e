"""`box_merging` post-processing of Uni3DETRHead.get_bboxes (the KITTI config,
uni3detr_kitti_3classes.py:115-117): uni3detr_head.py:881-892 ->
projects/mmdet3d_plugin/core/bbox/bbox_merging.py (`nms_boxes_3d_merge_only` with
`overlapped_boxes_3d_fast_poly`, overlapped_thres 0.1).

Host-side numpy (float64), like the reference, which moves the decoded boxes to the CPU for this step
(uni3detr_head.py:884: `labels.cpu().numpy()` ...): a greedy, data-dependent sequential merge of <= a few
hundred boxes per scene after the device-resident decode - not on the benchmarked path. The reference's
shapely polygons are replaced by a convex-quadrilateral clipper.

Semantics kept as they are in the reference, including its quirk: `boxes_3d_to_corners` is a camera-frame
routine (x, y, z, l, h, w, yaw about the y axis) that the head feeds with LiDAR boxes
(x, y, z_bottom, dx, dy, dz, yaw), so the "bird's-eye" polygon lives in the (x, z) plane with extents
(dx, dz) and the height axis is y with extent dy, measured downwards from y.
"""
import numpy as np


def boxes_3d_to_corners(boxes_3d):
    """bbox_merging.py:12-31, vectorised: (n,7) -> (n,8,3)."""
    b = np.asarray(boxes_3d, np.float64)
    l, h, w, yaw = b[:, 3], b[:, 4], b[:, 5], b[:, 6]
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1], np.float64) * 0.5
    sz = np.array([1, -1, -1, 1, 1, -1, -1, 1], np.float64) * 0.5
    sy = np.array([0, 0, 0, 0, -1, -1, -1, -1], np.float64)
    cx, cy, cz = l[:, None] * sx, h[:, None] * sy, w[:, None] * sz          # (n,8)
    c, s = np.cos(yaw)[:, None], np.sin(yaw)[:, None]
    # corners.dot(R^T), R = [[c,0,s],[0,1,0],[-s,0,c]]
    x = cx * c + cz * s
    z = -cx * s + cz * c
    out = np.stack([x, cy, z], -1)
    return out + b[:, None, :3]


def _ccw(p):
    a = np.sum(p[:, 0] * np.roll(p[:, 1], -1) - p[:, 1] * np.roll(p[:, 0], -1))
    return (p, 0.5 * a) if a >= 0 else (p[::-1], -0.5 * a)


def _clip_area(subject, clip):
    """Area of convex polygon `subject` inside convex counter-clockwise polygon `clip` (Sutherland-Hodgman)."""
    poly = subject
    for e in range(len(clip)):
        if len(poly) == 0:
            return 0.0
        p0, p1 = clip[e], clip[(e + 1) % len(clip)]
        ex, ey = p1[0] - p0[0], p1[1] - p0[1]
        d = ex * (poly[:, 1] - p0[1]) - ey * (poly[:, 0] - p0[0])          # >= 0: inside
        nxt = np.roll(poly, -1, 0)
        dn = np.roll(d, -1)
        out = []
        for i in range(len(poly)):
            if d[i] >= 0:
                out.append(poly[i])
            if (d[i] >= 0) != (dn[i] >= 0):
                u = d[i] / (d[i] - dn[i])
                out.append(poly[i] + u * (nxt[i] - poly[i]))
        poly = np.asarray(out, np.float64).reshape(-1, 2)
    if len(poly) < 3:
        return 0.0
    return 0.5 * abs(np.sum(poly[:, 0] * np.roll(poly[:, 1], -1) - poly[:, 1] * np.roll(poly[:, 0], -1)))


def overlapped_boxes_3d_fast_poly(single_box, box_list):
    """bbox_merging.py:68-93: 3-D overlap of one corner set against a list of corner sets."""
    mx0, mn0 = single_box.max(0), single_box.min(0)
    mx, mn = box_list.max(1), box_list.min(1)
    overlap = np.zeros(len(box_list))
    apart = np.any((mx0 < mn) | (mn0 > mx), axis=1)
    p1, area1 = _ccw(single_box[:4][:, [0, 2]])
    for i in np.nonzero(~apart)[0]:
        p2, area2 = _ccw(box_list[i][:4][:, [0, 2]])
        shared_area = _clip_area(p1, p2)
        shared_y = min(mx[i][1], mx0[1]) - max(mn[i][1], mn0[1])
        intersection = shared_y * shared_area
        union = (mx[i][1] - mn[i][1]) * area2 + (mx0[1] - mn0[1]) * area1
        overlap[i] = np.float32(intersection) / (union - intersection)
    return overlap


def nms_boxes_3d_merge_only(class_labels, detection_boxes_3d, detection_scores, overlapped_thres=0.1):
    """bbox_merging.py:159-174 (+ bboxes_sort :96-115, bboxes_nms_merge_only :118-157) with top_k = -1:
    sort by score (descending), then for every still-kept box i replace it by the per-coordinate MEDIAN of
    itself and the later same-class boxes overlapping it by more than the threshold, and drop those.
    Returns (labels, boxes, scores, kept indices into the score-sorted order)."""
    scores = np.asarray(detection_scores)
    order = np.argsort(-scores)
    classes = np.asarray(class_labels)[order]
    scores = scores[order]
    bboxes = np.array(detection_boxes_3d)[order]             # copy: rows are overwritten by the medians
    corners = boxes_3d_to_corners(bboxes)                    # computed once, before any merge (:131)
    n = scores.size
    keep = np.ones(n, bool)
    for i in range(n - 1):
        if not keep[i]:
            continue
        valid = np.nonzero(keep[i + 1:])[0] + i + 1
        if len(valid) == 0:
            continue
        overlap = overlapped_boxes_3d_fast_poly(corners[i], corners[valid])
        remove = (overlap > overlapped_thres) & (classes[valid] == classes[i])
        merged = np.concatenate([bboxes[valid][remove], bboxes[[i]]], axis=0)
        bboxes[i] = np.median(merged, axis=0)
        keep[valid[remove]] = False
    idx = np.nonzero(keep)[0]
    return classes[idx], bboxes[idx], scores[idx], idx
