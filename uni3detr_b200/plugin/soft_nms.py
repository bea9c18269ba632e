"""`soft_nms` post-processing option of Uni3DETRHead.get_bboxes (commented alternative in every shipped config:
"soft nms can generate a little higher result", uni3detr_sunrgbd.py:114-118): uni3detr_head.py:795-823 (the
Gaussian soft-NMS loop, first-party) with mmdet3d's `bbox_overlaps_3d(..., coordinate='lidar')` (third-party:
rotated-BEV intersection area x height overlap over the union volume, bottom-centre boxes) restated.

Host-side numpy (float64): a greedy, data-dependent loop over the <= a few hundred boxes of one class - not on
the benchmarked path (the reference runs it box by box through torch on whatever device the scores live on).
"""
import numpy as np

from .box_merging import _ccw, _clip_area


def _bev_corners(b):
    """[x, y, z, dx, dy, dz, yaw] -> (4,2) BEV corners (yaw counter-clockwise from +x, like mmcv iou3d)."""
    c, s = np.cos(b[6]), np.sin(b[6])
    hx, hy = 0.5 * b[3], 0.5 * b[4]
    loc = np.array([[hx, hy], [-hx, hy], [-hx, -hy], [hx, -hy]], np.float64)
    return loc @ np.array([[c, s], [-s, c]], np.float64) + b[:2]


def overlaps_3d_lidar(box, boxes):
    """mmdet3d BaseInstance3DBoxes.overlaps(mode='iou') for LiDAR boxes (z = bottom): one box vs (n,7)."""
    box = np.asarray(box, np.float64)
    boxes = np.asarray(boxes, np.float64)
    out = np.zeros(len(boxes))
    pa, _ = _ccw(_bev_corners(box))
    va = box[3] * box[4] * box[5]
    # cheap reject: centre distance vs the sum of the BEV half diagonals
    ra = 0.5 * np.hypot(box[3], box[4])
    rb = 0.5 * np.hypot(boxes[:, 3], boxes[:, 4])
    near = np.hypot(boxes[:, 0] - box[0], boxes[:, 1] - box[1]) <= ra + rb
    for i in np.nonzero(near)[0]:
        b = boxes[i]
        h = min(box[2] + box[5], b[2] + b[5]) - max(box[2], b[2])
        if h <= 0:
            continue
        pb, _ = _ccw(_bev_corners(b))
        inter = _clip_area(pa, pb) * h
        out[i] = inter / max(va + b[3] * b[4] * b[5] - inter, 1e-8)
    return out


def soft_nms(boxes, scores, gaussian_sigma=0.3, prune_threshold=1e-3):
    """uni3detr_head.py:795-823: repeatedly take the best remaining box, decay every remaining score by
    exp(-iou^2 / sigma), drop the ones that fall to the prune threshold. Returns (indices into `boxes` in
    selection order, their scores at selection time)."""
    boxes = np.array(boxes, np.float64)
    scores = np.array(scores, np.float64)
    idxs = np.arange(len(scores))
    idx_out, score_out = [], []
    while scores.size > 0:
        top = int(np.argmax(scores))
        idx_out.append(int(idxs[top]))
        score_out.append(float(scores[top]))
        ious = overlaps_3d_lidar(boxes[top], boxes)
        scores = scores * np.exp(-np.power(ious, 2) / gaussian_sigma)
        keep = scores > prune_threshold
        keep[top] = False
        boxes, scores, idxs = boxes[keep], scores[keep], idxs[keep]
    return np.asarray(idx_out, np.int64), np.asarray(score_out, np.float64)
