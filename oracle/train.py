"""ORACLE (test infrastructure only - never imported by uni3detr_b200/): CPU restatement of the training side
of Uni3DETRHead, written as the reference structures it: one decoder layer at a time, one image at a time,
one scipy.optimize.linear_sum_assignment call per query group.

Follows (first-party): projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:510-793 (`_get_target_single`,
`get_targets`, `loss_single`, `loss`), core/bbox/assigners/hungarian_assigner_3d.py:53-151,
core/bbox/match_costs/match_cost.py:9-33,92-104, models/losses/rdiouloss.py:94-223, core/bbox/util.py:8-80.
Third-party pieces restated from mmdet 2.x / mmdet3d 1.0.0rc5 (FocalLossCost, L1Loss, PseudoSampler,
bbox_overlaps, nearest_bev, bbox_overlaps_3d): PARITY UNPINNED for those; the first-party control flow and
arithmetic are pinned by tests/golden/golden_loss.npz, which the reference's own files produce
(tests/golden/make_golden_loss.py).
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from . import model as M
from . import postproc as PP


def normalize_bbox(b):
    """core/bbox/util.py:8-42, mmdet3d >= 1.0 branch."""
    rot = -b[..., 6:7] - np.pi / 2
    return torch.cat([b[..., 0:1], b[..., 1:2], (b[..., 4:5] + 1e-5).log(), (b[..., 3:4] + 1e-5).log(), b[..., 2:3],
                      (b[..., 5:6] + 1e-5).log(), rot.sin(), rot.cos()], dim=-1)


def nearest_bev(b):
    """mmdet3d BaseInstance3DBoxes.nearest_bev."""
    rot = b[:, 6]
    normed = (rot - torch.floor(rot / np.pi + 0.5) * np.pi).abs()
    dims = torch.where((normed > np.pi / 4)[:, None], b[:, [4, 3]], b[:, [3, 4]])
    return torch.cat([b[:, :2] - dims / 2, b[:, :2] + dims / 2], 1)


def iou2d(a, b, aligned, eps=1e-6):
    """mmdet bbox_overlaps(mode='iou')."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    if aligned:
        wh = (torch.min(a[:, 2:], b[:, 2:]) - torch.max(a[:, :2], b[:, :2])).clamp(min=0)
        inter = wh[:, 0] * wh[:, 1]
        union = area_a + area_b - inter
    else:
        wh = (torch.min(a[:, None, 2:], b[None, :, 2:]) - torch.max(a[:, None, :2], b[None, :, :2])).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        union = area_a[:, None] + area_b[None, :] - inter
    return inter / union.clamp(min=eps)


def nearest_iou(a, b, aligned=False):
    return iou2d(nearest_bev(a), nearest_bev(b), aligned)


def iou3d_rotated_aligned(a, b):
    """diag of mmdet3d bbox_overlaps_3d(coordinate='lidar'): float64 polygon clipping (oracle/postproc.py)."""
    out = np.zeros(len(a), np.float32)
    an, bn = a.detach().double().numpy(), b.detach().double().numpy()
    for i in range(len(a)):
        h = min(an[i, 2] + an[i, 5], bn[i, 2] + bn[i, 5]) - max(an[i, 2], bn[i, 2])
        if h <= 0:
            continue
        inter = PP.poly_clip_area(PP.rect_corners(an[i]), PP.rect_corners(bn[i])) * h
        out[i] = inter / max(an[i, 3] * an[i, 4] * an[i, 5] + bn[i, 3] * bn[i, 4] * bn[i, 5] - inter, 1e-8)
    return torch.from_numpy(out)


def assign(box, cls, gt, gl, nq, rep, w_cls=2.0, w_reg=0.25, w_iou=1.2, alpha=0.25, gamma=2, eps=1e-12):
    """hungarian_assigner_3d.py:53-151 -> assigned_gt_inds (Q,) (0 = background, i+1 = gt i)."""
    Q = box.shape[0]
    out = torch.zeros(Q, dtype=torch.long)
    if gt.shape[0] == 0:
        return out
    p = cls.sigmoid()
    neg = -(1 - p + eps).log() * (1 - alpha) * p.pow(gamma)
    pos = -(p + eps).log() * alpha * (1 - p).pow(gamma)
    cost = (pos[:, gl] - neg[:, gl]) * w_cls
    cost = cost + torch.cdist(box[:, :8], normalize_bbox(gt)[:, :8], p=1) * w_reg
    cost = cost + (1 - nearest_iou(M.denormalize_bbox(box), gt)) * w_iou
    cost = cost.detach().numpy()
    for g in range(Q // nq):
        r, c = linear_sum_assignment(np.tile(cost[g * nq:(g + 1) * nq], (1, rep)))
        out[g * nq + torch.from_numpy(r)] = torch.from_numpy(c % cost.shape[1]) + 1
    return out


def loss_single(cls, box, iou, gts, gls, nq, C, rep, code_weights, lw=(1.5, 0.25, 1.2), gamma=2.0, alpha=0.25):
    """uni3detr_head.py:623-698 for one decoder layer: cls (B,Q,C), box (B,Q,8), iou (B,Q,1).
    Reference quirk kept: `_get_target_single` (:542-543) passes `self.gt_repeattimes` POSITIONALLY into
    `assign(bbox_pred, cls_pred, gt_bboxes, gt_labels, num_query, gt_bboxes_ignore, eps, gt_repeattimes)`,
    where it lands in `eps` (unused) - so inside the loss the cost columns are never repeated (`rep` is ignored)."""
    B, Q, _ = cls.shape
    labels, targets, weights = [], [], []
    for b in range(B):
        inds = assign(box[b].detach(), cls[b].detach(), gts[b], gls[b], nq, 1)
        pos = inds > 0
        lab = torch.full((Q,), C, dtype=torch.long)
        lab[pos] = gls[b][inds[pos] - 1]
        tgt = torch.zeros(Q, 7)
        tgt[pos] = gts[b][inds[pos] - 1][:, :7]
        labels.append(lab); targets.append(tgt); weights.append(pos.float())
    labels, targets, w = torch.cat(labels), torch.cat(targets), torch.cat(weights)
    npos = max(float(w.sum()), 1.0)
    cls, box, iou = cls.reshape(-1, C), box.reshape(-1, 8), iou.reshape(-1)
    b3d = M.denormalize_bbox(box)
    i3d = nearest_iou(b3d, targets, aligned=True)
    z1, z2 = b3d[:, 2] - b3d[:, 5] / 2, b3d[:, 2] + b3d[:, 5] / 2
    z3, z4 = targets[:, 2] - targets[:, 5] / 2, targets[:, 2] + targets[:, 5] / 2
    iou_z = (torch.min(z2, z4) - torch.max(z1, z3)).clamp(min=0) / (torch.max(z2, z4) - torch.min(z1, z3))
    score = (i3d + iou_z) / 2
    # SoftFocalLoss (rdiouloss.py:162-181)
    soft = F.one_hot(labels, C + 1)[:, :C].float() * score[:, None]
    ps = cls.sigmoid()
    fw = ((1 - alpha) + (2 * alpha - 1) * soft) * (soft - ps).pow(gamma)
    loss_cls = lw[0] * (F.binary_cross_entropy_with_logits(cls, soft, reduction="none") * fw).sum() / npos
    bw = w[:, None] * code_weights[None, :8]
    loss_bbox = lw[1] * ((box - normalize_bbox(targets)).abs() * bw).sum() / npos
    loss_iou = lw[2] * ((1 - i3d) * bw.mean(-1)).sum() / npos + ((1 - iou_z) * bw[:, 0]).sum() / npos
    true = iou3d_rotated_aligned(b3d, targets)
    loss_iou_pred = (F.binary_cross_entropy_with_logits(iou, true, reduction="none") * bw[:, 0]).sum() / npos * 1.2
    return loss_cls, loss_bbox, loss_iou, loss_iou_pred


def loss(preds, gts, gls, nq, C, rep, code_weights=None):
    """uni3detr_head.py:716-793."""
    cw = torch.ones(8) if code_weights is None else code_weights
    L = preds["all_cls_scores"].shape[0]
    per = [loss_single(preds["all_cls_scores"][l], preds["all_bbox_preds"][l], preds["all_iou_preds"][l], gts, gls, nq, C,
                       rep, cw) for l in range(L)]
    names = ("loss_cls", "loss_bbox", "loss_iou", "loss_iou_pred")
    d = {n: v for n, v in zip(names, per[-1])}
    for i in range(L - 1):
        for n, v in zip(names, per[i]):
            d[f"d{i}.{n}"] = v
    return d


def forward_train(sd, model_cfg, points_list, gts, gls):
    """Uni3DETR.forward_train (detectors/uni3detr.py:232-266): train-mode forward (train max_voxels, BatchNorm
    batch statistics, 3 query groups, dropout off) + the loss dict, differentiable w.r.t. the tensors in `sd`."""
    head = model_cfg["pts_bbox_head"]
    M.BN_TRAIN = True
    try:
        outs, _, _ = M.forward(sd, model_cfg, points_list, random_point=None, training=True, keep_graph=True)
    finally:
        M.BN_TRAIN = False
    cw = torch.tensor(head.get("code_weights") or [1.0] * 8)
    return loss(outs, gts, gls, head["num_query"], head["num_classes"], head.get("gt_repeattimes", 1), cw)
