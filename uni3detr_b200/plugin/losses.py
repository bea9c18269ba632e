"""Training losses and the Hungarian assigner of Uni3DETRHead, device resident.

References (first-party): projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:510-793 (targets,
`loss_single`, `loss`), core/bbox/assigners/hungarian_assigner_3d.py:53-151, core/bbox/match_costs/
match_cost.py (BBox3DL1Cost :9-33, IoU3DCost :92-104), models/losses/rdiouloss.py (IoU3DLoss :94-160,
SoftFocalLoss :162-223), core/bbox/util.py:8-42 (normalize_bbox).
Restated third-party pieces (mmdet / mmdet3d, not vendored - marked [restated]): FocalLossCost, L1Loss,
weight_reduce_loss, PseudoSampler, bbox_overlaps (2-D IoU), BaseInstance3DBoxes.nearest_bev /
bbox_overlaps_nearest_3d, bbox_overlaps_3d (rotated 3-D IoU), reduce_mean.

What is different from the reference: nothing leaves the device. The reference computes one cost matrix per
(decoder layer, image), moves it to the CPU and calls scipy once per query group
(hungarian_assigner_3d.py:124-139: L x B x G host round trips per step); here the cost blocks of all layers
and groups of an image are solved by ONE launch of the matcher kernel (csrc/train.cu k_hungarian, one CTA
per problem) and targets / losses are batched tensor expressions under autograd.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..compat import LOSSES


# ------------------------------------------------------------------ box helpers ---
def normalize_bbox(bboxes, pc_range=None):
    """core/bbox/util.py:8-42 (mmdet3d >= 1.0 branch): (cx,cy,cz,l,w,h,rot[,vx,vy]) ->
    (cx, cy, log w, log l, cz, log h, sin r', cos r'[, vx, vy]) with r' = -rot - pi/2 and the sizes
    shifted by 1e-5 before the log."""
    cx, cy, cz = bboxes[..., 0:1], bboxes[..., 1:2], bboxes[..., 2:3]
    l = (bboxes[..., 3:4] + 1e-5).log()
    w = (bboxes[..., 4:5] + 1e-5).log()
    h = (bboxes[..., 5:6] + 1e-5).log()
    rot = -bboxes[..., 6:7] - np.pi / 2
    parts = [cx, cy, w, l, cz, h, rot.sin(), rot.cos()]
    if bboxes.size(-1) > 7:
        parts += [bboxes[..., 7:8], bboxes[..., 8:9]]
    return torch.cat(parts, dim=-1)


def nearest_bev(boxes):
    """[restated] mmdet3d BaseInstance3DBoxes.nearest_bev: the axis-aligned BEV box (x1,y1,x2,y2) of
    (x,y,z,dx,dy,dz,yaw) after snapping the yaw to the nearest multiple of pi/2."""
    rot = boxes[..., 6]
    normed = torch.abs(rot - torch.floor(rot / np.pi + 0.5) * np.pi)         # limit_period(rot, 0.5, pi)
    swap = (normed > np.pi / 4).unsqueeze(-1)
    dims = torch.where(swap, boxes[..., [4, 3]], boxes[..., [3, 4]])
    centers = boxes[..., 0:2]
    return torch.cat([centers - dims / 2, centers + dims / 2], dim=-1)


def bbox_overlaps_2d(b1, b2, is_aligned=False, eps=1e-6):
    """[restated] mmdet bbox_overlaps(mode='iou'): b1 (..., n, 4), b2 (..., m, 4) -> (..., n, m), or (..., n) aligned."""
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    if is_aligned:
        lt = torch.max(b1[..., :2], b2[..., :2])
        rb = torch.min(b1[..., 2:], b2[..., 2:])
        wh = (rb - lt).clamp(min=0)
        overlap = wh[..., 0] * wh[..., 1]
        union = a1 + a2 - overlap
    else:
        lt = torch.max(b1[..., :, None, :2], b2[..., None, :, :2])
        rb = torch.min(b1[..., :, None, 2:], b2[..., None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        overlap = wh[..., 0] * wh[..., 1]
        union = a1[..., None] + a2[..., None, :] - overlap
    union = torch.max(union, union.new_tensor([eps]))
    return overlap / union


def bbox_overlaps_nearest_3d(b1, b2, is_aligned=False):
    """[restated] mmdet3d bbox_overlaps_nearest_3d: 2-D IoU of the nearest-BEV boxes ('depth' and 'lidar'
    boxes share the BEV definition, so the `coordinate` argument of the reference does not matter)."""
    return bbox_overlaps_2d(nearest_bev(b1), nearest_bev(b2), is_aligned)


def bbox_overlaps_3d_aligned(b1, b2):
    """[restated] torch.diag(mmdet3d bbox_overlaps_3d(b1, b2, coordinate='lidar')) (uni3detr_head.py:690):
    rotated-BEV intersection x height overlap over the union volume, boxes taken as [x,y,z(bottom),dx,dy,dz,yaw].
    CUDA tensors go through u3d_iou3d_aligned; no gradient (the reference detaches it)."""
    from .. import ops
    a = b1[..., :7].detach().float().contiguous()
    b = b2[..., :7].detach().float().contiguous()
    return ops.iou3d_aligned(a, b)          # CUDA only: raises on CPU tensors (no CPU path in the product)


def bbox_to_corners_aa(bbox):
    """uni3detr_head.py:700-707 `_bbox_to_loss`: (x,y,z,w,h,l,...) -> (x1,y1,z1,x2,y2,z2)."""
    return torch.stack((bbox[..., 0] - bbox[..., 3] / 2, bbox[..., 1] - bbox[..., 4] / 2,
                        bbox[..., 2] - bbox[..., 5] / 2, bbox[..., 0] + bbox[..., 3] / 2,
                        bbox[..., 1] + bbox[..., 4] / 2, bbox[..., 2] + bbox[..., 5] / 2), dim=-1)


# ------------------------------------------------------------------ losses ---
def weight_reduce_loss(loss, weight=None, reduction="mean", avg_factor=None):
    """[restated] mmdet.models.losses.utils.weight_reduce_loss."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == "mean" else (loss.sum() if reduction == "sum" else loss)
    if reduction == "mean":
        return loss.sum() / avg_factor
    if reduction == "none":
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


def soft_focal_loss(pred, target, weight=None, gamma=2.0, alpha=0.25, reduction="mean", avg_factor=None):
    """models/losses/rdiouloss.py:162-181: quality focal loss with the soft target = IoU score at the class of
    a positive. target = (labels (N,) with C = background, scores (N,))."""
    pred_sigmoid = pred.sigmoid()
    labels, target_score = target[0], target[1]
    C = pred.shape[1]
    onehot = F.one_hot(labels.clamp(max=C), C + 1)[:, :C].to(pred.dtype)
    target_soft = onehot * target_score[:, None]
    pt = target_soft - pred_sigmoid
    focal_weight = ((1 - alpha) + (2 * alpha - 1) * target_soft) * pt.pow(gamma)
    loss = F.binary_cross_entropy_with_logits(pred, target_soft, reduction="none") * focal_weight
    return weight_reduce_loss(loss, weight.view(-1, 1), reduction, avg_factor)


@LOSSES.register_module()
class SoftFocalLoss(nn.Module):
    """models/losses/rdiouloss.py:183-223."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, "Only sigmoid focal loss supported now."
        self.use_sigmoid, self.gamma, self.alpha = use_sigmoid, gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        reduction = reduction_override if reduction_override else self.reduction
        return self.loss_weight * soft_focal_loss(pred, target, weight, self.gamma, self.alpha, reduction, avg_factor)


@LOSSES.register_module()
class L1Loss(nn.Module):
    """[restated] mmdet L1Loss."""

    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        reduction = reduction_override if reduction_override else self.reduction
        if target.numel() == 0:
            return pred.sum() * 0
        return self.loss_weight * weight_reduce_loss(torch.abs(pred - target), weight, reduction, avg_factor)


@LOSSES.register_module()
class IoU3DLoss(nn.Module):
    """models/losses/rdiouloss.py:94-160: 1 - nearest-BEV IoU of aligned box pairs."""

    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        if weight is not None and not torch.any(weight > 0):
            return pred.sum() * weight.sum()
        reduction = reduction_override if reduction_override else self.reduction
        if weight is not None and weight.dim() > 1:
            weight = weight.mean(-1)
        loss = 1 - bbox_overlaps_nearest_3d(pred, target, is_aligned=True)
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


def build_loss(cfg):
    cfg = dict(cfg)
    t = cfg.pop("type")
    return LOSSES.get(t)(**cfg)


# ------------------------------------------------------------------ match costs ---
class FocalLossCost:
    """[restated] mmdet FocalLossCost."""

    def __init__(self, weight=1.0, alpha=0.25, gamma=2, eps=1e-12, **kwargs):
        self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps

    def __call__(self, cls_pred, gt_labels):
        p = cls_pred.sigmoid()
        neg = -(1 - p + self.eps).log() * (1 - self.alpha) * p.pow(self.gamma)
        pos = -(p + self.eps).log() * self.alpha * (1 - p).pow(self.gamma)
        cost = pos - neg                                  # (..., Q, C)
        idx = gt_labels.unsqueeze(-2).expand(*cost.shape[:-1], gt_labels.shape[-1])
        return torch.gather(cost, -1, idx) * self.weight  # (..., Q, M)


class BBox3DL1Cost:
    """core/bbox/match_costs/match_cost.py:9-33."""

    def __init__(self, weight=1.0):
        self.weight = weight

    def __call__(self, bbox_pred, gt_bboxes):
        return torch.cdist(bbox_pred, gt_bboxes, p=1) * self.weight


class IoU3DCost:
    """core/bbox/match_costs/match_cost.py:92-104."""

    def __init__(self, weight=1.0):
        self.weight = weight

    def __call__(self, bbox_pred, gt_bboxes):
        return (1 - bbox_overlaps_nearest_3d(bbox_pred, gt_bboxes)) * self.weight


MATCH_COSTS = {"FocalLossCost": FocalLossCost, "BBox3DL1Cost": BBox3DL1Cost, "IoU3DCost": IoU3DCost}


def build_match_cost(cfg):
    cfg = dict(cfg)
    t = cfg.pop("type")
    if t not in MATCH_COSTS:
        raise NotImplementedError(f"match cost {t!r} (shipped Uni3DETR configs use FocalLossCost / BBox3DL1Cost / IoU3DCost)")
    return MATCH_COSTS[t](**cfg)


def linear_sum_assignment_device(cost):
    """cost (P, nq, m) CUDA -> (rows (P, k), cols (P, k)) with k = min(nq, m): the pairs of the minimum-cost
    assignment of each block, like scipy.optimize.linear_sum_assignment (rows ascending), by the matcher
    kernel (u3d_hungarian). CUDA only - there is no host matcher in the product."""
    P, nq, m = cost.shape
    from .. import ops
    if m <= nq:
        # the kernel assigns every ROW of its input to a column: rows = ground-truth slots, columns = queries
        q_of_gt = ops.hungarian(cost.transpose(1, 2).float().contiguous()).long()      # (P, m)
        rows, order = torch.sort(q_of_gt, dim=1)
        return rows, order
    gt_of_q = ops.hungarian(cost.float().contiguous()).long()                          # (P, nq)
    return torch.arange(nq, device=cost.device).expand(P, nq), gt_of_q


class HungarianAssigner3D:
    """core/bbox/assigners/hungarian_assigner_3d.py:17-151 for a whole image at once: all decoder layers and
    query groups of the image are matched together (the reference calls `assign` once per layer and scipy once
    per group)."""

    def __init__(self, cls_cost=dict(type="ClassificationCost", weight=1.), reg_cost=dict(type="BBoxL1Cost", weight=1.0),
                 iou_cost=dict(type="IoUCost", weight=0.0), pc_range=None, **kwargs):
        self.cls_cost, self.reg_cost, self.iou_cost = (build_match_cost(c) for c in (cls_cost, reg_cost, iou_cost))
        self.pc_range = pc_range

    @torch.no_grad()
    def cost_matrix(self, bbox_pred, cls_pred, gt_bboxes, gt_labels):
        """(..., Q, code) predictions, (M, 7) gravity-centre boxes -> (..., Q, M) weighted cost (:108-121)."""
        from .head import denormalize_bbox
        normalized_gt = normalize_bbox(gt_bboxes, self.pc_range)
        bboxes3d = denormalize_bbox(bbox_pred, self.pc_range)
        lead = bbox_pred.shape[:-2]
        gt_b = gt_bboxes.expand(*lead, *gt_bboxes.shape)
        cls_cost = self.cls_cost(cls_pred, gt_labels.expand(*lead, gt_labels.shape[0]))
        reg_cost = self.reg_cost(bbox_pred[..., :8], normalized_gt[..., :8].expand(*lead, *normalized_gt[..., :8].shape))
        iou_cost = self.iou_cost(bboxes3d, gt_b)
        return cls_cost + reg_cost + iou_cost

    @torch.no_grad()
    def assign(self, bbox_pred, cls_pred, gt_bboxes, gt_labels, num_query, gt_bboxes_ignore=None, eps=1e-7,
               gt_repeattimes=1):
        """bbox_pred (..., Q, code), cls_pred (..., Q, C) with Q = G * num_query; gt_bboxes (M, 7), gt_labels (M,).
        Returns assigned_gt_inds (..., Q) long: 0 = background, i + 1 = ground truth i (:141-149)."""
        assert gt_bboxes_ignore is None
        lead, Q = bbox_pred.shape[:-2], bbox_pred.shape[-2]
        M = gt_bboxes.shape[0]
        assigned = torch.zeros(*lead, Q, dtype=torch.long, device=bbox_pred.device)
        if M == 0 or Q == 0:
            return assigned
        cost = self.cost_matrix(bbox_pred.float(), cls_pred.float(), gt_bboxes.float(), gt_labels)
        G = Q // num_query
        cost = cost.reshape(-1, num_query, M).repeat(1, 1, gt_repeattimes)              # (P, nq, M * rep), :131
        rows, cols = linear_sum_assignment_device(cost)
        flat = assigned.view(-1, num_query)                                             # (P, nq): P = prod(lead) * G
        flat.scatter_(1, rows.to(flat.device), (cols.to(flat.device) % M) + 1)
        return assigned


def reduce_mean(t):
    """[restated] mmdet reduce_mean: average over the ranks of the default process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    t = t.clone()
    dist.all_reduce(t.div_(dist.get_world_size()), op=dist.ReduceOp.SUM)
    return t
