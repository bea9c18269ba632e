"""Golden vectors for the training side: Uni3DETRHead.loss / loss_single / get_targets / _get_target_single
(projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:510-793), HungarianAssigner3D.assign
(core/bbox/assigners/hungarian_assigner_3d.py:53-151), BBox3DL1Cost / IoU3DCost (core/bbox/match_costs/
match_cost.py), SoftFocalLoss / IoU3DLoss (models/losses/rdiouloss.py), normalize_bbox (core/bbox/util.py).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_loss.py
Output (committed): tests/golden/golden_loss.npz

All of the files above run UNMODIFIED from /root/reference, with the real scipy.optimize.linear_sum_assignment.
Stubbed third-party pieces, [restated] from mmdet 2.x / mmdet3d 1.0.0rc5 / mmcv and therefore NOT pinned by
this fixture: mmdet `multi_apply`, `reduce_mean` (single process: identity), `AssignResult`, `PseudoSampler`,
`FocalLossCost`, `L1Loss`, `weighted_loss` / `weight_reduce_loss`; mmdet3d `bbox_overlaps_nearest_3d`
(nearest-BEV 2-D IoU) and `bbox_overlaps_3d` (rotated 3-D IoU; the dense (N,N) matrix the reference takes the
diagonal of is built from the aligned restatement on the diagonal only).
"""
import functools
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as MG  # noqa: E402
from uni3detr_b200.plugin import losses as LS  # noqa: E402  (the [restated] third-party pieces live there)


# ---- [restated] mmdet pieces
def multi_apply(func, *args, **kwargs):
    pfunc = functools.partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


class AssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels


class SamplingResult:
    pass


class PseudoSampler:
    def sample(self, assign_result, bboxes, gt_bboxes, **kwargs):
        r = SamplingResult()
        r.pos_inds = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).squeeze(-1).unique()
        r.neg_inds = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).squeeze(-1).unique()
        r.pos_assigned_gt_inds = assign_result.gt_inds[r.pos_inds] - 1
        r.pos_gt_bboxes = gt_bboxes[r.pos_assigned_gt_inds.long(), :] if gt_bboxes.numel() else gt_bboxes.view(-1, gt_bboxes.shape[-1])
        return r


def weighted_loss(loss_func):
    @functools.wraps(loss_func)
    def wrapper(pred, target, weight=None, reduction="mean", avg_factor=None, **kwargs):
        return LS.weight_reduce_loss(loss_func(pred, target, **kwargs), weight, reduction, avg_factor)
    return wrapper


def bbox_overlaps_nearest_3d(b1, b2, mode="iou", is_aligned=False, coordinate="lidar"):
    return LS.bbox_overlaps_nearest_3d(b1[..., :7], b2[..., :7], is_aligned)


def bbox_overlaps_3d(b1, b2, mode="iou", coordinate="camera"):
    assert coordinate == "lidar" and b1.shape[0] == b2.shape[0]
    return torch.diag(LS.bbox_overlaps_3d_aligned(b1, b2))          # only the diagonal is read (:690)


MATCH = MG._Reg()
MATCH.d["FocalLossCost"] = LS.FocalLossCost
LOSSES = MG._Reg()
LOSSES.d["L1Loss"] = LS.L1Loss
ASSIGNERS = MG._Reg()


def install():
    MG.install_stubs()
    MG.stub("mmcv.ops", nms3d=None, nms_bev=None, diff_iou_rotated_3d=None)
    MG.stub("mmcv.cnn", Linear=nn.Linear, bias_init_with_prob=lambda p: float(-np.log((1 - p) / p)))
    MG.stub("mmdet.core", multi_apply=multi_apply, reduce_mean=lambda t: t)
    MG.stub("mmdet.core.bbox.builder", BBOX_ASSIGNERS=ASSIGNERS, BBOX_CODERS=MG.BBOX_CODERS)
    MG.stub("mmdet.core.bbox.assigners", AssignResult=AssignResult, BaseAssigner=object)
    MG.stub("mmdet.core.bbox.match_costs", build_match_cost=MATCH.build)
    MG.stub("mmdet.core.bbox.match_costs.builder", MATCH_COST=MATCH)
    MG.stub("mmdet.models", HEADS=MG.HEADS, LOSSES=LOSSES, DETECTORS=MG.DETECTORS, BACKBONES=MG.BACKBONES, NECKS=MG.NECKS)
    MG.stub("mmdet.models.losses.utils", weighted_loss=weighted_loss, weight_reduce_loss=LS.weight_reduce_loss)
    MG.stub("mmdet3d.core.bbox", AxisAlignedBboxOverlaps3D=object)
    MG.stub("mmdet3d.core.bbox.iou_calculators.iou3d_calculator", bbox_overlaps_3d=bbox_overlaps_3d,
            bbox_overlaps_nearest_3d=bbox_overlaps_nearest_3d)
    MG.stub("mmdet3d.models.builder", build_loss=LOSSES.build, MIDDLE_ENCODERS=MG.MIDDLE)
    MG.stub("mmdet3d.core.bbox.coders", build_bbox_coder=MG.BBOX_CODERS.build)
    util = MG.load_ref("projects/mmdet3d_plugin/core/bbox/util.py", "projects.mmdet3d_plugin.core.bbox.util")
    MG.load_ref("projects/mmdet3d_plugin/core/bbox/match_costs/match_cost.py", "ref_match_cost")
    asg = MG.load_ref("projects/mmdet3d_plugin/core/bbox/assigners/hungarian_assigner_3d.py", "ref_assigner")
    MG.load_ref("projects/mmdet3d_plugin/models/losses/rdiouloss.py", "ref_losses")
    head = MG.load_ref("projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py", "ref_head_loss")
    return util, asg, head


PCR = [-3.2, -0.2, -2.0, 3.2, 6.2, 0.56]
CASES = [   # (seed, L, B, groups, nq, classes, gts per image, gt_repeattimes)
    dict(seed=1, L=2, B=2, G=3, nq=40, C=10, n_gt=[4, 6], rep=1),
    dict(seed=2, L=3, B=2, G=3, nq=60, C=3, n_gt=[5, 0], rep=5),       # KITTI-style repeats, one empty image
    dict(seed=3, L=1, B=1, G=4, nq=30, C=10, n_gt=[7], rep=2),
]


def make_case(c):
    g = torch.Generator().manual_seed(c["seed"])
    lo, hi = torch.tensor(PCR[:3]), torch.tensor(PCR[3:])
    Q = c["G"] * c["nq"]
    gts, gls = [], []
    for n in c["n_gt"]:
        ctr = lo + (0.15 + 0.7 * torch.rand(n, 3, generator=g)) * (hi - lo)
        dims = 0.3 + torch.rand(n, 3, generator=g)
        yaw = (torch.rand(n, 1, generator=g) - 0.5) * 6.0
        gts.append(torch.cat([ctr, dims, yaw], 1))
        gls.append(torch.randint(0, c["C"], (n,), generator=g))
    box = torch.randn(c["L"], c["B"], Q, 8, generator=g) * 0.3
    ctr = lo + (0.1 + 0.8 * torch.rand(c["L"], c["B"], Q, 3, generator=g)) * (hi - lo)
    box[..., 0], box[..., 1], box[..., 4] = ctr[..., 0], ctr[..., 1], ctr[..., 2]
    # a few queries sit close to ground-truth boxes so the IoU terms are exercised
    for b, gt in enumerate(gts):
        for i in range(gt.shape[0]):
            q = (i * 7 + 3) % Q
            box[:, b, q, 0], box[:, b, q, 1], box[:, b, q, 4] = gt[i, 0] + 0.05, gt[i, 1] - 0.04, gt[i, 2] + 0.03
            box[:, b, q, 2], box[:, b, q, 3], box[:, b, q, 5] = gt[i, 4].log(), gt[i, 3].log(), gt[i, 5].log()
    cls = torch.randn(c["L"], c["B"], Q, c["C"], generator=g) - 1.0
    iou = torch.randn(c["L"], c["B"], Q, 1, generator=g)
    return dict(all_cls_scores=cls, all_bbox_preds=box, all_iou_preds=iou), gts, gls


class GT:
    """stands in for mmdet3d's box container (uni3detr_head.py:759-761 reads these two attributes)"""

    def __init__(self, t):
        self.gravity_center, self.tensor = t[:, :3], t


def main():
    util, asg, head = install()
    out = {}
    for ci, c in enumerate(CASES):
        preds, gts, gls = make_case(c)

        class Shell(nn.Module):
            pass
        hs = Shell()
        hs.num_query, hs.num_classes, hs.cls_out_channels = c["nq"], c["C"], c["C"]
        hs.gt_repeattimes, hs.bg_cls_weight, hs.sync_cls_avg_factor = c["rep"], 0, True
        hs.pc_range = PCR
        hs.code_weights = nn.Parameter(torch.tensor([1.0] * 8), requires_grad=False)
        hs.assigner = asg.HungarianAssigner3D(cls_cost=dict(type="FocalLossCost", weight=2.0),
                                              reg_cost=dict(type="BBox3DL1Cost", weight=0.25),
                                              iou_cost=dict(type="IoU3DCost", weight=1.2), pc_range=PCR)
        hs.sampler = PseudoSampler()
        hs.loss_cls = LOSSES.build(dict(type="SoftFocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.5))
        hs.loss_bbox = LOSSES.build(dict(type="L1Loss", loss_weight=0.25))
        hs.loss_iou = LOSSES.build(dict(type="IoU3DLoss", loss_weight=1.2))
        for name in ("loss_single", "get_targets", "_get_target_single", "loss"):
            setattr(hs, name, getattr(head.Uni3DETRHead, name).__get__(hs))
        hs._bbox_to_loss = head.Uni3DETRHead._bbox_to_loss
        p = {k: v.clone().requires_grad_() for k, v in preds.items()}
        losses = head.Uni3DETRHead.loss(hs, [GT(t) for t in gts], gls, p)
        total = sum(losses.values())
        total.backward()
        for k, v in preds.items():
            out[f"c{ci}_{k}"] = v.numpy()
            out[f"c{ci}_grad_{k}"] = p[k].grad.numpy()
        for b, (t, l) in enumerate(zip(gts, gls)):
            out[f"c{ci}_gt{b}"], out[f"c{ci}_gl{b}"] = t.numpy(), l.numpy()
        for k, v in losses.items():
            out[f"c{ci}_loss_{k}"] = np.float32(v.item())
        # the assignment itself, layer 0
        for b in range(c["B"]):
            r = hs.assigner.assign(preds["all_bbox_preds"][0, b], preds["all_cls_scores"][0, b], gts[b], gls[b], c["nq"],
                                   None, gt_repeattimes=c["rep"])
            out[f"c{ci}_assign{b}"] = r.gt_inds.numpy()
        print(f"case {ci}:", {k: round(float(v), 5) for k, v in losses.items()})
    np.savez_compressed(os.path.join(HERE, "golden_loss.npz"), **out)


if __name__ == "__main__":
    main()
