"""Golden vectors for Uni3DETRHead.get_bboxes with post_processing type 'nms'
(projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:826-918).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_getbboxes.py
Output (committed): tests/golden/golden_get_bboxes.npz

The reference's OWN `get_bboxes` and `NMSFreeCoder.decode` run from /root/reference: bottom-centre shift, the
per-class loop and its class-major output order, `score_thr` (scalar and per-class list), `num_thr`. Stubbed
third-party pieces: `mmcv.ops.nms3d` [restated: oracle/postproc.py nms3d - rotated-BEV IoU, greedy, score
descending, so the suppression arithmetic itself is NOT pinned] and the `box_type_3d` container (a class
holding `.tensor`).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as MG  # noqa: E402
from oracle import postproc as OP  # noqa: E402


def nms3d(boxes, scores, iou_threshold):                 # [restated] mmcv.ops.nms3d
    keep = OP.nms3d(boxes.numpy().astype(np.float64), scores.numpy().astype(np.float64), iou_threshold)
    return torch.from_numpy(keep)


class Boxes:                                              # stands in for LiDARInstance3DBoxes / DepthInstance3DBoxes
    def __init__(self, tensor, box_dim=7):
        self.tensor = tensor


def make_preds(g, L, B, Q, C, pcr):
    preds = {"all_cls_scores": torch.randn(L, B, Q, C, generator=g) - 0.5,
             "all_bbox_preds": torch.randn(L, B, Q, 8, generator=g) * 0.3,
             "all_iou_preds": torch.randn(L, B, Q, 1, generator=g)}
    p = preds["all_bbox_preds"]
    centres = torch.rand(B, 6, 3, generator=g)
    which = torch.randint(0, 6, (B, Q), generator=g)
    c = torch.gather(centres, 1, which.unsqueeze(-1).expand(-1, -1, 3)) + 0.03 * torch.randn(B, Q, 3, generator=g)
    lo, hi = torch.tensor(pcr[:3]), torch.tensor(pcr[3:])
    c = lo + (0.15 + 0.7 * c) * (hi - lo)
    p[..., 0], p[..., 1], p[..., 4] = c[..., 0], c[..., 1], c[..., 2]
    p[..., 2:4] = torch.log(torch.tensor([0.9, 0.6])) + 0.1 * torch.randn(L, B, Q, 2, generator=g)
    p[..., 5] = torch.log(torch.tensor(0.8)) + 0.1 * torch.randn(L, B, Q, generator=g)
    return preds


def main():
    MG.install_stubs()
    MG.stub("mmcv.ops", nms3d=nms3d)
    MG.load_ref("projects/mmdet3d_plugin/core/bbox/util.py", "projects.mmdet3d_plugin.core.bbox.util")
    coder = MG.load_ref("projects/mmdet3d_plugin/core/bbox/coders/nms_free_coder.py", "ref_coder_gb")
    head = MG.load_ref("projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py", "ref_head_gb")
    pcr = [-3.2, -0.2, -2.0, 3.2, 6.2, 0.56]
    C = 4
    out = {}
    cases = [dict(type="nms", nms_thr=0.5), dict(type="nms", nms_thr=0.3, score_thr=0.2),
             dict(type="nms", nms_thr=0.5, score_thr=[0.1, 0.3, 0.2, 0.25], num_thr=15),
             dict(type="nms", nms_thr=0.2, num_thr=20)]
    for ci, pp in enumerate(cases):
        g = torch.Generator().manual_seed(100 + ci)
        preds = make_preds(g, 3, 2, 60, C, pcr)

        class Shell:
            pass
        hs = Shell()
        hs.bbox_coder = coder.NMSFreeCoder(pc_range=pcr, post_center_range=pcr, max_num=40, alpha=0.2, num_classes=C)
        hs.post_processing, hs.num_classes = pp, C
        metas = [dict(box_type_3d=Boxes), dict(box_type_3d=Boxes)]
        with torch.no_grad():
            res = head.Uni3DETRHead.get_bboxes(hs, {k: v.clone() for k, v in preds.items()}, metas)
        for k, v in preds.items():
            out[f"c{ci}_{k}"] = v.numpy()
        for i, (b, s, l) in enumerate(res):
            out[f"c{ci}_s{i}_bboxes"] = b.tensor.numpy()
            out[f"c{ci}_s{i}_scores"] = s.numpy()
            out[f"c{ci}_s{i}_labels"] = np.asarray(l)
            print(f"case {ci} scene {i}: {len(s)} boxes kept")
    np.savez_compressed(os.path.join(HERE, "golden_get_bboxes.npz"), **out)


if __name__ == "__main__":
    main()
