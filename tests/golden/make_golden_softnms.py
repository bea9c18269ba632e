"""Golden vectors for the `soft_nms` option of Uni3DETRHead.get_bboxes (uni3detr_head.py:795-823, :862-867).

Run in the build container (needs /root/reference):  python tests/golden/make_golden_softnms.py
Output (committed): tests/golden/golden_soft_nms.npz

The reference's OWN `soft_nms` method and `get_bboxes` run from /root/reference. Stubbed third-party piece:
mmdet3d `bbox_overlaps_3d(coordinate='lidar')` [restated here independently of the product: rotated-BEV
intersection by the oracle's clipper x height overlap / union volume - the IoU arithmetic itself is NOT pinned].
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
import make_golden_getbboxes as GB  # noqa: E402
from oracle import postproc as OP  # noqa: E402


def bbox_overlaps_3d(a, b, mode="iou", coordinate="camera"):      # [restated] mmdet3d iou3d_calculator
    assert coordinate == "lidar" and mode == "iou"
    a, b = a.numpy().astype(np.float64), b.numpy().astype(np.float64)
    out = np.zeros((len(a), len(b)))
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            h = min(x[2] + x[5], y[2] + y[5]) - max(x[2], y[2])
            if h <= 0:
                continue
            inter = OP.poly_clip_area(OP.rect_corners(x), OP.rect_corners(y)) * h
            out[i, j] = inter / max(x[3] * x[4] * x[5] + y[3] * y[4] * y[5] - inter, 1e-8)
    return torch.from_numpy(out).float()


def main():
    MG.install_stubs()
    MG.stub("mmcv.ops", nms3d=GB.nms3d)
    MG.stub("mmdet3d.core.bbox.iou_calculators.iou3d_calculator", bbox_overlaps_3d=bbox_overlaps_3d)
    MG.load_ref("projects/mmdet3d_plugin/core/bbox/util.py", "projects.mmdet3d_plugin.core.bbox.util")
    coder = MG.load_ref("projects/mmdet3d_plugin/core/bbox/coders/nms_free_coder.py", "ref_coder_sn")
    head = MG.load_ref("projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py", "ref_head_sn")
    pcr = [-3.2, -0.2, -2.0, 3.2, 6.2, 0.56]
    C = 4
    out = {}
    cases = [dict(type="soft_nms", gaussian_sigma=0.3, prune_threshold=1e-2),
             dict(type="soft_nms", gaussian_sigma=0.5, prune_threshold=1e-3, score_thr=0.15, num_thr=25)]
    for ci, pp in enumerate(cases):
        g = torch.Generator().manual_seed(200 + ci)
        preds = GB.make_preds(g, 3, 2, 60, C, pcr)

        class Shell:
            soft_nms = head.Uni3DETRHead.soft_nms
        hs = Shell()
        hs.bbox_coder = coder.NMSFreeCoder(pc_range=pcr, post_center_range=pcr, max_num=40, alpha=0.2, num_classes=C)
        hs.post_processing, hs.num_classes = pp, C
        metas = [dict(box_type_3d=GB.Boxes), dict(box_type_3d=GB.Boxes)]
        with torch.no_grad():
            res = head.Uni3DETRHead.get_bboxes(hs, {k: v.clone() for k, v in preds.items()}, metas)
        for k, v in preds.items():
            out[f"c{ci}_{k}"] = v.numpy()
        for i, (b, s, l) in enumerate(res):
            out[f"c{ci}_s{i}_bboxes"] = b.tensor.numpy()
            out[f"c{ci}_s{i}_scores"] = s.numpy()
            out[f"c{ci}_s{i}_labels"] = np.asarray(l)
            print(f"case {ci} scene {i}: {len(s)} boxes")
    np.savez_compressed(os.path.join(HERE, "golden_soft_nms.npz"), **out)


if __name__ == "__main__":
    main()
