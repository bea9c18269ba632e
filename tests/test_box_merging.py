"""`box_merging` post-processing (KITTI config) vs golden vectors produced by the REFERENCE's own
bbox_merging.py (tests/golden/make_golden_postproc.py: only shapely is stubbed). CPU: the step is host-side
numpy in the reference (uni3detr_head.py:881-892) and here."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_box_merging.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_box_merging_matches_the_reference_file(gold, case):
    from uni3detr_b200.plugin import box_merging as BM
    g = {k[len(f"c{case}_"):]: v for k, v in gold.items() if k.startswith(f"c{case}_")}
    cl, bx, sc, idx = BM.nms_boxes_3d_merge_only(g["in_labels"], g["in_boxes"], g["in_scores"], overlapped_thres=0.1)
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(cl, g["labels"])
    np.testing.assert_array_equal(sc, g["scores"])
    np.testing.assert_allclose(bx, g["boxes"], rtol=0, atol=1e-6)


def test_corners_follow_the_reference_quirk():
    """A LiDAR box fed to the camera-frame corner routine: BEV polygon in (x, z), height along -y."""
    from uni3detr_b200.plugin import box_merging as BM
    c = BM.boxes_3d_to_corners(np.array([[1.0, 2.0, 3.0, 4.0, 0.5, 2.0, 0.0]]))[0]
    np.testing.assert_allclose(c[0], [3.0, 2.0, 4.0])      # (+l/2, 0, +w/2) + centre
    np.testing.assert_allclose(c[6], [-1.0, 1.5, 2.0])     # (-l/2, -h, -w/2) + centre
    half_turn = BM.boxes_3d_to_corners(np.array([[0, 0, 0, 4.0, 0.5, 2.0, np.pi / 2]]))[0]
    np.testing.assert_allclose(half_turn[0], [1.0, 0.0, -2.0], atol=1e-12)


def test_overlap_closed_forms():
    from uni3detr_b200.plugin import box_merging as BM
    a = BM.boxes_3d_to_corners(np.array([[0, 0, 0, 2.0, 1.0, 2.0, 0.0]]))[0]
    same = BM.overlapped_boxes_3d_fast_poly(a, a[None])
    np.testing.assert_allclose(same, [1.0], atol=1e-6)                       # identical boxes: IoU 1
    shifted = BM.boxes_3d_to_corners(np.array([[1.0, 0, 0, 2.0, 1.0, 2.0, 0.0]]))[0]
    np.testing.assert_allclose(BM.overlapped_boxes_3d_fast_poly(a, shifted[None]), [1.0 / 3.0], atol=1e-6)
    far = BM.boxes_3d_to_corners(np.array([[10.0, 0, 0, 2.0, 1.0, 2.0, 0.3]]))[0]
    assert BM.overlapped_boxes_3d_fast_poly(a, far[None])[0] == 0.0


def test_get_bboxes_box_merging_path_on_the_kitti_head(model_cfgs):
    """Uni3DETRHead.get_bboxes with the KITTI config's post_processing (box_merging + per-class score_thr):
    the batched decode_fixed glue must equal the reference's per-scene flow (decode -> shift -> merge ->
    thresholds), composed here from the golden-tested pieces. Runs on CPU tensors: the step is host-side."""
    import projects.mmdet3d_plugin  # noqa: F401
    from uni3detr_b200.compat import HEADS, build_from_cfg
    from uni3detr_b200.plugin import box_merging as BM
    cfg = dict(model_cfgs["kitti"]["pts_bbox_head"])
    assert cfg["post_processing"]["type"] == "box_merging"
    head = build_from_cfg(cfg, HEADS).eval()
    g = torch.Generator().manual_seed(5)
    L, B, Q, C = 3, 2, 400, head.num_classes
    preds = {"all_cls_scores": torch.randn(L, B, Q, C, generator=g) - 1.0,
             "all_bbox_preds": torch.randn(L, B, Q, 8, generator=g) * 0.3,
             "all_iou_preds": torch.randn(L, B, Q, 1, generator=g)}
    # plausible boxes: normalised centres inside the range, log sizes near a car, unit (sin, cos)
    p = preds["all_bbox_preds"]
    pcr = head.bbox_coder.pc_range
    p[..., 0] = torch.rand(L, B, Q, generator=g) * (pcr[3] - pcr[0]) * 0.3 + pcr[0] + 5
    p[..., 1] = torch.rand(L, B, Q, generator=g) * (pcr[4] - pcr[1]) * 0.2 - 8
    p[..., 4] = torch.rand(L, B, Q, generator=g) * 1.0 - 1.5
    p[..., 2:4] = torch.tensor([1.3, 0.5]) + 0.05 * torch.randn(L, B, Q, 2, generator=g)
    p[..., 5] = 0.4 + 0.05 * torch.randn(L, B, Q, generator=g)
    out = head.get_bboxes(preds, [{}, {}])
    assert len(out) == B
    thr = cfg["post_processing"]["score_thr"]
    for i, dec in enumerate(head.bbox_coder.decode(preds)):
        b = dec["bboxes"].clone()
        b[:, 2] = b[:, 2] - b[:, 5] * 0.5
        cl, bx, sc, _ = BM.nms_boxes_3d_merge_only(dec["labels"].numpy(), b.numpy(), dec["scores"].numpy(), 0.1)
        ind = np.zeros(len(sc), bool)
        for j in range(C):
            ind |= (cl == j) & (sc > thr[j])
        bboxes, scores, labels = out[i]
        np.testing.assert_array_equal(labels.numpy(), cl[ind])
        np.testing.assert_allclose(scores.numpy(), sc[ind], rtol=0, atol=1e-6)
        np.testing.assert_allclose(bboxes.numpy(), bx[ind], rtol=0, atol=1e-5)
        assert len(sc) < len(dec["scores"]) or len(dec["scores"]) <= 1      # something was merged


SOFT_CASES = [dict(type="soft_nms", gaussian_sigma=0.3, prune_threshold=1e-2),
              dict(type="soft_nms", gaussian_sigma=0.5, prune_threshold=1e-3, score_thr=0.15, num_thr=25)]


@pytest.mark.parametrize("case", [0, 1])
def test_get_bboxes_soft_nms_matches_the_reference_method(case, model_cfgs):
    """Uni3DETRHead.get_bboxes with post_processing 'soft_nms' vs golden vectors produced by the REFERENCE's own
    get_bboxes + soft_nms (tests/golden/make_golden_softnms.py; mmdet3d bbox_overlaps_3d stubbed by an
    independent float64 restatement). CPU tensors: the loop is host-side here."""
    import copy
    import projects.mmdet3d_plugin  # noqa: F401
    from uni3detr_b200.compat import HEADS, build_from_cfg
    g = dict(np.load(os.path.join(os.path.dirname(GOLDEN), "golden_soft_nms.npz")))
    pcr = [-3.2, -0.2, -2.0, 3.2, 6.2, 0.56]
    cfg = copy.deepcopy(model_cfgs["sunrgbd"]["pts_bbox_head"])
    cfg.update(num_classes=4, post_processing=SOFT_CASES[case])
    cfg["bbox_coder"].update(num_classes=4, max_num=40, alpha=0.2, pc_range=pcr, post_center_range=pcr)
    head = build_from_cfg(cfg, HEADS).eval()
    preds = {k: torch.from_numpy(g[f"c{case}_{k}"]) for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds")}
    out = head.get_bboxes(preds, [{}, {}])
    for i, (b, s, l) in enumerate(out):
        np.testing.assert_array_equal(l.numpy(), g[f"c{case}_s{i}_labels"])
        np.testing.assert_allclose(s.numpy(), g[f"c{case}_s{i}_scores"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(b.numpy(), g[f"c{case}_s{i}_bboxes"], rtol=0, atol=1e-5)
