"""Training side: losses, Hungarian assigner, targets.

CPU (`-m "not gpu"`): the oracle restatement (oracle/train.py) and the product's batched loss expressions
against tests/golden/golden_loss.npz - loss values, their gradients and the assignment produced by the
REFERENCE's own uni3detr_head.py / hungarian_assigner_3d.py / match_cost.py / rdiouloss.py
(tests/golden/make_golden_loss.py). The two CUDA-only pieces of the product (matcher kernel, rotated 3-D IoU
kernel) are replaced by the oracle's here; the GPU tests run the real ones.
GPU: the product loss end to end on the device, the matcher kernel against scipy, the IoU kernel against the
float64 polygon clipper.
"""
import os

import numpy as np
import pytest
import torch

from oracle import train as OT

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_loss as MGL  # noqa: E402  (case table only; nothing under /root/reference is touched)

PCR = MGL.PCR


@pytest.fixture(scope="module")
def golden_loss():
    return dict(np.load(os.path.join(HERE, "golden", "golden_loss.npz")))


def load_case(g, ci, c):
    preds = {k: torch.from_numpy(g[f"c{ci}_{k}"]) for k in ("all_cls_scores", "all_bbox_preds", "all_iou_preds")}
    gts = [torch.from_numpy(g[f"c{ci}_gt{b}"]) for b in range(c["B"])]
    gls = [torch.from_numpy(g[f"c{ci}_gl{b}"]) for b in range(c["B"])]
    return preds, gts, gls


def build_head(c, device="cpu"):
    import projects.mmdet3d_plugin  # noqa: F401
    from uni3detr_b200.compat import HEADS, build_from_cfg
    layer = dict(MGL_LAYER)
    cfg = dict(type="Uni3DETRHead", num_query=c["nq"], num_classes=c["C"], in_channels=256, sync_cls_avg_factor=True,
               with_box_refine=True, as_two_stage=False, code_size=8, gt_repeattimes=c["rep"],
               code_weights=[1.0] * 8,
               transformer=dict(type="Uni3DETRTransformer", decoder=dict(type="Uni3DETRTransformerDecoder", num_layers=1,
                                                                         return_intermediate=True, transformerlayers=layer)),
               bbox_coder=dict(type="NMSFreeCoder", post_center_range=PCR, pc_range=PCR, max_num=10, num_classes=c["C"]),
               loss_cls=dict(type="SoftFocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.5),
               loss_bbox=dict(type="L1Loss", loss_weight=0.25), loss_iou=dict(type="IoU3DLoss", loss_weight=1.2),
               train_cfg=dict(assigner=dict(type="HungarianAssigner3D", cls_cost=dict(type="FocalLossCost", weight=2.0),
                                            reg_cost=dict(type="BBox3DL1Cost", weight=0.25),
                                            iou_cost=dict(type="IoU3DCost", weight=1.2), pc_range=PCR)))
    return build_from_cfg(cfg, HEADS).to(device)


MGL_LAYER = dict(type="BaseTransformerLayer",
                 attn_cfgs=[dict(type="MultiheadAttention", embed_dims=256, num_heads=8, dropout=0.1),
                            dict(type="UniCrossAtten", num_points=1, embed_dims=256, num_sweeps=1)],
                 ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=64, num_fcs=2, ffn_drop=0.1,
                               act_cfg=dict(type="ReLU", inplace=True)),
                 norm_cfg=dict(type="LN"), operation_order=("self_attn", "norm", "cross_attn", "norm", "ffn", "norm"))


@pytest.mark.parametrize("ci", range(len(MGL.CASES)))
def test_oracle_loss_matches_the_reference(golden_loss, ci):
    c = MGL.CASES[ci]
    preds, gts, gls = load_case(golden_loss, ci, c)
    p = {k: v.clone().requires_grad_() for k, v in preds.items()}
    d = OT.loss(p, gts, gls, c["nq"], c["C"], c["rep"])
    for k, v in d.items():
        np.testing.assert_allclose(float(v), float(golden_loss[f"c{ci}_loss_{k}"]), rtol=2e-5, atol=1e-6, err_msg=k)
    sum(d.values()).backward()
    for k in preds:
        np.testing.assert_allclose(p[k].grad.numpy(), golden_loss[f"c{ci}_grad_{k}"], rtol=1e-4, atol=1e-6, err_msg=k)
    for b in range(c["B"]):
        inds = OT.assign(preds["all_bbox_preds"][0, b], preds["all_cls_scores"][0, b], gts[b], gls[b], c["nq"], c["rep"])
        np.testing.assert_array_equal(inds.numpy(), golden_loss[f"c{ci}_assign{b}"])


def _host_matcher(cost):
    from scipy.optimize import linear_sum_assignment
    rows, cols = zip(*(linear_sum_assignment(cost[p].detach().cpu().numpy()) for p in range(cost.shape[0])))
    return torch.from_numpy(np.stack(rows)).to(cost.device), torch.from_numpy(np.stack(cols)).to(cost.device)


@pytest.mark.parametrize("ci", range(len(MGL.CASES)))
def test_product_loss_expressions_match_the_reference_on_cpu(golden_loss, ci, monkeypatch):
    """The product's batched (all layers at once) targets / loss expressions, with the two CUDA-only kernels
    swapped for the oracle's (the product itself refuses CPU tensors there)."""
    from uni3detr_b200.plugin import losses as LS
    monkeypatch.setattr(LS, "linear_sum_assignment_device", _host_matcher)
    monkeypatch.setattr(LS, "bbox_overlaps_3d_aligned", lambda a, b: OT.iou3d_rotated_aligned(a[..., :7], b[..., :7]))
    c = MGL.CASES[ci]
    preds, gts, gls = load_case(golden_loss, ci, c)
    head = build_head(c)
    p = {k: v.clone().requires_grad_() for k, v in preds.items()}
    d = head.loss(gts, gls, p)
    assert set(d) == {k[len(f"c{ci}_loss_"):] for k in golden_loss if k.startswith(f"c{ci}_loss_")}
    for k, v in d.items():
        np.testing.assert_allclose(float(v), float(golden_loss[f"c{ci}_loss_{k}"]), rtol=2e-5, atol=1e-6, err_msg=k)
    sum(d.values()).backward()
    for k in preds:
        np.testing.assert_allclose(p[k].grad.numpy(), golden_loss[f"c{ci}_grad_{k}"], rtol=1e-4, atol=1e-6, err_msg=k)
    # un-normalised sums (the data-parallel step normalises after its single all-reduce): sums / npos == losses
    d2 = head.loss(gts, gls, preds, normalize=False)
    npos = max(float(d2.pop("num_total_pos")), 1.0)
    for k, v in d2.items():
        np.testing.assert_allclose(float(v) / npos, float(golden_loss[f"c{ci}_loss_{k}"]), rtol=2e-5, atol=1e-6)


def test_product_matcher_and_iou_refuse_cpu_tensors():
    from uni3detr_b200._lib import U3DError
    from uni3detr_b200.plugin import losses as LS
    with pytest.raises(U3DError):
        LS.linear_sum_assignment_device(torch.rand(1, 4, 3))
    with pytest.raises(U3DError):
        LS.bbox_overlaps_3d_aligned(torch.rand(3, 7), torch.rand(3, 7))


# ---------------------------------------------------------------------------------- GPU ----
@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,P", [(1, 5, 3), (20, 300, 12), (35, 300, 9), (300, 300, 2), (64, 900, 4)])
def test_hungarian_kernel_vs_scipy(rows, cols, P):
    """u3d_hungarian == scipy.optimize.linear_sum_assignment: same assignment (generic costs: unique optimum)."""
    from scipy.optimize import linear_sum_assignment
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(rows * 1000 + cols)
    cost = torch.randn(P, rows, cols, generator=g)
    got = ops.hungarian(cost.cuda()).cpu().numpy()
    for p in range(P):
        r, c = linear_sum_assignment(cost[p].numpy())
        np.testing.assert_array_equal(got[p], c)
    # repeated columns (gt_repeattimes): ties between copies - equal total cost, each row a distinct column
    cost = torch.randn(P, rows, max(cols // 5, rows), generator=g).repeat(1, 1, 5)[:, :, :max(cols, rows)]
    got = ops.hungarian(cost.cuda()).cpu().numpy()
    for p in range(P):
        r, c = linear_sum_assignment(cost[p].numpy())
        assert len(set(got[p].tolist())) == rows
        np.testing.assert_allclose(cost[p].numpy()[np.arange(rows), got[p]].sum(), cost[p].numpy()[r, c].sum(), rtol=1e-5)


@pytest.mark.gpu
def test_iou3d_aligned_kernel_vs_polygon_clipper():
    from uni3detr_b200 import ops
    g = torch.Generator().manual_seed(4)
    n = 500
    a = torch.cat([torch.rand(n, 3, generator=g) * 2, 0.3 + torch.rand(n, 3, generator=g), (torch.rand(n, 1, generator=g) - 0.5) * 6], 1)
    b = a + torch.cat([torch.randn(n, 3, generator=g) * 0.3, torch.randn(n, 3, generator=g) * 0.1, torch.randn(n, 1, generator=g)], 1)
    b[:, 3:6] = b[:, 3:6].abs() + 0.1
    b[:10] = a[:10]                                                    # identical boxes: IoU 1
    want = OT.iou3d_rotated_aligned(a, b)
    got = ops.iou3d_aligned(a.cuda(), b.cuda()).cpu()
    assert float(want[:10].min()) > 0.9999
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(MGL.CASES)))
def test_product_loss_on_device_matches_the_reference(golden_loss, ci):
    """Uni3DETRHead.loss on the GPU - matcher kernel, IoU kernel, batched targets - vs the reference's values
    and gradients."""
    c = MGL.CASES[ci]
    preds, gts, gls = load_case(golden_loss, ci, c)
    head = build_head(c, "cuda")
    p = {k: v.clone().cuda().requires_grad_() for k, v in preds.items()}
    d = head.loss([t.cuda() for t in gts], [t.cuda() for t in gls], p)
    for k, v in d.items():
        np.testing.assert_allclose(float(v), float(golden_loss[f"c{ci}_loss_{k}"]), rtol=1e-4, atol=1e-5, err_msg=k)
    sum(d.values()).backward()
    for k in preds:
        np.testing.assert_allclose(p[k].grad.cpu().numpy(), golden_loss[f"c{ci}_grad_{k}"], rtol=1e-3, atol=1e-5, err_msg=k)
    for b in range(c["B"]):
        inds = head.assigner.assign(p["all_bbox_preds"][0, b].detach(), p["all_cls_scores"][0, b].detach(), gts[b].cuda(),
                                    gls[b].cuda(), c["nq"], gt_repeattimes=c["rep"])
        np.testing.assert_array_equal(inds.cpu().numpy(), golden_loss[f"c{ci}_assign{b}"])
