"""Uni3DETRHead + NMSFreeCoder drop-ins.

References: projects/mmdet3d_plugin/models/dense_heads/uni3detr_head.py:311-508 (forward),
projects/mmdet3d_plugin/core/bbox/coders/nms_free_coder.py:9-136,
projects/mmdet3d_plugin/core/bbox/util.py:44-80 (denormalize_bbox, mmdet3d>=1.0 branch).
Training side (`loss`, targets, Hungarian assignment): uni3detr_head.py:510-793 via plugin/losses.py.
"""
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..compat import BBOX_CODERS, HEADS, TRANSFORMER, build_from_cfg
from .transformer import _ln, _run_mlp, _wb, inverse_sigmoid


def denormalize_bbox(normalized_bboxes, pc_range=None):
    """core/bbox/util.py:44-80 with __mmdet3d_version__ >= 1.0 (the pinned v1.0.0rc5)."""
    rot = torch.atan2(normalized_bboxes[..., 6:7], normalized_bboxes[..., 7:8])
    rot = -rot - np.pi / 2
    cx, cy, cz = normalized_bboxes[..., 0:1], normalized_bboxes[..., 1:2], normalized_bboxes[..., 4:5]
    w = normalized_bboxes[..., 2:3].exp()
    l = normalized_bboxes[..., 3:4].exp()
    h = normalized_bboxes[..., 5:6].exp()
    if normalized_bboxes.size(-1) > 8:
        vx, vy = normalized_bboxes[..., 8:9], normalized_bboxes[..., 9:10]
        return torch.cat([cx, cy, cz, l, w, h, rot, vx, vy], dim=-1)
    return torch.cat([cx, cy, cz, l, w, h, rot], dim=-1)


@BBOX_CODERS.register_module()
class NMSFreeCoder:
    def __init__(self, pc_range, voxel_size=None, post_center_range=None, max_num=100,
                 score_threshold=None, alpha=0.5, num_classes=10):
        self.pc_range = pc_range
        self.voxel_size = voxel_size
        self.post_center_range = post_center_range
        self.max_num = max_num
        self.score_threshold = score_threshold
        self.num_classes = num_classes
        self.alpha = alpha
        self._pcr = {}

    def _range(self, like):
        """post_center_range as a tensor on `like`'s device, created once per device (so a captured
        CUDA graph of the forward contains no host-to-device copy)."""
        key = (like.device, like.dtype)
        if key not in self._pcr:
            self._pcr[key] = like.new_tensor(self.post_center_range)
        return self._pcr[key]

    def encode(self):
        pass

    def decode_single(self, cls_scores, bbox_preds, all_iou_preds):
        cls_scores = cls_scores.sigmoid()
        scores, indexs = cls_scores.view(-1).topk(self.max_num)
        labels = indexs % self.num_classes
        bbox_index = torch.div(indexs, self.num_classes, rounding_mode="floor")
        bbox_preds = bbox_preds[bbox_index]
        final_box_preds = denormalize_bbox(bbox_preds, self.pc_range)
        final_ious = all_iou_preds.sigmoid()[bbox_index]
        if self.post_center_range is None:
            raise NotImplementedError("Need to reorganize output as a batch, only support "
                                      "post_center_range is not None for now!")
        pcr = self._range(scores)
        mask = (final_box_preds[..., :3] >= pcr[:3]).all(1)
        mask &= (final_box_preds[..., :3] <= pcr[3:]).all(1)
        if self.score_threshold:
            mask &= scores > self.score_threshold
        ious = final_ious[mask].reshape(-1)
        s = scores[mask]
        return {"bboxes": final_box_preds[mask],
                "scores": s ** self.alpha * ious ** (1 - self.alpha),
                "labels": labels[mask], "ious": ious}

    @torch.no_grad()
    def decode_fixed(self, preds_dicts):
        """Device-resident form of :meth:`decode` for the batched hot path: same arithmetic
        (nms_free_coder.py:57-97,121-123) over the whole batch at once, but fixed-size outputs and
        no host synchronisation - rows the reference would drop (centre outside
        post_center_range, score <= score_threshold) are flagged in ``mask`` instead of removed.
        Returns bboxes (B,max_num,7|9), scores (B,max_num), labels (B,max_num) int64, mask bool."""
        cls = torch.mean(preds_dicts["all_cls_scores"][1:].float(), 0)
        box = torch.mean(preds_dicts["all_bbox_preds"][1:].float(), 0)
        iou = torch.mean(preds_dicts["all_iou_preds"][1:].float(), 0)
        B, Q, C = cls.shape
        scores, idx = cls.sigmoid().view(B, Q * C).topk(self.max_num, dim=1)
        labels = idx % self.num_classes
        qi = torch.div(idx, self.num_classes, rounding_mode="floor")
        boxes = denormalize_bbox(torch.gather(box, 1, qi.unsqueeze(-1).expand(-1, -1, box.shape[-1])),
                                 self.pc_range)
        ious = torch.gather(iou.sigmoid().squeeze(-1), 1, qi)
        pcr = self._range(scores)
        mask = (boxes[..., :3] >= pcr[:3]).all(-1) & (boxes[..., :3] <= pcr[3:]).all(-1)
        if self.score_threshold:
            mask &= scores > self.score_threshold
        final = scores ** self.alpha * ious ** (1 - self.alpha)
        return boxes, final, labels, mask

    def decode(self, preds_dicts):
        cls = torch.mean(preds_dicts["all_cls_scores"][1:].float(), 0)
        box = torch.mean(preds_dicts["all_bbox_preds"][1:].float(), 0)
        iou = torch.mean(preds_dicts["all_iou_preds"][1:].float(), 0)
        return [self.decode_single(cls[i], box[i], iou[i]) for i in range(cls.size(0))]


@HEADS.register_module()
class Uni3DETRHead(nn.Module):
    def __init__(self, num_classes, in_channels, num_query=100, num_reg_fcs=2, transformer=None,
                 sync_cls_avg_factor=False, positional_encoding=None, loss_cls=None, loss_bbox=None,
                 loss_iou=None, train_cfg=None, test_cfg=None, init_cfg=None,
                 with_box_refine=False, as_two_stage=False, bbox_coder=None, num_cls_fcs=2,
                 code_weights=None, post_processing=None, gt_repeattimes=1, code_size=None, **kwargs):
        super().__init__()
        if as_two_stage:
            raise NotImplementedError("as_two_stage is not used by any Uni3DETR config")
        self.with_box_refine, self.as_two_stage = with_box_refine, as_two_stage
        self.code_size = code_size if code_size is not None else 10
        cw = code_weights if code_weights is not None else \
            [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2]
        self.bbox_coder = build_from_cfg(dict(bbox_coder), BBOX_CODERS)
        self.pc_range = self.bbox_coder.pc_range
        self.num_cls_fcs = num_cls_fcs - 1
        self.num_classes, self.in_channels, self.num_query = num_classes, in_channels, num_query
        self.num_reg_fcs = num_reg_fcs
        self.sync_cls_avg_factor = sync_cls_avg_factor
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.loss_cfgs = dict(loss_cls=loss_cls, loss_bbox=loss_bbox, loss_iou=loss_iou)
        # training side (uni3detr_head.py:340-356 + mmdet DETRHead.__init__): losses, assigner, pseudo sampler
        from . import losses as LS
        self.loss_cls = LS.build_loss(loss_cls) if loss_cls and loss_cls.get("type") in LS.LOSSES else None
        self.loss_bbox = LS.build_loss(loss_bbox) if loss_bbox and loss_bbox.get("type") in LS.LOSSES else None
        self.loss_iou = LS.build_loss(loss_iou) if loss_iou and loss_iou.get("type") in LS.LOSSES else None
        self.bg_cls_weight = 0          # mmdet DETRHead: sigmoid classification has no background class weight
        self.assigner = None
        if train_cfg and train_cfg.get("assigner"):
            a = dict(train_cfg["assigner"])
            if a.pop("type") != "HungarianAssigner3D":
                raise NotImplementedError("only HungarianAssigner3D (all shipped configs)")
            self.assigner = LS.HungarianAssigner3D(**a)
        use_sigmoid = bool((loss_cls or {}).get("use_sigmoid", False))
        self.cls_out_channels = num_classes if use_sigmoid else num_classes + 1
        self.use_sigmoid_cls = use_sigmoid
        self.transformer = build_from_cfg(dict(transformer), TRANSFORMER)
        self.embed_dims = self.transformer.embed_dims
        self.code_weights = nn.Parameter(torch.tensor(cw), requires_grad=False)
        self.fp16_enabled = False
        self.post_processing = post_processing
        self.gt_repeattimes = gt_repeattimes
        self.compute_dtype = torch.float32
        self._plan = None
        self._rng = None  # optional torch.Generator for the test-time random query group
        self._init_layers()
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())

    def _init_layers(self):
        E = self.embed_dims
        cls_branch = []
        for _ in range(self.num_reg_fcs):
            cls_branch += [nn.Linear(E, E), nn.LayerNorm(E), nn.ReLU(inplace=True)]
        cls_branch.append(nn.Linear(E, self.cls_out_channels))
        fc_cls = nn.Sequential(*cls_branch)
        reg_branch = []
        for _ in range(self.num_reg_fcs):
            reg_branch += [nn.Linear(E, E), nn.ReLU()]
        reg_branch.append(nn.Linear(E, self.code_size))
        reg_branch = nn.Sequential(*reg_branch)
        iou_branch = []
        for _ in range(self.num_reg_fcs):
            iou_branch += [nn.Linear(E, E), nn.ReLU()]
        iou_branch.append(nn.Linear(E, 1))
        iou_branch = nn.Sequential(*iou_branch)
        num_pred = self.transformer.decoder.num_layers
        if self.with_box_refine:
            clone = lambda m: nn.ModuleList([copy.deepcopy(m) for _ in range(num_pred)])
            self.cls_branches, self.reg_branches, self.iou_branches = \
                clone(fc_cls), clone(reg_branch), clone(iou_branch)
        else:
            self.cls_branches = nn.ModuleList([fc_cls for _ in range(num_pred)])
            self.reg_branches = nn.ModuleList([reg_branch for _ in range(num_pred)])
            self.iou_branches = nn.ModuleList([iou_branch for _ in range(num_pred)])
        self.tgt_embed = nn.Embedding(self.num_query * 2, E)
        self.refpoint_embed = nn.Embedding(self.num_query, 3)

    def init_weights(self):
        self.transformer.init_weights()
        if self.use_sigmoid_cls:
            bias_init = float(-math.log((1 - 0.01) / 0.01))
            for m in self.cls_branches:
                nn.init.constant_(m[-1].bias, bias_init)

    def invalidate(self):
        self._plan = None

    def train(self, mode=True):
        self._plan = None
        return super().train(mode)

    def set_compute_dtype(self, dtype):
        self.compute_dtype = dtype
        self.transformer.decoder.compute_dtype = dtype
        self.invalidate()
        self.transformer.decoder.invalidate()

    @torch.no_grad()
    def prepare(self):
        dt = self.compute_dtype
        lin = lambda br: [_wb(m, dt) for m in br if isinstance(m, nn.Linear)]
        cls = []
        for br in self.cls_branches:
            lns = [(m.weight.detach().to(dt), m.bias.detach().to(dt), m.eps)
                   for m in br if isinstance(m, nn.LayerNorm)]
            cls.append(dict(lin=lin(br), ln=lns))
        self._plan = dict(dtype=dt, cls=cls, reg=[lin(b) for b in self.reg_branches],
                          iou=[lin(b) for b in self.iou_branches], tc=None)
        if self.transformer.decoder.uses_tc():
            from .. import ops
            PL = ops.PackedLinear
            tc_cls = []
            for br in self.cls_branches:
                mods = list(br)
                tc_cls.append([PL(m.weight, m.bias, (mods[i + 1].weight, mods[i + 1].bias, mods[i + 1].eps)
                                  if i + 1 < len(mods) and isinstance(mods[i + 1], nn.LayerNorm) else None)
                               for i, m in enumerate(mods) if isinstance(m, nn.Linear)])
            plin = lambda br: [PL(m.weight, m.bias) for m in br if isinstance(m, nn.Linear)]
            self._plan["tc"] = dict(cls=tc_cls, reg=[plin(b) for b in self.reg_branches],
                                    iou=[plin(b) for b in self.iou_branches])
        return self._plan

    def build_queries(self, fpsbpts, train_mode, random_point=None):
        """uni3detr_head.py:437-449 - mixed queries: learned + 2 FPS groups (+ random at test)."""
        nq, bs = self.num_query, fpsbpts.shape[0]
        tgt = self.tgt_embed.weight
        refanchor = self.refpoint_embed.weight
        if train_mode:
            tgt = torch.cat([tgt[0:nq], tgt[nq:], tgt[nq:]])
            refs = torch.cat([refanchor.unsqueeze(0).expand(bs, -1, -1), inverse_sigmoid(fpsbpts)], 1)
        else:
            if random_point is None:
                random_point = torch.rand(fpsbpts.shape, device=fpsbpts.device,
                                          generator=self._rng)[:, :nq, :]
            tgt = torch.cat([tgt[0:nq], tgt[nq:], tgt[nq:], tgt[nq:]])
            refs = torch.cat([refanchor.unsqueeze(0).expand(bs, -1, -1), inverse_sigmoid(fpsbpts),
                              inverse_sigmoid(random_point)], 1)
        return torch.cat([tgt.unsqueeze(0).expand(bs, -1, -1), refs], -1)

    def forward_train(self, pts_feats, img_metas, fpsbpts):
        """uni3detr_head.py:422-508 in training mode under autograd (fp32): 3 query groups (learned + the two
        FPS groups, :437-443), decoder, per-level branches and box assembly as torch ops on the modules."""
        query_embeds = self.build_queries(fpsbpts.float(), True)
        if pts_feats.dim() == 5:
            pts_feats = pts_feats.unsqueeze(1)
        hs, init_reference, inter_references = self.transformer(
            pts_feats, query_embeds, self.num_query,
            reg_branches=self.reg_branches if self.with_box_refine else None, img_metas=img_metas)
        hs = hs.permute(0, 2, 1, 3)
        pc = self.pc_range
        classes, coords, ious = [], [], []
        for lvl in range(hs.shape[0]):
            reference = inverse_sigmoid(init_reference if lvl == 0 else inter_references[lvl - 1])
            tmp = self.reg_branches[lvl](hs[lvl])
            xy = (tmp[..., 0:2] + reference[..., 0:2]).sigmoid()
            z = (tmp[..., 4:5] + reference[..., 2:3]).sigmoid()
            cx = xy[..., 0:1] * (pc[3] - pc[0]) + pc[0]
            cy = xy[..., 1:2] * (pc[4] - pc[1]) + pc[1]
            cz = z * (pc[5] - pc[2]) + pc[2]
            coords.append(torch.cat([cx, cy, tmp[..., 2:4], cz, tmp[..., 5:]], dim=-1))
            classes.append(self.cls_branches[lvl](hs[lvl]))
            ious.append(self.iou_branches[lvl](hs[lvl]))
        return {"all_cls_scores": torch.stack(classes), "all_bbox_preds": torch.stack(coords),
                "all_iou_preds": torch.stack(ious)}

    def forward(self, pts_feats, img_metas, fpsbpts, random_point=None):
        """pts_feats (B,C,D,H,W); fpsbpts (B,2nq,3) in [0,1]. Returns the reference's dict of
        all_cls_scores (L,B,Q,cls), all_bbox_preds (L,B,Q,code), all_iou_preds (L,B,Q,1), fp32."""
        if self.training:
            return self.forward_train(pts_feats, img_metas, fpsbpts)
        with torch.no_grad():
            return self._forward_eval(pts_feats, img_metas, fpsbpts, random_point)

    def _forward_eval(self, pts_feats, img_metas, fpsbpts, random_point=None):
        p = self._plan
        if p is None or p["dtype"] != self.compute_dtype:
            p = self.prepare()
        dt = p["dtype"]
        query_embeds = self.build_queries(fpsbpts.float(), pts_feats.requires_grad, random_point)
        if pts_feats.dim() == 5:
            pts_feats = pts_feats.unsqueeze(1)
        tc = p.get("tc")
        reg_plans = (tc["reg"] if tc is not None else p["reg"]) if self.with_box_refine else None
        hs, init_reference, inter_references = self.transformer(
            pts_feats, query_embeds, self.num_query, reg_plans=reg_plans, img_metas=img_metas)
        hs = hs.permute(0, 2, 1, 3)  # (L,B,Q,E)
        E = self.embed_dims
        pc = self.pc_range
        if tc is not None and hs.is_cuda and hs.dtype == torch.bfloat16:
            return self._forward_heads_tc(tc, hs)
        classes, coords, ious = [], [], []
        for lvl in range(hs.shape[0]):
            reference = init_reference if lvl == 0 else inter_references[lvl - 1]
            reference = inverse_sigmoid(reference)
            x = hs[lvl]
            c = x
            cp = p["cls"][lvl]
            c = c.reshape(-1, E)
            for i, (w, b) in enumerate(cp["lin"][:-1]):
                c = _ln(F.linear(c, w, b), cp["ln"][i], relu=True)
            c = c.reshape(x.shape)
            outputs_class = F.linear(c, *cp["lin"][-1]).float()
            tmp = _run_mlp(x, p["reg"][lvl]).float()
            outputs_iou = _run_mlp(x, p["iou"][lvl]).float()
            assert reference.shape[-1] == 3
            xy = (tmp[..., 0:2] + reference[..., 0:2]).sigmoid()
            z = (tmp[..., 4:5] + reference[..., 2:3]).sigmoid()
            cx = xy[..., 0:1] * (pc[3] - pc[0]) + pc[0]
            cy = xy[..., 1:2] * (pc[4] - pc[1]) + pc[1]
            cz = z * (pc[5] - pc[2]) + pc[2]
            coords.append(torch.cat([cx, cy, tmp[..., 2:4], cz, tmp[..., 5:]], dim=-1))
            classes.append(outputs_class)
            ious.append(outputs_iou)
        return {"all_cls_scores": torch.stack(classes), "all_bbox_preds": torch.stack(coords),
                "all_iou_preds": torch.stack(ious)}

    def _forward_heads_tc(self, tc, hs):
        """bf16 serving path of the branches (uni3detr_head.py:456-496): Linear-LN-ReLU / Linear-ReLU stacks
        as ops.linear_tc launches writing straight into the stacked fp32 outputs; the reg branch of level
        l is the one the decoder already evaluated for its reference refinement (same weights, same
        input: uni3detr_transformer.py:194-196), so its raw output is reused; one box-assembly kernel."""
        from .. import ops
        L_ = ops.linear_tc
        nL, B, Q, E = hs.shape
        R = B * Q
        dev = hs.device
        cls_out = torch.empty((nL, R, self.cls_out_channels), dtype=torch.float32, device=dev)
        iou_out = torch.empty((nL, R, 1), dtype=torch.float32, device=dev)
        box_out = torch.empty((nL, R, self.code_size), dtype=torch.float32, device=dev)
        tmps = self.transformer.decoder.last_reg_tmp if self.with_box_refine else None
        for lvl in range(nL):
            x = hs[lvl].reshape(R, E)
            c = x
            for lin in tc["cls"][lvl][:-1]:
                c = L_(c, lin, ln=True, relu_out=True)
            L_(c, tc["cls"][lvl][-1], out_f32=True, out=cls_out[lvl])
            i = L_(L_(x, tc["iou"][lvl][0], relu=True), tc["iou"][lvl][1], relu=True)
            L_(i, tc["iou"][lvl][2], out_f32=True, out=iou_out[lvl])
            if tmps is not None:
                tmp = tmps[lvl]
            else:
                r = tc["reg"][lvl]
                tmp = L_(L_(L_(x, r[0], relu=True), r[1], relu=True), r[2], out_f32=True)
            # the transformer returns sigmoid(ref) and the reference takes inverse_sigmoid of it again
            # (:129 -> :475); box_assemble does that round trip from the logits level l started from
            ops.box_assemble(tmp, self.transformer.decoder.last_ref_logits[lvl], self.pc_range, out=box_out[lvl])
        return {"all_cls_scores": cls_out.view(nL, B, Q, -1), "all_bbox_preds": box_out.view(nL, B, Q, -1),
                "all_iou_preds": iou_out.view(nL, B, Q, 1)}

    # ------------------------------------------------------------------ training ---
    def get_targets_batched(self, all_cls, all_box, gt_bboxes_list, gt_labels_list):
        """uni3detr_head.py:510-621 (`_get_target_single` + `get_targets`) for every decoder layer at once.
        all_cls (L,B,Q,C), all_box (L,B,Q,code); gt boxes (n_i,7) gravity-centre, labels (n_i,).
        Returns labels (L,B,Q) long (num_classes = background), bbox_targets (L,B,Q,7), pos mask (L,B,Q) bool."""
        L, B, Q, _ = all_cls.shape
        dev = all_cls.device
        labels = torch.full((L, B, Q), self.num_classes, dtype=torch.long, device=dev)
        targets = torch.zeros((L, B, Q, 7), dtype=torch.float32, device=dev)
        pos = torch.zeros((L, B, Q), dtype=torch.bool, device=dev)
        for b in range(B):                                   # one matcher launch per image: L x G problems
            gb, gl = gt_bboxes_list[b].to(dev).float(), gt_labels_list[b].to(dev).long()
            if gb.shape[0] == 0:
                continue
            # reference quirk kept (uni3detr_head.py:542-543): `self.gt_repeattimes` is passed positionally and
            # lands in assign()'s unused `eps` parameter, so the loss never repeats the ground-truth columns
            inds = self.assigner.assign(all_box[:, b].detach(), all_cls[:, b].detach(), gb, gl, self.num_query,
                                        None, self.gt_repeattimes)                     # (L, Q)
            m = inds > 0
            gi = (inds - 1).clamp(min=0)
            pos[:, b] = m
            labels[:, b] = torch.where(m, gl[gi], labels[:, b])
            targets[:, b] = torch.where(m.unsqueeze(-1), gb[gi][..., :7], targets[:, b])
        return labels, targets, pos

    def loss_layers(self, all_cls, all_box, all_iou, gt_bboxes_list, gt_labels_list, normalize=True):
        """`loss_single` (uni3detr_head.py:623-698) for all L decoder layers as batched tensor expressions.
        Returns (loss_cls, loss_bbox, loss_iou, loss_iou_pred) each of shape (L,), and num_total_pos.
        normalize=False returns the un-normalised sums (x loss weights) instead: the data-parallel train step
        divides by the rank-averaged positive count AFTER its single gradient all-reduce."""
        from . import losses as LS
        L, B, Q, C = all_cls.shape
        all_cls, all_box, all_iou = all_cls.float(), all_box.float(), all_iou.float()
        labels, bbox_targets, pos = self.get_targets_batched(all_cls, all_box, gt_bboxes_list, gt_labels_list)
        num_total_pos = int(pos[0].sum())                   # identical for every layer: all gts are matched
        num_total_neg = B * Q - num_total_pos
        if normalize:
            cls_avg_factor = num_total_pos * 1.0 + num_total_neg * self.bg_cls_weight
            if self.sync_cls_avg_factor:
                cls_avg_factor = float(LS.reduce_mean(all_cls.new_tensor([cls_avg_factor])))
            cls_avg_factor = max(cls_avg_factor, 1)
            npos = float(torch.clamp(LS.reduce_mean(all_cls.new_tensor([float(num_total_pos)])), min=1))
        else:
            cls_avg_factor = npos = 1.0
        N = B * Q
        cls = all_cls.reshape(L, N, C)
        box = all_box.reshape(L, N, -1)
        tgt = bbox_targets.reshape(L, N, 7)
        w = pos.reshape(L, N).float()
        norm_tgt = LS.normalize_bbox(tgt, self.pc_range)
        boxes3d = denormalize_bbox(box, self.pc_range)
        iou3d = LS.bbox_overlaps_nearest_3d(boxes3d, tgt, is_aligned=True)
        pc, tc = LS.bbox_to_corners_aa(boxes3d), LS.bbox_to_corners_aa(tgt)
        z1, z2, z3, z4 = pc[..., 2], pc[..., 5], tc[..., 2], tc[..., 5]
        iou_z = torch.max(torch.min(z2, z4) - torch.max(z1, z3), torch.zeros_like(z1)) / (torch.max(z2, z4) - torch.min(z1, z3))
        iou3d_dec = (iou3d + iou_z) / 2
        isnotnan = torch.isfinite(norm_tgt).all(dim=-1)
        bw = w.unsqueeze(-1) * self.code_weights[:box.shape[-1]]
        iou_true = LS.bbox_overlaps_3d_aligned(boxes3d.reshape(-1, 7), tgt.reshape(-1, 7)).reshape(L, N)
        out = [[], [], [], []]
        for l in range(L):                                   # L small reductions; everything above is batched
            lc = self.loss_cls(cls[l], [labels[l].reshape(-1), iou3d_dec[l]], torch.ones(N, device=cls.device),
                               avg_factor=cls_avg_factor)
            nn_ = isnotnan[l]
            lb = self.loss_bbox(box[l][nn_, :10], norm_tgt[l][nn_, :10], bw[l][nn_, :10], avg_factor=npos)
            li = self.loss_iou(boxes3d[l][nn_, :10], tgt[l][nn_, :10], bw[l][nn_, :10], avg_factor=npos)
            li = li + torch.sum((1 - iou_z[l][nn_]) * bw[l][nn_, 0]) / npos
            lp = torch.sum(F.binary_cross_entropy_with_logits(all_iou[l].reshape(-1), iou_true[l], reduction="none")
                           * bw[l][nn_, 0]) / npos * 1.2
            for o, v in zip(out, (lc, lb, li, lp)):
                o.append(v)
        return tuple(torch.stack(o) for o in out), num_total_pos

    def loss_single(self, cls_scores, bbox_preds, iou_preds, gt_bboxes_list, gt_labels_list, gt_bboxes_ignore_list=None):
        """Reference API (uni3detr_head.py:623-698): one decoder layer, (B,Q,.) tensors."""
        (lc, lb, li, lp), _ = self.loss_layers(cls_scores[None], bbox_preds[None], iou_preds[None], gt_bboxes_list,
                                               gt_labels_list)
        return lc[0], lb[0], li[0], lp[0]

    @staticmethod
    def _gt_tensor(gt_bboxes):
        """(gravity_center, dims, yaw) rows from an mmdet3d box container or a plain (n,7+) tensor whose
        first three columns already are the gravity centre (uni3detr_head.py:759-761)."""
        if hasattr(gt_bboxes, "gravity_center"):
            return torch.cat((gt_bboxes.gravity_center, gt_bboxes.tensor[:, 3:]), dim=1)
        return gt_bboxes

    def loss(self, gt_bboxes_list, gt_labels_list, preds_dicts, gt_bboxes_ignore=None, normalize=True):
        """uni3detr_head.py:716-793: dict of the last layer's four losses + `d{i}.` entries of the others."""
        assert gt_bboxes_ignore is None, f"{self.__class__.__name__} only supports for gt_bboxes_ignore setting to None."
        if self.assigner is None or self.loss_cls is None:
            raise RuntimeError("Uni3DETRHead.loss needs train_cfg.assigner and the loss configs")
        if preds_dicts["all_bbox_preds"].shape[-1] != 8:
            raise NotImplementedError("loss: code_size 8 (the reference truncates targets to 7 box values, :554)")
        dev = preds_dicts["all_cls_scores"].device
        gts = [self._gt_tensor(g).to(dev) for g in gt_bboxes_list]
        (lc, lb, li, lp), npos = self.loss_layers(preds_dicts["all_cls_scores"], preds_dicts["all_bbox_preds"],
                                                  preds_dicts["all_iou_preds"], gts, gt_labels_list, normalize)
        d = {"loss_cls": lc[-1], "loss_bbox": lb[-1], "loss_iou": li[-1], "loss_iou_pred": lp[-1]}
        for i in range(len(lc) - 1):
            d[f"d{i}.loss_cls"], d[f"d{i}.loss_bbox"] = lc[i], lb[i]
            d[f"d{i}.loss_iou"], d[f"d{i}.loss_iou_pred"] = li[i], lp[i]
        if not normalize:
            d["num_total_pos"] = torch.tensor(float(npos), device=dev)
        return d

    def _score_thr(self, thr, like):
        """Per-class score_thr list as a device tensor, cached per (device, dtype): building it inside a
        step would be a pageable host-to-device copy (a sync, and illegal during graph capture)."""
        key = (like.device, like.dtype, tuple(float(t) for t in thr))
        cache = self.__dict__.setdefault("_thr_cache", {})
        if key not in cache:
            cache[key] = torch.tensor(key[2], dtype=like.dtype, device=like.device)
        return cache[key]

    @torch.no_grad()
    def postprocess_fixed(self, preds_dicts):
        """Device-resident get_bboxes (uni3detr_head.py:827-918) for post_processing None / 'nms':
        NMSFreeCoder top-k, gravity-centre -> bottom-centre shift (:842), per-class nms3d for every
        scene in one launch pair (:847-871), score_thr (:895-908), num_thr (:910-914). Fixed-size
        outputs, no host synchronisation: returns boxes (B,M,7|9), scores (B,M), labels (B,M) int64
        and keep (B,M) bool; the kept rows, in order, are the reference's result."""
        from .. import ops
        boxes, scores, labels, keep = self.bbox_coder.decode_fixed(preds_dicts)
        boxes = boxes.clone()
        boxes[..., 2] = boxes[..., 2] - boxes[..., 5] * 0.5
        pp = self.post_processing
        if pp is None:
            return boxes, scores, labels, keep
        if pp["type"] != "nms":
            raise NotImplementedError(f"post_processing type {pp['type']!r}: box_merging / soft_nms are greedy "
                                      "sequential loops that run host-side through get_bboxes (like the reference)")
        # class-major, score-descending order inside a class (the reference's output order);
        # rows the coder dropped sort to the end
        B, M = scores.shape
        o1 = torch.sort(scores, dim=1, descending=True, stable=True)[1]
        lab = torch.where(keep, labels, torch.full_like(labels, self.num_classes))
        o2 = torch.sort(torch.gather(lab, 1, o1), dim=1, stable=True)[1]
        order = torch.gather(o1, 1, o2)
        boxes = torch.gather(boxes, 1, order.unsqueeze(-1).expand(-1, -1, boxes.shape[-1]))
        scores, labels, keep = (torch.gather(t, 1, order) for t in (scores, labels, keep))
        keep = keep & ops.nms3d_bev(boxes[..., :7].float().contiguous(), labels.int().contiguous(), keep,
                                    pp["nms_thr"])
        if "score_thr" in pp:
            thr = pp["score_thr"]
            if isinstance(thr, (list, tuple)):
                assert len(thr) == self.num_classes
                keep = keep & (scores > self._score_thr(thr, scores)[labels.clamp(max=self.num_classes - 1)])
            else:
                keep = keep & (scores > thr)
        if "num_thr" in pp:
            k = min(int(pp["num_thr"]), M)
            masked = torch.where(keep, scores, scores.new_full((), -1.0))
            top, order = torch.sort(masked, dim=1, descending=True, stable=True)
            order, top = order[:, :k], top[:, :k]
            boxes = torch.gather(boxes, 1, order.unsqueeze(-1).expand(-1, -1, boxes.shape[-1]))
            scores, labels = torch.gather(scores, 1, order), torch.gather(labels, 1, order)
            keep = top >= 0
        return boxes, scores, labels, keep

    @torch.no_grad()
    def _get_bboxes_host(self, preds_dicts, img_metas):
        """post_processing types whose core is a greedy, data-dependent sequential loop, run on the host:
        'box_merging' (uni3detr_kitti_3classes.py:115-117; uni3detr_head.py:881-892) - exactly like the
        reference, which calls `.cpu().numpy()` here - the same-class median merge of plugin/box_merging.py;
        'soft_nms' (the commented alternative of every config; uni3detr_head.py:795-823, :862-867) - per class,
        Gaussian score decay by 3-D IoU (plugin/soft_nms.py), class-major output like the 'nms' branch.
        Device-resident decode + bottom-centre shift before, `score_thr` (scalar or per-class list) and
        `num_thr` (:895-914) after."""
        from . import box_merging as BM
        from . import soft_nms as SN
        pp = self.post_processing
        boxes, scores, labels, keep = self.bbox_coder.decode_fixed(preds_dicts)
        boxes = boxes.clone()
        boxes[..., 2] = boxes[..., 2] - boxes[..., 5] * 0.5
        if pp["type"] == "box_merging" and boxes.shape[-1] != 7:
            raise NotImplementedError("box_merging: the reference's corner routine takes 7-value boxes only")
        dev = boxes.device
        ret = []
        for i in range(boxes.shape[0]):
            k = keep[i]
            l_np, b_np = labels[i][k].cpu().numpy(), boxes[i][k].float().cpu().numpy()
            s_np = scores[i][k].float().cpu().numpy()
            if pp["type"] == "box_merging":
                cl, bx, sc, _ = BM.nms_boxes_3d_merge_only(l_np, b_np, s_np, overlapped_thres=0.1)
            else:
                ob, os_, ol = [], [], []
                for j in range(self.num_classes):
                    ind = l_np == j
                    if not ind.any():
                        continue
                    sel, soft = SN.soft_nms(b_np[ind][:, :7], s_np[ind], pp["gaussian_sigma"], pp["prune_threshold"])
                    ob.append(b_np[ind][sel])
                    os_.append(soft.astype(np.float32))
                    ol.extend([j] * len(sel))
                bx = np.concatenate(ob) if ob else np.zeros((0, b_np.shape[1]), np.float32)
                sc = np.concatenate(os_) if os_ else np.zeros(0, np.float32)
                cl = np.asarray(ol, np.int64)
            if "score_thr" in pp:
                thr = pp["score_thr"]
                if isinstance(thr, (list, tuple)):
                    assert len(thr) == self.num_classes
                    ind = np.zeros(len(sc), bool)
                    for j in range(self.num_classes):
                        ind |= (cl == j) & (sc > thr[j])
                else:
                    ind = sc > thr
                cl, bx, sc = cl[ind], bx[ind], sc[ind]
            if "num_thr" in pp:
                ind = np.argsort(-sc)[: pp["num_thr"]]
                cl, bx, sc = cl[ind], bx[ind], sc[ind]
            bboxes = torch.from_numpy(np.ascontiguousarray(bx)).to(dev)
            meta = img_metas[i] if img_metas is not None and i < len(img_metas) else {}
            box_type = meta.get("box_type_3d") if isinstance(meta, dict) else None
            if box_type is not None:
                bboxes = box_type(bboxes, bboxes.shape[-1])
            ret.append([bboxes, torch.from_numpy(np.ascontiguousarray(sc)).to(dev),
                        torch.from_numpy(np.ascontiguousarray(cl)).to(dev)])
        return ret

    @torch.no_grad()
    def get_bboxes(self, preds_dicts, img_metas, rescale=False):
        """Reference API (uni3detr_head.py:827-918): list over scenes of [bboxes, scores, labels].
        post_processing None / 'nms' run on the device (:meth:`postprocess_fixed`); the only host
        round trip is the final compaction to exact-size tensors."""
        pp = self.post_processing
        if pp is not None and pp["type"] in ("box_merging", "soft_nms"):
            return self._get_bboxes_host(preds_dicts, img_metas)
        boxes, scores, labels, keep = self.postprocess_fixed(preds_dicts)
        ret = []
        for i in range(boxes.shape[0]):
            k = keep[i]
            bboxes = boxes[i][k]
            meta = img_metas[i] if img_metas is not None and i < len(img_metas) else {}
            box_type = meta.get("box_type_3d") if isinstance(meta, dict) else None
            if box_type is not None:
                bboxes = box_type(bboxes, bboxes.shape[-1])
            ret.append([bboxes, scores[i][k], labels[i][k]])
        return ret

