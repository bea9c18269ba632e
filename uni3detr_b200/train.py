"""Data-parallel training step: whole scenes sharded per GPU, ONE all-reduce.

Reference: extra_tools/train.py:247-254 -> mmdet3d `train_model` -> MMDistributedDataParallel + the mmcv runner's
OptimizerHook (config `optimizer = AdamW(lr, weight_decay=0.01)`, `grad_clip max_norm=10`,
projects/configs/uni3detr/uni3detr_sunrgbd.py:233-235). That path all-reduces the gradients in DDP buckets AND,
inside the loss, all-reduces two scalars per decoder layer (`reduce_mean` of `cls_avg_factor` and `num_total_pos`,
uni3detr_head.py:660-662,680-681) before the backward pass can start.

Here every gradient lives in one flat fp32 buffer (`p.grad` are views into it, autograd accumulates in place) whose
last element carries the rank's positive count. The loss is back-propagated UN-normalised
(`Uni3DETR.forward_train(normalize=False)`): every term of the reference's loss is divided by the same number -
the rank-averaged count of matched queries, identical for all decoder layers because every ground-truth box is always
matched and the classification background weight is 0 - so that division commutes with the gradient sum. One
`all_reduce(SUM)` of the flat buffer (NCCL over NVLink / NVSwitch) then yields both the summed gradients and the global
count; a single in-place scale by 1 / (world * max(mean count, 1)) gives exactly the gradient DDP's averaging + the
reference's normalisation produce.
"""
import torch
import torch.distributed as dist


class DataParallelTrainer:
    def __init__(self, model, lr=2e-4, weight_decay=0.01, max_grad_norm=10.0, optimizer=None):
        self.model = model
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n + 1, dtype=torch.float32, device=dev)     # gradients | positive count
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        if self.world > 1:                                                  # identical start on every rank
            init = torch.cat([p.detach().reshape(-1).float() for p in self.params])
            dist.broadcast(init, src=0)
            off = 0
            with torch.no_grad():
                for p in self.params:
                    p.copy_(init[off:off + p.numel()].view_as(p))
                    off += p.numel()
        self.max_grad_norm = max_grad_norm
        self.optimizer = optimizer or torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)
        self.collectives_per_step = 1 if self.world > 1 else 0

    def step(self, points, gt_bboxes_3d, gt_labels_3d, img_metas=None):
        """One optimisation step on this rank's scenes. Returns the loss dict normalised like the reference's
        (each value divided by the global mean positive count)."""
        self.flat.zero_()                                                   # grads stay views of the bucket
        losses = self.model.forward_train(points=points, img_metas=img_metas, gt_bboxes_3d=gt_bboxes_3d,
                                          gt_labels_3d=gt_labels_3d, normalize=False)
        npos = losses.pop("num_total_pos")
        torch.stack(list(losses.values())).sum().backward()
        self.flat[-1] = npos
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)               # THE collective of the step
        denom = torch.clamp(self.flat[-1] / self.world, min=1.0)
        self.flat[:-1].mul_(1.0 / (denom * self.world))
        if self.max_grad_norm:
            norm = self.flat[:-1].norm()
            self.flat[:-1].mul_(torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0))
        self.optimizer.step()
        return {k: v.detach() / denom for k, v in losses.items()}
