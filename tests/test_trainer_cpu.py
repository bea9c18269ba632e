"""world_size-2 gloo test of the data-parallel train step's single flat all-reduce (uni3detr_b200/train.py):
gradients + loss normaliser travel in one collective and equal what per-parameter DDP averaging + the
reference's `reduce_mean` normalisation (uni3detr_head.py:660-662,680-681) would give. Host logic only: a tiny
stand-in model with the `forward_train(normalize=False)` contract (the real model needs a GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a, self.b = nn.Linear(4, 8), nn.Linear(8, 1)

    def forward_train(self, points=None, gt_bboxes_3d=None, gt_labels_3d=None, img_metas=None, normalize=True):
        x, npos = points
        y = self.b(torch.relu(self.a(x))).pow(2)
        return {"loss_a": y.sum(), "loss_b": 0.5 * y.abs().sum(), "num_total_pos": torch.tensor(float(npos))}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(rank):
    g = torch.Generator().manual_seed(10 + rank)
    return torch.randn(6 + rank, 4, generator=g), 3 + 2 * rank


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uni3detr_b200.train import DataParallelTrainer
    calls = []
    real = dist.all_reduce

    def counting(*a, **k):
        calls.append(1)
        return real(*a, **k)
    dist.all_reduce = counting
    model = Toy()
    tr = DataParallelTrainer(model, max_grad_norm=0, optimizer=torch.optim.SGD(model.parameters(), lr=0.0))
    losses = tr.step(_data(rank), None, None)
    q.put((rank, len(calls), tr.flat[:-1].tolist(), {k: float(v) for k, v in losses.items()}))   # plain lists: no fd passing
    dist.destroy_process_group()


def test_one_allreduce_equals_ddp_average_with_reference_normaliser():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # expected: loss_r = sums_r / mean_r(npos); DDP averages the per-rank gradients
    mean_pos = sum(_data(r)[1] for r in range(world)) / world
    grads = []
    for r in range(world):
        m = Toy()
        d = m.forward_train(points=_data(r))
        d.pop("num_total_pos")
        (sum(d.values()) / mean_pos).backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
    want = sum(grads) / world
    for rank, n_calls, flat, losses in res:
        assert n_calls == 1                       # exactly one collective per step
        torch.testing.assert_close(torch.tensor(flat), want, rtol=1e-5, atol=1e-6)
