#!/usr/bin/env python
"""Experiment: do two CUDA-graph instances of the forward, replayed alternately on two streams, overlap enough
(decoder tail of step i with the sparse front of step i+1) to raise throughput over back-to-back replays of one?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from uni3detr_b200 import GraphedForward, synth

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "32"))
model, cfg = synth.build_model("sunrgbd", seed=0)
model = model.to(dev)
model.set_compute_dtype(torch.bfloat16)
nq = cfg["pts_bbox_head"]["num_query"]
host = [torch.from_numpy(synth.make_scene("sunrgbd", i)) for i in range(B)]
rp = torch.rand(B, nq, 3, generator=torch.Generator().manual_seed(1234)).to(dev)
batch = torch.cat(host, 0).to(dev)
gs = []
for i in range(2):
    g = GraphedForward(model, [p.shape[0] for p in host], host[0].shape[1], random_point=rp, postprocess=False)
    g.points.copy_(batch)
    gs.append(g)
torch.cuda.synchronize()
K = 20


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e)


def seq():
    for _ in range(K):
        gs[0].graph.replay()


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def alt():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    for _ in range(K // 2):
        with torch.cuda.stream(s1):
            gs[0].graph.replay()
        with torch.cuda.stream(s2):
            gs[1].graph.replay()
    cur.wait_stream(s1)
    cur.wait_stream(s2)


t_seq = timed(seq)
t_alt = timed(alt)
print(f"B={B} sequential {t_seq / K:.3f} ms/step ({B * K / t_seq * 1e3:.1f} scenes/s)   two graphs on two streams "
      f"{t_alt / K:.3f} ms/step ({B * K / t_alt * 1e3:.1f} scenes/s)")
c0 = float(gs[0].outputs[1].float().sum()); c1 = float(gs[1].outputs[1].float().sum())
print("checksums", c0, c1)
