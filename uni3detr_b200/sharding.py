"""Multi-GPU plumbing: whole scenes are the unit of sharding.

The forward needs no communication (no op mixes scenes in eval mode), so W ranks each run
their own scenes - the same partition mmdet's DistributedSampler gives the reference
(extra_tools/dist_test.sh -> test.py:217-222) - and the job ends with ONE all-reduce of a small
metrics vector, replacing multi_gpu_test's tmpdir/pickle gather. Works on any torch.distributed
backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def scene_indices(total, rank, world):
    """Scenes of `rank`: {i : i mod world == rank} (DistributedSampler order, no padding)."""
    return list(range(rank, total, world))


def reduce_metrics(scenes_done, elapsed_s, checksum, device="cpu"):
    """One all-reduce: SUM of scenes and checksum, MAX of elapsed time (packed as a second
    MAX-reduced lane of the same launch group). Returns (total scenes, max elapsed, checksum)."""
    vec = torch.tensor([float(scenes_done), float(checksum)], dtype=torch.float64, device=device)
    tmax = torch.tensor([float(elapsed_s)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        work = [dist.all_reduce(vec, op=dist.ReduceOp.SUM, async_op=True),
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX, async_op=True)]
        for w in work:
            w.wait()
    return float(vec[0]), float(tmax[0]), float(vec[1])
