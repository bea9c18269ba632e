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
    """ONE all-reduce for the whole job: every rank writes its (scenes, checksum, elapsed) into its own
    3 lanes of a (world, 3) float64 vector that is zero elsewhere, the vector is SUM-reduced once, and
    each rank reads the scene / checksum totals and the slowest rank's time from it.
    Returns (total scenes, max elapsed, checksum)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    vec = torch.zeros((world, 3), dtype=torch.float64, device=device)
    vec[rank] = torch.tensor([float(scenes_done), float(checksum), float(elapsed_s)], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return float(vec[:, 0].sum()), float(vec[:, 2].max()), float(vec[:, 1].sum())
