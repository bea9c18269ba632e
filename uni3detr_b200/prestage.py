"""Device input pre-stage built from the reference's `test_pipeline` dicts (round 1; validated on hardware in round 2).

`PointsPreStage(cfg.test_pipeline)` reads the `LoadPointsFromFile`, `PointsRangeFilter` and `PointSample`
entries of the reference config (same `type=` names and keys: coord_type, load_dim, use_dim, shift_height,
point_cloud_range, num_points; uni3detr_sunrgbd.py:175-191) and runs them on the device through
libu3d (csrc/points.cu), returning the (points, offsets) pair the voxelizer consumes - no DataContainer,
no per-sample numpy round trip. Image / formatting entries (`DefaultFormatBundle3D`, `Collect3D`,
`MultiScaleFlipAug3D` with its identity test-time transforms) carry no point arithmetic and are skipped;
`LoadPointsFromMultiSweeps` (nuScenes) is not covered yet.

Not on the benchmarked path (the bench starts from points already in the model's format): see csrc/points.cu.
"""
import numpy as np
import torch

from . import ops


def _flatten(pipeline):
    for t in pipeline:
        yield t
        if isinstance(t, dict) and "transforms" in t:
            yield from _flatten(t["transforms"])


class PointsPreStage:
    def __init__(self, pipeline, seed=0):
        self.load_dim, self.use_dim, self.shift_height = None, None, False
        self.pc_range, self.num_points = None, None
        for t in _flatten(pipeline):
            ty = t.get("type")
            if ty == "LoadPointsFromFile":
                self.load_dim = int(t.get("load_dim", 6))
                ud = t.get("use_dim", [0, 1, 2])
                self.use_dim = list(range(ud)) if isinstance(ud, int) else [int(u) for u in ud]
                self.shift_height = bool(t.get("shift_height", False))
            elif ty == "PointsRangeFilter":
                self.pc_range = [float(v) for v in t["point_cloud_range"]]
            elif ty == "PointSample":
                self.num_points = int(t["num_points"])
            elif ty == "LoadPointsFromMultiSweeps":
                raise NotImplementedError("PointsPreStage: multi-sweep loading is a next row (SURVEY.md 8f)")
        if self.load_dim is None:
            raise ValueError("PointsPreStage: the pipeline has no LoadPointsFromFile entry")
        # the reference draws PointSample's indices from numpy's global legacy stream: a RandomState seeded the
        # same way reproduces its choices (tests/golden/golden_point_sample.npz)
        self.rng = np.random.RandomState(seed)

    @property
    def channels(self):
        return len(self.use_dim) + (1 if self.shift_height else 0)

    def draw_choices(self, kept):
        """PointSample's host-side index draw (numpy, like the reference), one array per scene."""
        out = []
        for n in kept:
            out.append(self.rng.choice(int(n), self.num_points, replace=int(n) < self.num_points))
        return out

    @torch.no_grad()
    def __call__(self, raw_list, device="cuda"):
        """raw_list: per scene, a float32 array/tensor of the .bin contents (any shape, load_dim columns).
        Returns (points (N, C) f32 on `device`, offsets (B+1) int32 on `device`)."""
        B = len(raw_list)
        raws = [torch.as_tensor(np.asarray(r, np.float32)).reshape(-1, self.load_dim) for r in raw_list]
        lens = [r.shape[0] for r in raws]
        off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
        raw = torch.cat(raws, 0).pin_memory().to(device, non_blocking=True)
        pts, out_off, _ = ops.points_prepare(raw, off.to(device), B, self.use_dim, self.shift_height, self.pc_range)
        if self.num_points is None:
            return pts, out_off
        o = out_off.cpu().numpy()               # PointSample draws on the host: it needs the kept counts
        kept = np.diff(o)
        choices = np.concatenate([c + o[b] for b, c in enumerate(self.draw_choices(kept))]).astype(np.int32)
        sampled = ops.points_gather(pts, torch.from_numpy(choices).to(device))
        new_off = torch.arange(B + 1, dtype=torch.int32, device=device) * self.num_points
        return sampled, new_off
