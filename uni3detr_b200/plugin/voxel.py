"""Voxel layer / voxel encoders named by the reference configs.

``pts_voxel_layer`` dicts are built into :class:`Voxelization` (mmcv.ops.Voxelization API,
used by MVXTwoStageDetector.voxelize -> detectors/uni3detr.py:148,158); ``HardSimpleVFE`` /
``DynamicSimpleVFE`` are the mmdet3d voxel encoders the configs name
(uni3detr_sunrgbd.py:31, uni3detr_scannet_large.py:31). All of them run on
libu3d_b200 (u3d_voxelize_hard / u3d_voxelize_dynamic); the detector's fused path calls the
batched op once per batch instead of once per sample.
"""
import torch
from torch import nn

from .. import ops
from ..compat import VOXEL_ENCODERS


def grid_size_zyx(point_cloud_range, voxel_size):
    """mmcv: grid = round((range[3:] - range[:3]) / voxel_size) as (x,y,z); returned (z,y,x)."""
    r = torch.tensor(point_cloud_range, dtype=torch.float32)
    v = torch.tensor(voxel_size, dtype=torch.float32)
    g = torch.round((r[3:] - r[:3]) / v).long().tolist()
    return (int(g[2]), int(g[1]), int(g[0]))


class Voxelization(nn.Module):
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000,
                 deterministic=True):
        super().__init__()
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.max_num_points = int(max_num_points)
        self.max_voxels = tuple(max_voxels) if isinstance(max_voxels, (tuple, list)) \
            else (max_voxels, max_voxels)
        self.deterministic = deterministic
        self.grid_zyx = grid_size_zyx(point_cloud_range, voxel_size)

    @property
    def dynamic(self):
        return self.max_num_points == -1

    def current_max_voxels(self):
        return int(self.max_voxels[0] if self.training else self.max_voxels[1])

    def forward(self, points):
        """Per-sample API of mmcv.ops.Voxelization: hard -> (voxels, coors(zyx), num_points);
        dynamic -> per-point coors (N,3) zyx with -1 rows for out-of-range points."""
        points = points.contiguous().float()
        off = torch.tensor([0, points.shape[0]], dtype=torch.int32, device=points.device)
        if self.dynamic:
            v = ops.voxelize_dynamic(points, off, 1, self.point_cloud_range, self.voxel_size,
                                     self.grid_zyx)
            return v.pt_coors[:points.shape[0], 1:]
        v = ops.voxelize_hard(points, off, 1, self.point_cloud_range, self.voxel_size,
                              self.grid_zyx, self.max_num_points, self.current_max_voxels(),
                              deterministic=self.deterministic, want_voxels=True)
        m = int(v.scene_rows[-1].item())  # the per-sample API returns exact-size tensors
        return v.voxels[:m], v.coors[:m, 1:], v.num_points[:m]

    @staticmethod
    def concat(points_list):
        """list[B] of (N_i,C) -> (points (Ntot,C) f32, pt_off (B+1) int32 on device, lens)."""
        lens = [int(p.shape[0]) for p in points_list]
        pts = torch.cat([p.float() for p in points_list], 0).contiguous()
        off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0).tolist()), dtype=torch.int32)
        return pts, off.to(pts.device, non_blocking=True), lens

    def batched(self, points_list, index_dims=None, concat=None):
        """Fused batch path: one launch sequence for the whole batch, VFE mean included.
        index_dims: (D,H,W) index space of the VoxelMap = the encoder's sparse_shape (SECOND-style
        configs declare it one cell deeper in z than the voxel grid); default: the voxel grid."""
        dims = tuple(int(v) for v in index_dims) if index_dims is not None else self.grid_zyx
        if any(g > d for g, d in zip(self.grid_zyx, dims)):
            raise ValueError(f"voxel grid {self.grid_zyx} does not fit sparse_shape {dims}")
        pts, off, lens = concat if concat is not None else self.concat(points_list)
        B = len(lens)
        if self.dynamic:
            v = ops.voxelize_dynamic(pts, off, B, self.point_cloud_range, self.voxel_size, dims)
        else:
            v = ops.voxelize_hard(pts, off, B, self.point_cloud_range, self.voxel_size,
                                  dims, self.max_num_points, self.current_max_voxels(),
                                  deterministic=self.deterministic)
        return pts, off, lens, v


@VOXEL_ENCODERS.register_module()
class HardSimpleVFE(nn.Module):
    """mmdet3d HardSimpleVFE: mean of the points of each voxel."""

    def __init__(self, num_features=4):
        super().__init__()
        self.num_features = num_features
        self.fp16_enabled = False

    def forward(self, features, num_points, coors=None):
        s = features[:, :, :self.num_features].sum(dim=1)
        return (s / num_points.type_as(features).view(-1, 1)).contiguous()


@VOXEL_ENCODERS.register_module()
class DynamicSimpleVFE(nn.Module):
    """mmdet3d DynamicSimpleVFE: DynamicScatter(mean) over (b,z,y,x)."""

    def __init__(self, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1)):
        super().__init__()
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.grid_zyx = grid_size_zyx(point_cloud_range, voxel_size)
        self.fp16_enabled = False

    @torch.no_grad()
    def forward(self, features, coors):
        """features (N,C), coors (N,4) [b,z,y,x] (rows with -1 are dropped) -> (feats, coors)
        in lexicographic (b,z,y,x) order. Re-voxelises the points on the device: the per-point
        coordinates are a pure function of the points, so they are recomputed, not re-read."""
        B = int(coors[-1, 0].item()) + 1
        counts = torch.bincount(coors[:, 0].long(), minlength=B)
        off = torch.zeros(B + 1, dtype=torch.int32, device=features.device)
        off[1:] = counts.cumsum(0).int()
        v = ops.voxelize_dynamic(features.contiguous().float(), off, B, self.point_cloud_range,
                                 self.voxel_size, self.grid_zyx)
        m = int(v.scene_rows[-1].item())
        return v.feats[:m], v.coors[:m]
