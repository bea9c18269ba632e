"""Deterministic synthetic scenes and random-init models.

There is no network for datasets or checkpoints: the bench and the parity tests use these
generators (seed = 1000*config_id + scene_idx) and ``torch.manual_seed(0)`` default inits with
randomised BatchNorm running statistics so eval-mode BN is a non-trivial affine.
"""
import os

import numpy as np
import torch

CONFIG_FILES = {
    "sunrgbd": "uni3detr_sunrgbd.py",
    "scannet_large": "uni3detr_scannet_large.py",
    "kitti": "uni3detr_kitti_3classes.py",
    "nuscenes": "uni3detr_nuscenes.py",
}

# model-relevant facts of the four BASELINE configs, so nothing has to read the reference tree at
# run time on the GPU box (tests/golden/configs.json holds the full model dicts)
WORKLOADS = {
    "sunrgbd": dict(config_id=2, n_points=20000, C=4, generator="room",
                    pc_range=[-3.2, -0.2, -2., 3.2, 6.2, 0.56]),
    "scannet_large": dict(config_id=3, n_points=100000, C=4, generator="room",
                          pc_range=[-6.4, -6.4, -0.1, 6.4, 6.4, 2.46]),
    "kitti": dict(config_id=4, n_points=20000, C=4, generator="lidar",
                  pc_range=[0, -40, -3, 70.4, 40, 1], fov_deg=45.0),
    "nuscenes": dict(config_id=5, n_points=200000, C=5, generator="lidar",
                     pc_range=[-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], fov_deg=180.0),
}


def room_scene(n, pc_range, seed, C=4):
    """Indoor 'room': 40% floor, 20%+20% two walls, 20% on the faces of 10 random boxes;
    5 mm Gaussian noise on the constrained axis; 4th feature = z - percentile(z, 0.99)."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(pc_range[:3], np.float64), np.asarray(pc_range[3:], np.float64)
    ext = hi - lo
    counts = [int(0.4 * n), int(0.2 * n), int(0.2 * n)]
    counts.append(n - sum(counts))
    parts = []
    p = lo + rng.random((counts[0], 3)) * ext
    p[:, 2] = lo[2] + 0.05 * ext[2] + rng.normal(0, 0.005, counts[0])
    parts.append(p)
    p = lo + rng.random((counts[1], 3)) * ext
    p[:, 1] = hi[1] - 0.05 * ext[1] + rng.normal(0, 0.005, counts[1])
    parts.append(p)
    p = lo + rng.random((counts[2], 3)) * ext
    p[:, 0] = lo[0] + 0.05 * ext[0] + rng.normal(0, 0.005, counts[2])
    parts.append(p)
    nb = 10
    centres = lo + (0.2 + 0.6 * rng.random((nb, 3))) * ext
    sizes = 0.3 + 0.7 * rng.random((nb, 3))
    which = rng.integers(0, nb, counts[3])
    face = rng.integers(0, 6, counts[3])
    u = rng.random((counts[3], 3)) - 0.5
    p = centres[which] + u * sizes[which]
    ax = face % 3
    sign = np.where(face < 3, -0.5, 0.5)
    rows = np.arange(counts[3])
    p[rows, ax] = centres[which, ax] + sign * sizes[which, ax] + rng.normal(0, 0.005, counts[3])
    parts.append(p)
    pts = np.concatenate(parts)
    pts = np.clip(pts, lo + 1e-4, hi - 1e-4)
    rng.shuffle(pts)
    out = np.zeros((n, C), np.float32)
    out[:, :3] = pts
    if C > 3:
        out[:, 3] = pts[:, 2] - np.percentile(pts[:, 2], 0.99)
    if C > 4:
        out[:, 4:] = rng.random((n, C - 4))
    return out


def lidar_scene(n, pc_range, seed, C=4, fov_deg=180.0):
    """Outdoor 'lidar': 64 beams (-24.8..+2 deg), sensor at 1.7 m, 65% ground / 35% obstacles."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(pc_range[:3], np.float64), np.asarray(pc_range[3:], np.float64)
    pts = np.zeros((0, 3))
    while len(pts) < n:
        m = 2 * n
        el = np.deg2rad(rng.choice(np.linspace(-24.8, 2.0, 64), m))
        az = np.deg2rad(rng.uniform(-fov_deg, fov_deg, m))
        ground = rng.random(m) < 0.65
        with np.errstate(divide="ignore", invalid="ignore"):
            r_ground = np.where(el < -0.01, 1.7 / np.tan(-el), 1e9)
        r_obs = rng.gamma(2.0, 12.0, m)
        r = np.where(ground & (r_ground < 120), r_ground, r_obs) + rng.normal(0, 0.02, m)
        x, y = r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az)
        z = r * np.sin(el)  # sensor frame: ground at -1.7 m
        q = np.stack([x, y, z], 1)
        ok = np.all((q > lo + 1e-3) & (q < hi - 1e-3), axis=1)
        pts = np.concatenate([pts, q[ok]])
    pts = pts[:n]
    out = np.zeros((n, C), np.float32)
    out[:, :3] = pts
    if C > 3:
        out[:, 3] = rng.random(n)
    if C > 4:
        out[:, 4] = rng.integers(0, 10, n) * 0.05
    return out


def make_scene(workload, scene_idx, n_points=None):
    w = WORKLOADS[workload]
    n = n_points or w["n_points"]
    seed = 1000 * w["config_id"] + scene_idx
    if w["generator"] == "room":
        return room_scene(n, w["pc_range"], seed, w["C"])
    return lidar_scene(n, w["pc_range"], seed, w["C"], w["fov_deg"])


def randomize_bn_(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def randomize_gates_(model, seed=1):
    """attention_weights are zero-initialised in the reference (gate == 0.5); give them small
    random values so the gate kernel path is actually exercised by parity tests."""
    g = torch.Generator().manual_seed(seed)
    from .plugin.transformer import UniCrossAtten
    for m in model.modules():
        if isinstance(m, UniCrossAtten):
            m.attention_weights.weight.data.copy_(0.05 * torch.randn(m.attention_weights.weight.shape, generator=g))
            m.attention_weights.bias.data.copy_(0.1 * torch.randn(m.attention_weights.bias.shape, generator=g))


def load_model_cfg(workload):
    """Model dict of a BASELINE config: from the committed JSON copy of the reference's
    config `model=` sections (tests/golden/configs.json) so it works on the GPU box."""
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "..", "tests", "golden", "configs.json")
    with open(path) as f:
        return json.load(f)[workload]


def build_model(workload, seed=0, overrides=None, randomize=True):
    from . import compat, register_all
    register_all()
    cfg = load_model_cfg(workload)
    if overrides:
        cfg = compat._merge(cfg, overrides)
    torch.manual_seed(seed)
    model = compat.build_model(cfg)
    model.pts_bbox_head.init_weights()
    if randomize:
        randomize_bn_(model, seed)
        randomize_gates_(model, seed + 1)
    model.eval()
    return model, cfg
