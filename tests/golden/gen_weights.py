"""Seeded weight generator shared by make_golden.py and the tests: the golden fixtures store
inputs/outputs only; weights are regenerated from (name, shape, seed)."""
import zlib

import torch


def gen_state_dict(shapes, seed):
    """shapes: {name: shape}. Values depend only on (seed, name, shape)."""
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros(shape, dtype=torch.long)
        elif name.endswith("running_var"):
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif len(shape) <= 1:
            base = 1.0 if (name.endswith("weight") and "embed" not in name) else 0.0
            sd[name] = base + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if "embed" in name:
                fan_in = 1
            sd[name] = torch.randn(shape, generator=g) / max(fan_in, 1) ** 0.5
    return sd
