"""uni3detr_b200 — Blackwell-native (sm_100a) implementation of the Uni3DETR per-scene forward
hot path, exposed under the reference's plugin registry names.

Layout: ``csrc/`` hand-written CUDA behind the C ABI of ``include/u3d.h`` (built into
``libu3d_b200.so``), ``ops.py`` tensor-level bindings, ``plugin/`` the drop-in modules
(`Uni3DETR`, `SparseEncoderHD`, `SECOND3D`, `SECOND3DFPN`, `Uni3DETRHead`,
`Uni3DETRTransformer`, `Uni3DETRTransformerDecoder`, `UniCrossAtten`, `NMSFreeCoder`),
``compat.py`` the registry/config shim, ``synth.py`` the synthetic scene generators.
"""
from . import compat  # noqa: F401
from .compat import Config, build_model  # noqa: F401

__version__ = "0.1.0"


def GraphedForward(*args, **kwargs):
    """CUDA-graph replay of the forward for fixed batch signatures (see graphed.py)."""
    from .graphed import GraphedForward as _G
    return _G(*args, **kwargs)


def register_all():
    """Import the plugin modules (registers every drop-in class)."""
    from . import plugin  # noqa: F401
    return plugin
