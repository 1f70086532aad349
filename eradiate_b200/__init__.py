"""
eradiate_b200 -- B200-native Monte Carlo volumetric path tracer behind
``eradiate.kernel.mi_load_dict / mi_traverse / mi_render``.

Only the hot path lives here: the host-side mirror of the reference's kernel
boundary (``eradiate_b200.kernel``), the CUDA kernels + C ABI (``csrc/``) and the
sample-sharding helper for multi-GPU runs (``eradiate_b200.dist``).
"""

__version__ = "0.1.0"

from . import kernel  # noqa: F401
