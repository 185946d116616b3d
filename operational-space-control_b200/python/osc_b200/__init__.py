"""osc_b200: host-side Python mirror of the batched operational-space controller.

The product path is the CUDA library `libosc_b200.so` (C-ABI in `include/osc_b200.h`);
importing this package does not load it -- `osc_b200.capi` does, and fails
loudly when the library or a GPU is missing (there is no CPU fallback)."""
from .specs import RobotSpec, load_preset, preset_names, from_yaml  # noqa: F401
from . import synth  # noqa: F401
