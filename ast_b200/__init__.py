"""ast_b200 — B200-native (sm_100a) hot path of 0xSameer/ast behind the reference's Python API."""
import os as _os

# The library drives ~10 CUDA streams per model (layer wavefront, GEMM streams, side stream, packer, NCCL); with the default of
# 8 hardware queues they alias and pick up false dependencies.  Read by the driver when the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .symbols import SYMBOLS  # noqa: E402,F401

__all__ = ["SYMBOLS"]
