"""ast_b200 — B200-native (sm_100a) hot path of 0xSameer/ast behind the reference's Python API."""
from .symbols import SYMBOLS  # noqa: F401

__all__ = ["SYMBOLS"]
