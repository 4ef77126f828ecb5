"""TEST INFRASTRUCTURE ONLY - `cupy` stand-in: numpy plus the handful of cupy-only names the reference touches
(`cupy.cuda.Device(gpuid)` context manager seq2seq.py:152,487; `xp.asnumpy` nn.py:269)."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import random, newaxis, bool_, float32, int32, float64, int64  # noqa: F401


def asnumpy(a):
    return _np.asarray(a)


class _Device:
    def __init__(self, i=0):
        self.id = i

    def use(self):
        return None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class cuda:
    Device = _Device
