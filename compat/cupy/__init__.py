"""The `cupy` names the reference's drivers touch (copy_params.py:7,10,61-63: `xp.all(a == b)` on parameter arrays), over the
torch tensors that `.W.data` views return.  Not CuPy."""
import numpy as _np
import torch as _torch


def asnumpy(a):
    return a.detach().cpu().numpy() if isinstance(a, _torch.Tensor) else _np.asarray(a)


def all(a):  # noqa: A001
    return bool(_torch.all(a)) if isinstance(a, _torch.Tensor) else bool(_np.all(a))


def asarray(a, dtype=None):
    return _np.asarray(a, dtype=dtype)


float32, int32 = _np.float32, _np.int32
