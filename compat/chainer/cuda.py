"""`from chainer import cuda`; `xp = cuda.cupy` (copy_params.py:8-10)"""
import cupy  # noqa: F401  (compat/cupy)


class _Device:
    def __init__(self, i=0):
        self.id = i

    def use(self):
        import torch
        torch.cuda.set_device(self.id)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def get_device(i=0):
    return _Device(i)


get_device_from_id = get_device
