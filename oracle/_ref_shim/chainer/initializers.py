"""chainer.initializers used by the reference.  The golden generator overwrites every
parameter with oracle.init_params afterwards, so the random stream only has to be well-formed."""
import numpy as np

_rng = np.random.RandomState(12345)


def _fans(shape):
    fan_out = shape[0]
    fan_in = int(np.prod(shape[1:]))
    return fan_in, fan_out


class HeNormal:
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, a):
        fan_in, _ = _fans(a.shape)
        a[...] = _rng.normal(0.0, self.scale * np.sqrt(2.0 / fan_in), a.shape)


class LeCunNormal:
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, a):
        fan_in, _ = _fans(a.shape)
        a[...] = _rng.normal(0.0, self.scale * np.sqrt(1.0 / fan_in), a.shape)


class Normal:
    def __init__(self, scale=0.05):
        self.scale = scale

    def __call__(self, a):
        a[...] = _rng.normal(0.0, self.scale, a.shape)


class Constant:
    def __init__(self, v):
        self.v = v

    def __call__(self, a):
        a[...] = self.v


Zero = lambda: Constant(0.0)   # noqa: E731
One = lambda: Constant(1.0)    # noqa: E731
