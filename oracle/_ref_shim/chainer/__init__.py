"""TEST INFRASTRUCTURE ONLY - a stand-in for the `chainer` names the reference's live path imports.

Purpose: Chainer/CuPy are neither vendored nor installable here, so the
reference's own source could never run.  With this package (and the sibling `cupy`, `nltk` stubs) first on
sys.path, `/root/reference/{seq2seq,nn,dataloader,config,eval}.py` import and execute UNMODIFIED on the CPU.
That pins the *control flow* of the reference (the `X[-i]` order, per-layer dropout placement, (c, h) argument
order, the scheduled-sampling draw order, PAD-weighted CE, hook order, the beam bookkeeping and its stable sort) to
its own source; what stays unpinned is the op semantics restated below from SURVEY Appendix A (public Chainer v5
behaviour), because the real library is absent.

Arithmetic is torch-CPU (conv2d, matmul, autograd) in `config.dtype` (float64 for golden generation): an
implementation independent of oracle/ast_oracle.py's numpy im2col + hand-written backward.

Nothing outside tests/, oracle/ and tools/ may import this.
"""
import contextlib

import numpy as np
import torch

torch.set_num_threads(max(1, torch.get_num_threads()))


class _Config:
    train = True
    dtype = np.float64          # compute dtype of every floating Variable / parameter
    enable_backprop = True


config = _Config()
configuration = type("configuration", (), {"config": config})


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)


def no_backprop_mode():
    return using_config("enable_backprop", False)


def _tdtype():
    return torch.float64 if np.dtype(config.dtype) == np.float64 else torch.float32


def _as_tensor(x):
    """numpy / scalar / tensor -> torch tensor; floating data is cast to the compute dtype."""
    if isinstance(x, Variable):
        return x._t
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype == np.bool_:
            a = a.copy()
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.is_floating_point() and t.dtype != _tdtype():
        t = t.to(_tdtype())
    return t


class Variable:
    """chainer.Variable: `.data` / `.array` are numpy views of the value, autograd is torch's."""

    def __init__(self, data=None, name=None, grad=None, requires_grad=False):
        self.name = name
        self._t = None
        if data is not None:
            t = _as_tensor(data)
            if requires_grad:
                t = t.detach().clone().requires_grad_(True)
            self._t = t

    @classmethod
    def _wrap(cls, t):
        v = cls.__new__(cls)
        v.name = None
        v._t = t
        return v

    # -- data access --------------------------------------------------------------------------------
    @property
    def data(self):
        return None if self._t is None else self._t.detach().numpy()

    @data.setter
    def data(self, value):
        t = _as_tensor(value)
        if self._t is not None and self._t.requires_grad and self._t.is_leaf:
            with torch.no_grad():
                if tuple(self._t.shape) == tuple(t.shape):
                    self._t.copy_(t)
                    return
            self._t = t.detach().clone().requires_grad_(True)
        else:
            self._t = t

    array = data

    @property
    def grad(self):
        return None if self._t is None or self._t.grad is None else self._t.grad.numpy()

    @grad.setter
    def grad(self, g):
        self._t.grad = None if g is None else _as_tensor(g).clone()

    def cleargrad(self):
        if self._t is not None:
            self._t.grad = None

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def size(self):
        return self._t.numel()

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def T(self):
        return Variable._wrap(self._t.t())

    def __len__(self):
        return self._t.shape[0]

    def __getitem__(self, idx):
        return Variable._wrap(self._t[idx])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def to_gpu(self, device=None):
        return self

    def to_cpu(self):
        return self

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Variable._wrap(self._t.reshape(*shape))

    def backward(self, retain_grad=False):
        self._t.backward()

    def unchain_backward(self):
        self._t = self._t.detach()

    # -- arithmetic -----------------------------------------------------------------------------------
    def _bin(self, other, fn):
        return Variable._wrap(fn(self._t, _as_tensor(other)))

    def __add__(self, o):
        return self._bin(o, torch.add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, torch.sub)

    def __rsub__(self, o):
        return Variable._wrap(_as_tensor(o) - self._t)

    def __mul__(self, o):
        return self._bin(o, torch.mul)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, torch.div)

    def __neg__(self):
        return Variable._wrap(-self._t)

    def __float__(self):
        return float(self._t)

    def __repr__(self):
        return f"variable({self.data!r})"


class Parameter(Variable):
    """chainer.Parameter: a leaf Variable, possibly uninitialised (lazy in_size / in_channels)."""

    def __init__(self, initializer=None, shape=None, name=None):
        super().__init__(None, name=name)
        self.initializer = initializer
        self.update_rule = _UpdateRule()
        if shape is not None:
            self.initialize(shape)

    def initialize(self, shape):
        a = np.zeros(shape, dtype=config.dtype)
        init = self.initializer
        if callable(init):
            init(a)
        elif init is not None:
            a[...] = init
        self._t = torch.from_numpy(a).to(_tdtype()).requires_grad_(True)


class _UpdateRule:
    def __init__(self):
        self.enabled = True
        self.state = None
        self.t = 0


class Link:
    """chainer.Link: parameters and persistents registered by name; children in Chain."""

    def __init__(self):
        self.__dict__["_params"] = []
        self.__dict__["_persistent"] = []
        self.__dict__["name"] = None

    def add_param(self, name, shape=None, initializer=None):
        p = Parameter(initializer, shape, name=name)
        self.__dict__[name] = p
        self._params.append(name)
        return p

    def add_persistent(self, name, value):
        self.__dict__[name] = value
        self._persistent.append(name)

    def params(self, include_uninit=True):
        for n in sorted(self._params):
            p = self.__dict__[n]
            if include_uninit or p._t is not None:
                yield p

    def namedparams(self, include_uninit=True):
        for n in sorted(self._params):
            p = self.__dict__[n]
            if include_uninit or p._t is not None:
                yield "/" + n, p

    def links(self, skipself=False):
        if not skipself:
            yield self

    def namedlinks(self, skipself=False):
        if not skipself:
            yield "/", self

    def to_gpu(self, device=None):
        return self

    def to_cpu(self):
        return self

    def cleargrads(self):
        for p in self.params():
            p.cleargrad()

    def zerograds(self):
        for p in self.params():
            if p._t is not None:
                p._t.grad = torch.zeros_like(p._t)

    def disable_update(self):
        for p in self.params():
            p.update_rule.enabled = False

    def enable_update(self):
        for p in self.params():
            p.update_rule.enabled = True

    def _serialize_items(self, prefix=""):
        """(key, kind, getter, setter) for serializers: params then persistents."""
        for n in sorted(self._params):
            yield prefix + n, "param", self, n
        for n in sorted(self._persistent):
            yield prefix + n, "persistent", self, n


class Chain(Link):
    def __init__(self, **links):
        super().__init__()
        self.__dict__["_children"] = []
        for name, link in links.items():
            self.add_link(name, link)

    def add_link(self, name, link):
        link.__dict__["name"] = name
        self.__dict__[name] = link
        self._children.append(name)

    def __setattr__(self, name, value):
        # re-binding a child (copy_params.py:26-43) keeps it a registered child
        self.__dict__[name] = value

    def __getitem__(self, name):
        return self.__dict__[name]

    def children(self):
        for n in self._children:
            yield self.__dict__[n]

    def params(self, include_uninit=True):
        yield from super().params(include_uninit)
        for n in sorted(self._children):
            yield from self.__dict__[n].params(include_uninit)

    def namedparams(self, include_uninit=True):
        yield from super().namedparams(include_uninit)
        for n in sorted(self._children):
            for path, p in self.__dict__[n].namedparams(include_uninit):
                yield "/" + n + path, p

    def links(self, skipself=False):
        if not skipself:
            yield self
        for n in sorted(self._children):
            yield from self.__dict__[n].links()

    def cleargrads(self):
        for p in self.params():
            p.cleargrad()

    def disable_update(self):
        for p in self.params():
            p.update_rule.enabled = False

    def enable_update(self):
        for p in self.params():
            p.update_rule.enabled = True

    def _serialize_items(self, prefix=""):
        yield from super()._serialize_items(prefix)
        for n in sorted(self._children):
            yield from self.__dict__[n]._serialize_items(prefix + n + "/")


class Function:                 # imported by name only (seq2seq.py:15, nn.py:22)
    pass


from . import cuda, utils, initializers, functions, links, optimizer, optimizers, serializers  # noqa: E402,F401
from . import backends  # noqa: E402,F401
