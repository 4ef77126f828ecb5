"""chainer.links stand-in: Convolution2D, BatchNormalization, LSTM, Linear, EmbedID, LayerNormalization with the
parameter names, shapes, lazy shaping and defaults of public Chainer v5."""
import numpy as np
import torch

import chainer
from chainer import Chain, Link, Variable, _as_tensor, initializers
from chainer import functions as F


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class Linear(Link):
    """A.4: y = x W^T + b, W (out, in) LeCunNormal, b = 0; in_size None => shaped at the first call."""

    def __init__(self, in_size, out_size=None, nobias=False, initialW=None, initial_bias=None):
        super().__init__()
        if out_size is None:
            in_size, out_size = None, in_size
        self.out_size = out_size
        self.add_param("W", None, initialW if initialW is not None else initializers.LeCunNormal())
        if in_size is not None:
            self.W.initialize((out_size, in_size))
        if nobias:
            self.__dict__["b"] = None
        else:
            self.add_param("b", (out_size,), initial_bias if initial_bias is not None else 0.0)

    def __call__(self, x):
        t = _as_tensor(x)
        if t.dim() > 2:
            t = t.reshape(t.shape[0], -1)
        if self.W._t is None:
            self.W.initialize((self.out_size, t.shape[1]))
        y = t @ self.W._t.t()
        if self.b is not None:
            y = y + self.b._t
        return Variable._wrap(y)


class EmbedID(Link):
    """A.4: W (V, E) ~ N(0, 1); backward scatter-adds."""

    def __init__(self, in_size, out_size, initialW=None, ignore_label=None):
        super().__init__()
        self.add_param("W", (in_size, out_size), initialW if initialW is not None else initializers.Normal(1.0))

    def __call__(self, x):
        idx = torch.from_numpy(np.asarray(x.data if isinstance(x, Variable) else x).astype(np.int64))
        return Variable._wrap(self.W._t[idx])


class Convolution2D(Link):
    """A.1: cross-correlation, W (out, in, kh, kw), out = floor((n + 2p - d(k-1) - 1) / s) + 1; in_channels None => lazy."""

    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, nobias=False, initialW=None,
                 initial_bias=None, dilate=1, groups=1):
        super().__init__()
        self.out_channels, self.ksize = out_channels, _pair(ksize)
        self.stride, self.pad, self.dilate = _pair(stride), _pair(pad), _pair(dilate)
        self.add_param("W", None, initialW if initialW is not None else initializers.LeCunNormal())
        if in_channels is not None:
            self.W.initialize((out_channels, in_channels) + self.ksize)
        if nobias:
            self.__dict__["b"] = None
        else:
            self.add_param("b", (out_channels,), initial_bias if initial_bias is not None else 0.0)

    def __call__(self, x):
        t = _as_tensor(x)
        if self.W._t is None:
            self.W.initialize((self.out_channels, t.shape[1]) + self.ksize)
        y = torch.nn.functional.conv2d(t, self.W._t, None if self.b is None else self.b._t, stride=self.stride,
                                       padding=self.pad, dilation=self.dilate)
        return Variable._wrap(y)


class BatchNormalization(Link):
    """A.2: gamma = 1, beta = 0, avg_mean = 0, avg_var = 1, decay 0.9, eps 2e-5, statistics over every axis but 1.
    Train: biased batch variance normalises; running variance accumulates the unbiased one (no eps, cuDNN path)."""

    def __init__(self, size, decay=0.9, eps=2e-5, dtype=None):
        super().__init__()
        self.decay, self.eps = decay, eps
        self.add_param("gamma", (size,), 1.0)
        self.add_param("beta", (size,), 0.0)
        self.add_persistent("avg_mean", np.zeros(size, dtype=chainer.config.dtype))
        self.add_persistent("avg_var", np.ones(size, dtype=chainer.config.dtype))
        self.add_persistent("N", 0)

    def __call__(self, x, finetune=False):
        t = _as_tensor(x)
        axes = tuple(i for i in range(t.dim()) if i != 1)
        shp = [1] * t.dim()
        shp[1] = -1
        g, b = self.gamma._t.reshape(shp), self.beta._t.reshape(shp)
        if chainer.config.train:
            mean = t.mean(dim=axes, keepdim=True)
            var = ((t - mean) ** 2).mean(dim=axes, keepdim=True)
            y = g * (t - mean) / torch.sqrt(var + self.eps) + b
            m = t.numel() // t.shape[1]
            adjust = m / max(m - 1.0, 1.0)
            mean_np = mean.detach().reshape(-1).numpy()
            var_np = var.detach().reshape(-1).numpy()
            self.__dict__["avg_mean"] = self.decay * self.avg_mean + (1 - self.decay) * mean_np
            self.__dict__["avg_var"] = self.decay * self.avg_var + (1 - self.decay) * adjust * var_np
            self.__dict__["N"] = self.N + 1
        else:
            mean = _as_tensor(self.avg_mean).reshape(shp)
            var = _as_tensor(self.avg_var).reshape(shp)
            y = g * (t - mean) / torch.sqrt(var + self.eps) + b
        return Variable._wrap(y)


class LayerNormalization(Link):
    def __init__(self, size=None, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.add_param("gamma", (size,), 1.0)
        self.add_param("beta", (size,), 0.0)

    def __call__(self, x):
        t = _as_tensor(x)
        mu = t.mean(dim=1, keepdim=True)
        var = ((t - mu) ** 2).mean(dim=1, keepdim=True)
        return Variable._wrap(self.gamma._t * (t - mu) / torch.sqrt(var + self.eps) + self.beta._t)


def _lstm_bias_init(a):
    """A.3: bias 0 except the forget gate (interleaved index 4j + 2) = 1."""
    a[...] = 0.0
    a.reshape(-1, 4)[:, 2] = 1.0


class LSTM(Chain):
    """A.3: upward = Linear(in, 4*out) with bias, lateral = Linear(out, 4*out, nobias); gates = upward(x)
    (+ lateral(h) if h is not None); c starts at 0; set_state(c, h); h / c are the pre-dropout link states."""

    def __init__(self, in_size, out_size=None, lateral_init=None, upward_init=None, bias_init=None, forget_bias_init=None):
        if out_size is None:
            in_size, out_size = None, in_size
        super().__init__()
        self.add_link("upward", Linear(in_size, 4 * out_size, initial_bias=_lstm_bias_init))
        self.add_link("lateral", Linear(out_size, 4 * out_size, nobias=True))
        self.__dict__["state_size"] = out_size
        self.reset_state()

    def reset_state(self):
        self.__dict__["c"] = None
        self.__dict__["h"] = None

    def set_state(self, c, h):
        assert isinstance(c, Variable) and isinstance(h, Variable)
        self.__dict__["c"] = c
        self.__dict__["h"] = h

    def __call__(self, x):
        gates = self.upward(x)
        if self.h is not None:
            gates = gates + self.lateral(self.h)
        if self.c is None:
            self.__dict__["c"] = Variable(np.zeros((x.shape[0], self.state_size), dtype=chainer.config.dtype))
        c, h = F.lstm(self.c, gates)
        self.__dict__["c"], self.__dict__["h"] = c, h
        return h
