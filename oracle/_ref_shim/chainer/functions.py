"""chainer.functions stand-in: exactly the F.* names seq2seq.py / nn.py / dataloader.py call, with the semantics of
SURVEY Appendix A (public Chainer v5).  torch-CPU arithmetic + autograd in chainer.config.dtype."""
import numpy as np
import torch

import chainer
from chainer import Variable, _as_tensor, _tdtype

# ---- dropout mask injection -------------------------------------------------------------------------------
# Chainer's RNG stream cannot be reproduced; golden generation injects the masks.  `dropout_hook(shape, ratio)`
# is called once per F.dropout call with ratio > 0 in train mode, in call order, and returns the SCALED keep
# mask (0 or 1/(1-ratio)) as a numpy array, or None for a fresh random mask.
dropout_hook = None


def _w(t):
    return Variable._wrap(t)


def expand_dims(x, axis):
    return _w(_as_tensor(x).unsqueeze(axis))


def swapaxes(x, a, b):
    return _w(_as_tensor(x).transpose(a, b))


def rollaxis(x, axis, start=0):
    return _w(torch.movedim(_as_tensor(x), axis, start))


def reshape(x, shape):
    return _w(_as_tensor(x).reshape(tuple(shape)))


def squeeze(x, axis=None):
    t = _as_tensor(x)
    return _w(t.squeeze() if axis is None else t.squeeze(axis))


def flipud(x):
    return _w(torch.flip(_as_tensor(x), dims=(0,)))


def concat(xs, axis=1):
    return _w(torch.cat([_as_tensor(x) for x in xs], dim=axis))


def relu(x):
    return _w(torch.relu(_as_tensor(x)))


def tanh(x):
    return _w(torch.tanh(_as_tensor(x)))


def sigmoid(x):
    return _w(torch.sigmoid(_as_tensor(x)))


def dropout(x, ratio=0.5):
    """A.5: train => x * mask / (1 - ratio), mask ~ Bernoulli(1 - ratio); test => identity."""
    t = _as_tensor(x)
    if not chainer.config.train:
        return _w(t)
    if ratio <= 0:
        return _w(t * 1.0)
    m = None
    if dropout_hook is not None:
        m = dropout_hook(tuple(t.shape), ratio)
    if m is None:
        keep = np.random.rand(*t.shape) >= ratio
        m = keep.astype(np.float64) / (1.0 - ratio)
    return _w(t * _as_tensor(np.asarray(m)))


def softmax(x, axis=1):
    return _w(torch.softmax(_as_tensor(x), dim=axis))


def log_softmax(x, axis=1):
    return _w(torch.log_softmax(_as_tensor(x), dim=axis))


def batch_matmul(a, b, transa=False, transb=False):
    """A.6: operands are batches of matrices; a 2-D operand (B, K) is a batch of (K, 1) column vectors."""
    ta, tb = _as_tensor(a), _as_tensor(b)
    if ta.dim() == 2:
        ta = ta.unsqueeze(2)
    if tb.dim() == 2:
        tb = tb.unsqueeze(2)
    if transa:
        ta = ta.transpose(1, 2)
    if transb:
        tb = tb.transpose(1, 2)
    return _w(torch.bmm(ta, tb))


def argmax(x, axis=None):
    """Ties -> lowest index (numpy semantics)."""
    a = _as_tensor(x).detach().numpy()
    return Variable(np.argmax(a, axis=axis).astype(np.int32))


def softmax_cross_entropy(x, t, class_weight=None, normalize=True, ignore_label=-1, reduce="mean"):
    """A.7: logp = log_softmax(x) * w[None, :]; loss = -sum_n logp[n, t_n] / max(#{t_n != ignore_label}, 1)."""
    assert reduce == "mean" and normalize
    z = _as_tensor(x)
    tt = torch.from_numpy(np.asarray(t.data if isinstance(t, Variable) else t).astype(np.int64))
    logp = torch.log_softmax(z, dim=1)
    if class_weight is not None:
        logp = logp * _as_tensor(np.asarray(class_weight)).to(z.dtype).unsqueeze(0)
    valid = tt != ignore_label
    count = max(int(valid.sum()), 1)
    picked = logp.gather(1, tt.clamp(min=0).unsqueeze(1)).squeeze(1)
    picked = torch.where(valid, picked, torch.zeros_like(picked))
    return _w(-picked.sum() / count)


def pad_sequence(xs, length=None, padding=0):
    """A.8: pad every array to the longest (axis 0) -> (B, Lmax, ...) Variable (dtype of the inputs)."""
    arrs = [np.asarray(x.data if isinstance(x, Variable) else x) for x in xs]
    L = max(a.shape[0] for a in arrs) if length is None else length
    out = np.full((len(arrs), L) + arrs[0].shape[1:], padding, dtype=arrs[0].dtype)
    for i, a in enumerate(arrs):
        out[i, :a.shape[0]] = a
    return Variable(out)


def lstm(c_prev, x):
    """A.3: gates interleaved, index 4*j + k with k in (a, i, f, o) for unit j."""
    c_prev, x = _as_tensor(c_prev), _as_tensor(x)
    B, n4 = x.shape
    g = x.reshape(B, n4 // 4, 4)
    a, i, f, o = torch.tanh(g[:, :, 0]), torch.sigmoid(g[:, :, 1]), torch.sigmoid(g[:, :, 2]), torch.sigmoid(g[:, :, 3])
    c = a * i + f * c_prev
    h = o * torch.tanh(c)
    return _w(c), _w(h)
