"""chainer.optimizer stand-in: the hooks nn.py:100-110 attaches and the GradientMethod
update order: hooks in insertion order over ALL parameters (frozen ones included), then t += 1, then each enabled
parameter's rule."""
import numpy as np
import torch


class WeightDecay:
    name = "WeightDecay"

    def __init__(self, rate):
        self.rate = rate

    def __call__(self, opt):
        for p in opt.target.params():
            if p._t is not None and p._t.grad is not None:
                p._t.grad.add_(p._t.detach(), alpha=self.rate)


class GradientClipping:
    name = "GradientClipping"

    def __init__(self, threshold):
        self.threshold = threshold

    def __call__(self, opt):
        sq = 0.0
        for p in opt.target.params():
            if p._t is not None and p._t.grad is not None:
                sq += float((p._t.grad.double() ** 2).sum())
        norm = np.sqrt(sq)
        opt.last_grad_norm = norm
        rate = self.threshold / norm if norm > 0 else np.inf
        if rate < 1:
            for p in opt.target.params():
                if p._t is not None and p._t.grad is not None:
                    p._t.grad.mul_(rate)


class GradientNoise:
    name = "GradientNoise"

    def __init__(self, eta, noise_func=None):
        self.eta = eta

    def __call__(self, opt):
        std = np.sqrt(self.eta / np.power(1 + opt.t, 0.55))
        for p in opt.target.params():
            if p._t is not None and p._t.grad is not None:
                p._t.grad.add_(torch.from_numpy(np.random.normal(0, std, tuple(p._t.shape))).to(p._t.dtype))


class GradientMethod:
    def __init__(self):
        self.t = 0
        self.target = None
        self._hooks = []
        self.last_grad_norm = None

    def setup(self, link):
        self.target = link
        self.t = 0
        return self

    def add_hook(self, hook, name=None):
        self._hooks.append(hook)

    def update(self, lossfun=None, *args, **kwds):
        assert lossfun is None
        for p in self.target.params():               # reallocate_cleared_grads
            if p._t is not None and p._t.grad is None:
                p._t.grad = torch.zeros_like(p._t)
        for h in self._hooks:
            h(self)
        self.t += 1
        for p in self.target.params():
            if p._t is not None and p.update_rule.enabled:
                self.update_one(p)
