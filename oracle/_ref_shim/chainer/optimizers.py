"""chainer.optimizers stand-in: Adam(amsgrad=True) and SGD with the arithmetic of SURVEY Appendix A.10."""
import math

import torch

from .optimizer import GradientMethod


class Adam(GradientMethod):
    def __init__(self, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-08, eta=1.0, weight_decay_rate=0, amsgrad=False):
        super().__init__()
        self.alpha, self.beta1, self.beta2, self.eps, self.eta = alpha, beta1, beta2, eps, eta
        self.weight_decay_rate, self.amsgrad = weight_decay_rate, amsgrad

    @property
    def lr(self):
        fix1 = 1.0 - math.pow(self.beta1, self.t)
        fix2 = 1.0 - math.pow(self.beta2, self.t)
        return self.alpha * math.sqrt(fix2) / fix1

    def update_one(self, p):
        r = p.update_rule
        g = p._t.grad
        if r.state is None:
            r.state = {"m": torch.zeros_like(p._t), "v": torch.zeros_like(p._t)}
            if self.amsgrad:
                r.state["vhat"] = torch.zeros_like(p._t)
        m, v = r.state["m"], r.state["v"]
        m += (1 - self.beta1) * (g - m)
        v += (1 - self.beta2) * (g * g - v)
        if self.amsgrad:
            vhat = r.state["vhat"]
            torch.maximum(vhat, v, out=vhat)
        else:
            vhat = v
        with torch.no_grad():
            p._t -= self.eta * (self.lr * m / (torch.sqrt(vhat) + self.eps) + self.weight_decay_rate * p._t)


class SGD(GradientMethod):
    def __init__(self, lr=0.01):
        super().__init__()
        self.lr = lr

    def update_one(self, p):
        with torch.no_grad():
            p._t -= self.lr * p._t.grad
