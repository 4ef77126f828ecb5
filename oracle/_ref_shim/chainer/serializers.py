"""chainer.serializers stand-in: numpy.savez_compressed, keys = slash-joined link paths,
params then persistents; loading shapes a lazily-unshaped parameter from the stored array."""
import numpy as np

import chainer


def save_npz(file, obj, compression=True):
    out = {}
    for key, kind, link, name in obj._serialize_items():
        v = link.__dict__[name]
        if kind == "param":
            if v is None or v._t is None:
                continue
            out[key] = v.data.copy()
        else:
            out[key] = np.asarray(v)
    with open(file, "wb") as f:
        (np.savez_compressed if compression else np.savez)(f, **out)


def load_npz(file, obj, path="", strict=True):
    with np.load(file) as npz:
        for key, kind, link, name in obj._serialize_items():
            k = path + key
            if k not in npz.files:
                if strict and not (kind == "param" and link.__dict__[name] is None):
                    raise KeyError(k)
                continue
            a = npz[k]
            if kind == "param":
                p = link.__dict__[name]
                if p._t is None:
                    p.initialize(a.shape)
                p.data = a.astype(chainer.config.dtype)
            elif name == "N":
                link.__dict__[name] = int(a)
            else:
                link.__dict__[name] = a.astype(chainer.config.dtype)
