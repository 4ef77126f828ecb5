"""The `chainer` names the reference's DRIVERS touch (train.py:11,75; beam.py:11; copy_params.py:6-8), answered by ast_b200.
Not Chainer: no links / functions / autograd live here - the model behind `nn.NN` is the CUDA engine."""
from ast_b200 import serializers  # noqa: F401
from ast_b200.nn import using_config  # noqa: F401
from ast_b200.seq2seq import Variable, config  # noqa: F401
from . import cuda  # noqa: F401


class Function:          # imported by name only (copy_params.py:8)
    pass


class utils:             # imported by name only
    pass
