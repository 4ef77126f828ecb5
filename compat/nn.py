"""`from nn import NN` (train.py:7, beam.py:7, copy_params.py:1) -> ast_b200.nn"""
from ast_b200.nn import *  # noqa: F401,F403
from ast_b200.nn import NN, Adam, SGD, WeightDecay, GradientClipping, GradientNoise, using_config, _ADAM, _SGD  # noqa: F401
