"""`from chainer import serializers` -> save_npz / load_npz with the reference's key set (ast_b200/serializers.py)"""
from ast_b200.serializers import load_npz, save_npz  # noqa: F401
