"""`from eval import Eval` (train.py:8, beam.py:8, copy_params.py:2) -> ast_b200.eval"""
from ast_b200.eval import *  # noqa: F401,F403
from ast_b200.eval import Eval, corpus_bleu  # noqa: F401
