"""`from dataloader import FisherDataLoader, GlobalPhoneDataLoader, SYMBOLS` (nn.py:18, seq2seq.py:20) -> ast_b200.dataloader"""
from ast_b200.dataloader import *  # noqa: F401,F403
from ast_b200.dataloader import DataLoader, FisherDataLoader, GlobalPhoneDataLoader, SyntheticDataLoader  # noqa: F401
from ast_b200.symbols import SYMBOLS  # noqa: F401
