"""`from config import Config` (nn.py:17) -> ast_b200.config"""
from ast_b200.config import Config  # noqa: F401
