"""`from seq2seq import SpeechEncoderDecoder` (nn.py:16) -> ast_b200.seq2seq"""
from ast_b200.seq2seq import *  # noqa: F401,F403
from ast_b200.seq2seq import SpeechEncoderDecoder, Variable, SYMBOLS  # noqa: F401
