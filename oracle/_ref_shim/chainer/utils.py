"""chainer.utils: imported by name only (seq2seq.py:15)."""
