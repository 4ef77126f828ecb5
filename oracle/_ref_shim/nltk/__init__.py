"""TEST INFRASTRUCTURE ONLY - `nltk` stub so the reference's eval.py (imported by dataloader.py:12) resolves."""
