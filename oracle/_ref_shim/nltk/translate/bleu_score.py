"""nltk.translate.bleu_score stub: the three names eval.py:7-8 imports.  corpus_bleu delegates to the repo's restatement of
the published algorithm (ast_b200/eval.py; host-side text metric, no device code); the unmodified reference eval.py only
works as a smoke path here, it is not a BLEU pin (nltk itself is absent)."""
from ast_b200.eval import corpus_bleu as _corpus_bleu, modified_precision  # noqa: F401


class SmoothingFunction:
    def method2(self, *a, **k):
        raise NotImplementedError("marker object only: corpus_bleu below applies method-2 smoothing itself")


def corpus_bleu(list_of_references, hypotheses, weights=(0.25, 0.25, 0.25, 0.25), smoothing_function=None, **kw):
    return _corpus_bleu(list_of_references, hypotheses, weights=weights)


def sentence_bleu(references, hypothesis, weights=(0.25, 0.25, 0.25, 0.25), smoothing_function=None, **kw):
    return _corpus_bleu([references], [hypothesis], weights=weights)
