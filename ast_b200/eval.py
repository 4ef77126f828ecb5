"""Evaluation: multi-reference corpus BLEU and hypothesis dump, drop-in for the reference's `eval.py`
(`Eval(path, n_evals).calc_bleu(hyps)`, `write_to_file`; callers `train.py:44,64`, `beam.py:87,138`).

The reference delegates to `nltk.translate.bleu_score.corpus_bleu(refs, hyps, weights=(.25,)*4,
smoothing_function=SmoothingFunction().method2)` (`eval.py:29-38`); nltk is not vendored and not installable here,
so the published algorithm is restated (Papineni et al. 2002 corpus BLEU; Lin & Och 2004 "add one" smoothing):
  * per order n: clipped n-gram matches (clip = max count over the references) summed over the corpus, divided by
    the summed hypothesis n-gram counts (a hypothesis shorter than n contributes a denominator of 1, as nltk's
    `modified_precision` does with `max(1, sum(counts))`);
  * reference length = per sentence the reference length closest to the hypothesis length (ties -> shorter);
  * brevity penalty exp(1 - r/c) when c <= r (0 when c == 0);
  * BLEU = 0 if there is no unigram match; otherwise BP * exp(sum_n w_n log p_n) with method2 smoothing.
Version-sensitive (the reference pins no nltk version): nltk <= 3.4 (the 2018 era of the reference) adds 1 to numerator
and denominator of EVERY order; later releases leave the unigram precision unsmoothed.  `smooth_unigram` selects
which; the default follows the 2018 behaviour.  Host-side text metric: not on the GPU path.
"""
import math
import os
from collections import Counter
from typing import Dict, List, Sequence


def _ngrams(tokens: Sequence[str], n: int) -> Counter:
    return Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


def modified_precision(references: Sequence[Sequence[str]], hypothesis: Sequence[str], n: int):
    """(clipped matches, max(1, hypothesis n-gram count)) for one sentence."""
    counts = _ngrams(hypothesis, n) if len(hypothesis) >= n else Counter()
    max_counts: Dict[tuple, int] = {}
    for ref in references:
        rc = _ngrams(ref, n) if len(ref) >= n else Counter()
        for ng in counts:
            max_counts[ng] = max(max_counts.get(ng, 0), rc[ng])
    clipped = sum(min(c, max_counts.get(ng, 0)) for ng, c in counts.items())
    return clipped, max(1, sum(counts.values()))


def closest_ref_length(references: Sequence[Sequence[str]], hyp_len: int) -> int:
    return min((len(r) for r in references), key=lambda rl: (abs(rl - hyp_len), rl))


def brevity_penalty(ref_len: int, hyp_len: int) -> float:
    if hyp_len > ref_len:
        return 1.0
    if hyp_len == 0:
        return 0.0
    return math.exp(1.0 - ref_len / hyp_len)


def corpus_bleu(list_of_references, hypotheses, weights=(0.25, 0.25, 0.25, 0.25), smooth_unigram: bool = True) -> float:
    assert len(list_of_references) == len(hypotheses), "one reference set per hypothesis"
    num = [0] * len(weights)
    den = [0] * len(weights)
    hyp_len = ref_len = 0
    for refs, hyp in zip(list_of_references, hypotheses):
        for i in range(len(weights)):
            a, b = modified_precision(refs, hyp, i + 1)
            num[i] += a
            den[i] += b
        hyp_len += len(hyp)
        ref_len += closest_ref_length(refs, len(hyp))
    if num[0] == 0:
        return 0.0
    bp = brevity_penalty(ref_len, hyp_len)
    p = [((num[i] + 1) / (den[i] + 1)) if (i > 0 or smooth_unigram) else (num[i] / den[i]) for i in range(len(weights))]
    return bp * math.exp(math.fsum(w * math.log(pi) for w, pi in zip(weights, p)))


class Eval:
    """`eval.py:12-47`: reads `eval.ids` and `ref.en0..ref.en{n-1}` from `path`."""

    def __init__(self, path: str, n_evals: int) -> None:
        with open(os.path.join(path, "eval.ids"), "r", encoding="utf-8") as f:
            self.ids = [line.strip() for line in f]
        refs: List[List[List[str]]] = []
        for i in range(n_evals):
            with open(os.path.join(path, "ref.en{0:d}".format(i)), "r", encoding="utf-8") as f:
                refs.append([line.strip().split() for line in f])
        self.refs = list(zip(*refs))

    def calc_bleu(self, hyps, smooth_unigram: bool = True) -> float:
        en_hyp = [hyps[u] for u in self.ids]
        return corpus_bleu(self.refs, en_hyp, smooth_unigram=smooth_unigram)

    def write_to_file(self, hyps, fname) -> None:
        with open(fname, "w", encoding="utf-8") as out_f:
            for u in self.ids:
                out_f.write("{0:s}\n".format(" ".join(hyps[u])))
