"""Length-normalised reranking of beam results (beam.py:30-42); the search itself is NN.decode_beam."""
import math


def rerank_hypothesis(beam_hyps, weight):
    """beam.py:30-32: score / (len(hyp) - 2) ** weight, best first."""
    return sorted([(i[0], i[1] / math.pow(len(i[0]) - 2, weight), len(i[0])) for i in beam_hyps],
                  reverse=True, key=lambda t: t[1])


def get_best_hyps(utts_beam, W):
    """beam.py:34-42"""
    preds = {}
    for u in utts_beam:
        rerank_hyp = rerank_hypothesis(utts_beam[u], weight=W)
        preds[u] = [i for i in rerank_hyp[0][0]]
    return preds


def decode_set(nn, set_key, N, K, stop_limit=None):
    """beam.py:105-124: per-utterance decode_beam over a data set -> {utt: [(hyp, score, attn_history)]}."""
    stop_limit = nn.cfg.train["data"]["max_pred"] if stop_limit is None else stop_limit
    beam = {}
    for utt in nn.data_loader.get_batch(1, set_key, train=False, labels=False):
        n_best = nn.decode_beam(utt["X"], stop_limit=stop_limit, N=N, K=K)
        beam[utt["utts"][0]] = [(e["hyp"], e["score"], e["attn_history"]) for e in n_best]
    return beam
