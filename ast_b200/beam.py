"""Length-normalised reranking of beam results (beam.py:30-42); the search itself is NN.decode_beam.  BeamPool decodes
independent utterances concurrently (the reference's beam.py:105-124 loop is sequential and batch-size-1)."""
import math
import queue
import threading

from .symbols import SYMBOLS


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


def decode_set(nn, set_key, N, K, stop_limit=None, in_flight=1, batch=0):
    """beam.py:105-124: per-utterance decode_beam over a data set -> {utt: [(hyp, score, attn_history)]}.
    batch > 1 searches that many utterances in lock-step on the device (NN.decode_beam_batch: one pass over the decoder weights
    per step for all of them); in_flight > 1 decodes that many utterances concurrently on engine replicas (BeamPool).
    Either way every utterance gets the hypotheses of the sequential loop."""
    stop_limit = nn.cfg.train["data"]["max_pred"] if stop_limit is None else stop_limit
    beam = {}
    if batch > 1:
        pend = []

        def flush():
            for b, n_best in zip(pend, nn.decode_beam_batch([b["X"] for b in pend], stop_limit, N, K)):
                beam[b["utts"][0]] = [(e["hyp"], e["score"], e["attn_history"]) for e in n_best]
            pend.clear()
        for utt in nn.data_loader.get_batch(1, set_key, train=False, labels=False):
            pend.append(utt)
            if len(pend) == min(batch, 32):
                flush()
        if pend:
            flush()
        return beam
    if in_flight > 1:
        from .nn import beam_result_to_entries, using_config
        batches = list(nn.data_loader.get_batch(1, set_key, train=False, labels=False))
        with using_config("train", False):
            nn.model._require(batches[0]["X"])
            pool = BeamPool(nn.model, n=in_flight)
            res = pool.decode([b["X"] for b in batches], stop_limit, N, K, SYMBOLS.GO_ID, SYMBOLS.EOS_ID,
                              convert=lambda r: beam_result_to_entries(r, SYMBOLS.GO_ID) if r["n_steps"] > 0 else [nn.init_hyp()])
        for b, n_best in zip(batches, res):
            beam[b["utts"][0]] = [(e["hyp"], e["score"], e["attn_history"]) for e in n_best]
        return beam
    for utt in nn.data_loader.get_batch(1, set_key, train=False, labels=False):
        n_best = nn.decode_beam(utt["X"], stop_limit=stop_limit, N=N, K=K)
        beam[utt["utts"][0]] = [(e["hyp"], e["score"], e["attn_history"]) for e in n_best]
    return beam


class BeamPool:
    """Throughput mode of beam decoding: `n` replicas of the model's engine on one GPU (same weights, own workspace and CUDA
    streams), each driven by its own host thread, decode independent utterances at the same time.  One beam-10 search is a chain
    of small latency-bound kernels (10 hypotheses x one decoder step) that leaves most of the GPU idle; utterances are independent
    ("replicas only"), so several searches interleave on the SMs.  Every search is the same code path as
    `NN.decode_beam` - hypotheses and scores are identical to the sequential loop."""

    def __init__(self, model, n=4):
        import torch
        from .engine import Engine
        src = model._require() if hasattr(model, "_require") else model
        self.engines, self.streams = [src], [torch.cuda.Stream(device=src.device)]
        for _ in range(max(1, n) - 1):
            e = Engine(src.cfg, src.feat_dim, src.device.index or 0)
            for k in ("exact", "tc_gemm", "beam_fused", "dec_fused"):
                e.set_option(k, src.get_option(k))
            self.engines.append(e)
            self.streams.append(torch.cuda.Stream(device=src.device))
        self.sync_weights()

    def sync_weights(self):
        """Copy the source model's parameters and BatchNorm running statistics into the replicas."""
        import torch
        src = self.engines[0]
        torch.cuda.current_stream(src.device).synchronize()
        for e in self.engines[1:]:
            e.params.copy_(src.params)
            e.bn_state.copy_(src.bn_state)
            e.bn_N = list(src.bn_N)
            e.weights_changed()
        torch.cuda.current_stream(src.device).synchronize()

    def decode(self, utterances, stop_limit, N, K, go=1, eos=2, convert=None):
        """utterances: list of (1, T, D) float32 arrays / tensors -> list of results in the same order (`convert(r)` applied in
        the worker thread if given, e.g. nn.beam_result_to_entries)."""
        import torch
        jobs = queue.Queue()
        for i, x in enumerate(utterances):
            jobs.put((i, x))
        out = [None] * len(utterances)
        errs = []

        def worker(e, st):
            try:
                with torch.cuda.stream(st):
                    while True:
                        try:
                            i, x = jobs.get_nowait()
                        except queue.Empty:
                            break
                        r = e.beam_search(x, int(stop_limit), int(N), int(K), go, eos)
                        out[i] = convert(r) if convert else r
                    st.synchronize()
            except Exception as ex:     # surface in the caller
                errs.append(ex)

        # the utterances may still be in flight on other streams (e.g. the loader's pack stream): the workers' streams do not
        # know about them
        torch.cuda.synchronize(self.engines[0].device)
        ts = [threading.Thread(target=worker, args=(e, st)) for e, st in zip(self.engines, self.streams)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]
        return out
