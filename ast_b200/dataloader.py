"""Data loaders with the reference's bucketing and batch protocol (dataloader.py:111-164,
preprocessing/prep_buckets.py:41-63) and a device-side pack: the per-utterance feature rows are copied
once (pinned host -> device) and zero-padded / truncated / frame-dropped (and CMVN-normalised when
statistics are given) by one coalesced CUDA kernel (ast_pack_cmvn) instead of numpy pad_sequence.

get_batch(batch_size, set_key, train, labels) yields {"X": (B,T,D) f32 cuda tensor,
"y": (B,L) i32 cuda tensor, "utts": [...]} exactly like the reference's generator.
"""
import ctypes as C
import os
import zlib
import pickle
import random

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .symbols import SYMBOLS


# ---- preprocessing/prep_buckets.py ---------------------------------------------------------------------
def create_buckets(cat_dict, num_b, width_b, key, scale, seed):
    """prep_buckets.py:41-63: bucket = min(len // width, num_b - 1); optional down-sampling of train."""
    buckets_info = {"buckets": [[] for _ in range(num_b)], "num_b": num_b, "width_b": width_b}
    for utt_id in cat_dict:
        bucket = min(cat_dict[utt_id][key] // width_b, num_b - 1)
        buckets_info["buckets"][bucket].append(utt_id)
    if scale > 1:
        random.seed(seed)
        for i in range(len(buckets_info["buckets"])):
            sample_len = int(len(buckets_info["buckets"][i]) // scale)
            buckets_info["buckets"][i] = random.sample(buckets_info["buckets"][i], sample_len)
    return buckets_info


def buckets_main(save_path, num_b, width_b, key, scale=1, seed="haha", info_path="", info_dict=None):
    """prep_buckets.py:67-108 (also writes buckets_<key>.dict into the model dir when it exists)."""
    if info_dict is None:
        with open(info_path, "rb") as f:
            info_dict = pickle.load(f)
    bucket_dict = {}
    for cat in info_dict:
        scale_val = scale if "train" in cat else 1
        bucket_dict[cat] = create_buckets(info_dict[cat], num_b, width_b, key, scale_val, seed)
    if save_path and os.path.isdir(save_path):
        with open(os.path.join(save_path, "buckets_{0:s}.dict".format(key)), "wb") as f:
            pickle.dump(bucket_dict, f)
    return bucket_dict


def plan_batches(buckets, batch_size, rng=None):
    """dataloader.py:125-135: shuffle inside each bucket, slice by batch_size, shuffle the batches.
    Draws from Python's global `random` in the reference's order; `rng` (a random.Random) replaces it under data
    parallelism, where every rank must draw the SAME plan although their global streams have diverged (each rank's
    forward_loss draws a rank-local number of scheduled-sampling values, seq2seq.py:432)."""
    rng = random if rng is None else rng
    batches = []
    width_b = buckets["width_b"]
    for b, bucket in enumerate(buckets["buckets"]):
        rng.shuffle(bucket)
        for i in range(0, len(bucket), batch_size):
            batches.append((bucket[i:i + batch_size], (b + 1) * width_b))
    rng.shuffle(batches)
    return batches


def drop_frame_mask(n_frames, drop_rate):
    """dataloader.py:83-93: int(rate*n) indices drawn WITH replacement from unseeded np.random."""
    keep = np.ones(n_frames, dtype=np.uint8)
    num_drop = int(drop_rate * n_frames)
    if num_drop > 0:
        keep[np.random.choice(np.arange(n_frames), size=num_drop)] = 0
    return keep


class DevicePacker:
    """Varlen utterances -> padded (B,T,D) batch on the device: ONE pinned staging buffer, ONE async H2D copy, ONE kernel.

    The host side of `dataloader.py:156-162` (`F.pad_sequence` + `to_gpu`) without its per-batch allocations: features,
    row offsets, lengths and the frame-keep masks are written straight into a persistent pinned buffer (two slots, so the
    host can fill batch i+1 while the copy of batch i is in flight), copied with a single `cudaMemcpyAsync`, and expanded
    on the device by the pack kernel (CMVN, frame zeroing and input noise fused; `csrc/misc.cu::pack_cmvn_kernel`).

    With `async_copy=True` (default) the copy and the pack kernel run on the packer's own stream: a training loop that asks
    for batch i+1 right after enqueueing step i (the `get_batch` generator does) gets the next batch's H2D transfer and
    packing overlapped with the running step; the consumer's stream only waits on an event."""

    SLOTS = 2

    def __init__(self, device, async_copy=True):
        self.device = device
        self.lib = _lib.load()
        self.stream = torch.cuda.Stream(device=device) if async_copy else None
        self._host = [None] * self.SLOTS
        self._dev = [None] * self.SLOTS
        self._done = [None] * self.SLOTS
        self._slot = 0

    @staticmethod
    def _up(n, a=16):
        return (n + a - 1) // a * a

    def _buffers(self, slot, nbytes):
        if self._host[slot] is None or self._host[slot].numel() < nbytes:
            if self._done[slot] is not None:
                self._done[slot].synchronize()      # nothing may still read the buffers that are about to be replaced
            cap = max(nbytes * 3 // 2, 1 << 20)
            self._host[slot] = torch.empty(cap, dtype=torch.uint8).pin_memory()
            self._dev[slot] = torch.empty(cap, dtype=torch.uint8, device=self.device)
            self._done[slot] = None
        if self._done[slot] is not None:
            self._done[slot].synchronize()          # the copy that last used this slot (two batches ago) has finished
        return self._host[slot], self._dev[slot]

    def pack(self, utts, max_sp, keep_masks=None, cmvn=None, noise_sigma=0.0, seed=0, labels=None, bits=None):
        """-> X (B,T,D) on the device; with `labels` ((B,L) int32) and optionally `bits` ((L-1,) uint8 scheduled-sampling
        draws) riding on the same copy: -> (X, y_dev, bits_dev)."""
        B, D = len(utts), utts[0].shape[1]
        lens = np.asarray([min(len(u), max_sp) for u in utts], dtype=np.int32)
        T = int(lens.max())
        n_raw = int(lens.sum()) * D * 4
        o_off = self._up(n_raw)
        o_len = o_off + self._up(B * 8)
        o_keep = o_len + self._up(B * 4)
        o_cmvn = o_keep + (self._up(B * T) if keep_masks is not None else 0)
        o_lab = o_cmvn + (self._up(2 * B * D * 4) if cmvn is not None else 0)        # per-utterance (speaker) scale / offset
        n_lab = labels.size * 4 if labels is not None else 0
        o_bits = o_lab + self._up(n_lab)
        total = o_bits + (self._up(len(bits)) if bits is not None else 0)
        slot = self._slot
        self._slot = (slot + 1) % self.SLOTS
        host, dev = self._buffers(slot, total)
        hnp = host.numpy()
        raw = hnp[:n_raw].view(np.float32).reshape(-1, D)
        off = hnp[o_off:o_off + B * 8].view(np.int64)
        r = 0
        for i, u in enumerate(utts):
            n = int(lens[i])
            raw[r:r + n] = u[:n]
            off[i] = r
            r += n
        hnp[o_len:o_len + B * 4].view(np.int32)[:] = lens
        if keep_masks is not None:
            km = hnp[o_keep:o_keep + B * T].reshape(B, T)
            km[:] = 0
            for i, k in enumerate(keep_masks):
                n = min(len(k), T)
                km[i, :n] = k[:n]
        if cmvn is not None:
            cm = hnp[o_cmvn:o_cmvn + 2 * B * D * 4].view(np.float32).reshape(2, B, D)
            cm[0] = np.broadcast_to(np.asarray(cmvn[0], dtype=np.float32), (B, D))
            cm[1] = np.broadcast_to(np.asarray(cmvn[1], dtype=np.float32), (B, D))
        if labels is not None:
            hnp[o_lab:o_lab + n_lab].view(np.int32)[:] = np.asarray(labels, dtype=np.int32).ravel()
        if bits is not None:
            hnp[o_bits:o_bits + len(bits)] = np.asarray(bits, dtype=np.uint8)
        consumer = torch.cuda.current_stream(self.device)
        work = self.stream if self.stream is not None else consumer
        with torch.cuda.stream(work):
            dev[:total].copy_(host[:total], non_blocking=True)
            d_raw = dev[:n_raw].view(torch.float32)
            d_off = dev[o_off:o_off + B * 8].view(torch.int64)
            d_len = dev[o_len:o_len + B * 4].view(torch.int32)
            keep = dev[o_keep:o_keep + B * T] if keep_masks is not None else None
            scale = offset = None
            if cmvn is not None:
                d_cm = dev[o_cmvn:o_cmvn + 2 * B * D * 4].view(torch.float32)
                scale, offset = d_cm[:B * D], d_cm[B * D:]
            X = torch.empty(B, T, D, dtype=torch.float32, device=self.device)
            check(self.lib.ast_pack_cmvn(ptr(d_raw), ptr(d_off), ptr(d_len), ptr(scale), ptr(offset), ptr(keep), None,
                                         float(noise_sigma), int(seed), ptr(X), B, T, D, C.c_void_p(work.cuda_stream)), "ast_pack_cmvn")
            outs = [X]
            if labels is not None:
                # the staging slot is recycled two batches later: hand out copies of the (tiny) label / bit tensors
                outs.append(dev[o_lab:o_lab + n_lab].view(torch.int32).reshape(labels.shape).clone())
                outs.append(dev[o_bits:o_bits + len(bits)].clone() if bits is not None else None)
            ev = torch.cuda.Event()
            ev.record(work)
        self._done[slot] = ev
        if work is not consumer:
            consumer.wait_event(ev)
            for t in outs:
                if t is not None:
                    t.record_stream(consumer)
        return X if labels is None else tuple(outs)


def cmvn_scale_offset(stats_sum, stats_sumsq, count, norm_vars=True):
    """Kaldi apply-cmvn coefficients from accumulated statistics (double precision, as Kaldi):
    y = x*scale + offset, scale = 1/sqrt(var) (var floored at 1e-20), offset = -mean*scale."""
    mean = np.asarray(stats_sum, dtype=np.float64) / count
    if norm_vars:
        var = np.maximum(np.asarray(stats_sumsq, dtype=np.float64) / count - mean * mean, 1e-20)
        scale = 1.0 / np.sqrt(var)
    else:
        scale = np.ones_like(mean)
    return scale.astype(np.float32), (-mean * scale).astype(np.float32)


class DataLoader:
    def __init__(self):
        self.map, self.vocab, self.info = {}, {}, {}

    def get_batch(self, batch_size, set_key, train=True, labels=False):
        raise NotImplementedError


class _BucketedLoader(DataLoader):
    """Shared get_batch / get_hyps logic of FisherDataLoader and GlobalPhoneDataLoader."""

    def _finish_init(self, data_cfg, model_dir, gpuid, info_dict=None):
        self.gpuid = gpuid
        self.data_cfg = data_cfg
        self.model_dir = model_dir
        self.buckets = buckets_main(self.model_dir, data_cfg["buckets_num"], data_cfg["buckets_width"], key="sp",
                                    scale=data_cfg["train_scale"], seed="haha", info_path=data_cfg.get("info_path", ""),
                                    info_dict=info_dict)
        self.n_utts = {k: sum(len(b) for b in self.buckets[k]["buckets"]) for k in self.buckets}
        self._packer = None

    def _load_utt(self, utt, set_key):
        raise NotImplementedError

    @property
    def feat_dim(self):
        """Feature dimension D of the corpus, read from one utterance (the reference shapes its first layers lazily at
        the first batch: `in_channels: null`, `L.LSTM(None, ...)`; NN needs D to build the engine up front)."""
        if getattr(self, "_feat_dim", None) is None:
            for set_key in self.buckets:
                for bucket in self.buckets[set_key]["buckets"]:
                    if bucket:
                        self._feat_dim = int(np.asarray(self._load_utt(bucket[0], set_key)).shape[-1])
                        return self._feat_dim
            raise RuntimeError("empty corpus: cannot infer the feature dimension")
        return self._feat_dim

    def _labels(self, utt, set_key):
        dec_key = self.data_cfg["dec_key"]
        return [self.vocab[dec_key]["w2i"].get(w, SYMBOLS.UNK_ID) for w in self.map[set_key][utt][dec_key]]

    def host_batches(self, batch_size, set_key, train, labels=False, rank=0, world=1, plan_rng=None):
        """The host half of dataloader.py:111-164 (pure numpy, no device): batch plan with the reference's `random` draw
        order, per-utterance load, frame-zeroing masks drawn from numpy's global RNG in load order (:83-93,105-106), labels
        [GO] + ids[:max_pred-2] + [EOS] zero-padded to the batch maximum (:149-150,160).
        Yields (utts, feats [list of (T_i, D)], keep [list of uint8 masks] | None, y (B, L) int32 | None, max_sp).
        Data parallel (`world` > 1): the plan is drawn for the GLOBAL batch (batch_size * world, from `plan_rng`,
        identical on every rank) and rank r takes utterances r::world of each global batch - same bucket on every rank, so
        padded lengths stay balanced.  A tail batch smaller than the world leaves some ranks with an EMPTY shard (utts == []):
        they still take part in the gradient all-reduce with a zero contribution (NN.train_epoch)."""
        num_b, width_b = self.buckets[set_key]["num_b"], self.buckets[set_key]["width_b"]
        max_sp = (num_b + 1) * width_b                                  # dataloader.py:118
        max_pred = self.data_cfg["max_pred"]
        zero_input = self.data_cfg.get("zero_input", 0)
        for gutts, _ in plan_batches(self.buckets[set_key], batch_size * world, plan_rng):
            utts = list(gutts[rank::world]) if world > 1 else gutts
            self.last_global_batch = len(gutts)
            if not utts:
                yield [], [], None, None, max_sp
                continue
            feats = [self._load_utt(u, set_key) for u in utts]
            keep = None
            if "train" in set_key and zero_input > 0:                   # dataloader.py:105-106
                keep = [drop_frame_mask(min(len(f), max_sp), zero_input) for f in feats]
            ypad = None
            if labels:
                ys = [np.asarray([SYMBOLS.GO_ID] + self._labels(u, set_key)[:max_pred - 2] + [SYMBOLS.EOS_ID], dtype=np.int32)
                      for u in utts]
                L = max(len(v) for v in ys)
                ypad = np.zeros((len(ys), L), dtype=np.int32)           # PAD_ID = 0
                for i, v in enumerate(ys):
                    ypad[i, :len(v)] = v
            yield list(utts), feats, keep, ypad, max_sp

    def get_batch(self, batch_size, set_key, train, labels=False, rank=0, world=1, plan_rng=None):
        dev = torch.device("cuda", self.gpuid if self.gpuid is not None and self.gpuid >= 0 else 0)
        if self._packer is None:
            self._packer = DevicePacker(dev)
        for utts, feats, keep, ypad, max_sp in self.host_batches(batch_size, set_key, train, labels, rank, world, plan_rng):
            batch = {"utts": utts, "global_batch": self.last_global_batch}
            if not utts:                       # empty data-parallel shard of a tail batch
                batch["X"] = batch["y"] = None
                yield batch
                continue
            if labels:
                batch["X"], batch["y"], _ = self._packer.pack(feats, max_sp, keep, labels=ypad)   # labels ride the same copy
            else:
                batch["X"] = self._packer.pack(feats, max_sp, keep)
            yield batch

    def get_hyps(self, preds):
        """dataloader.py:167-183: ids -> words (ids < 4 dropped, BPE '@@ ' joins undone)."""
        dec_key = self.data_cfg["dec_key"]
        join_str = " " if dec_key.endswith("_w") else ""
        en_hyps = {}
        for utt, p in preds:
            en_hyps[utt] = []
            if type(p) == list:
                t_str = join_str.join([self.vocab[dec_key]["i2w"][i].decode() for i in p if i >= len(SYMBOLS.START_VOCAB)])
                if "bpe_w" in dec_key:
                    t_str = t_str.replace("@@ ", "")
                en_hyps[utt].extend(t_str.strip().split())
        return en_hyps


class FisherDataLoader(_BucketedLoader):
    """dataloader.py:49-183: per-utterance .npy files under speech_path/<set>/[<spk>/]<utt>.npy."""

    def __init__(self, data_cfg, model_dir, gpuid):
        super().__init__()
        with open(data_cfg["map_path"], "rb") as f:
            self.map = pickle.load(f)
        with open(data_cfg["vocab_path"], "rb") as f:
            self.vocab = pickle.load(f)
        with open(data_cfg["info_path"], "rb") as f:
            self.info = pickle.load(f)
        self._finish_init(data_cfg, model_dir, gpuid, info_dict=self.info)

    def _load_utt(self, utt, set_key):
        sp_path = os.path.join(self.data_cfg["speech_path"], set_key)
        utt_path = os.path.join(sp_path, "{0:s}.npy".format(utt))
        if not os.path.exists(utt_path):
            utt_path = os.path.join(sp_path, utt.split("_", 1)[0], "{0:s}.npy".format(utt))
        return np.load(utt_path)


class GlobalPhoneDataLoader(_BucketedLoader):
    """dataloader.py:185-316: all speech in one pickle {set: {utt: (T,D) array}}."""

    def __init__(self, data_cfg, model_dir, gpuid):
        super().__init__()
        with open(data_cfg["map_path"], "rb") as f:
            self.map = pickle.load(f)
        with open(data_cfg["vocab_path"], "rb") as f:
            self.vocab = pickle.load(f)
        with open(data_cfg["info_path"], "rb") as f:
            self.info = pickle.load(f)
        with open(data_cfg["speech_path"], "rb") as f:
            self.speech_data = pickle.load(f)
        self._finish_init(data_cfg, model_dir, gpuid, info_dict=self.info)

    def _load_utt(self, utt, set_key):
        return np.asarray(self.speech_data[set_key][utt])


class SyntheticDataLoader(_BucketedLoader):
    """Fisher-shaped synthetic corpus for benchmarks and tests: utterance
    lengths and target lengths are given (or resampled from the bucket histogram), features are
    N(0,1) (what CMVN'd features look like), tokens uniform over [4, V)."""

    # Appendix C: fisher_train bucket counts (width 80, 20 buckets)
    FISHER_20H_BUCKETS = [1025, 3516, 2543, 1939, 1486, 1188, 932, 736, 674, 603, 550, 505, 420, 342, 277, 189, 138, 86, 63, 94]

    def __init__(self, data_cfg, model_dir, gpuid, feat_dim, vocab_size, lengths, target_lengths, seed=0, set_key="fisher_train"):
        super().__init__()
        self._feat_dim, self.vocab_size = feat_dim, vocab_size
        rng = np.random.default_rng(seed)
        self._seed = seed
        names = ["utt{0:06d}".format(i) for i in range(len(lengths))]
        self.lengths = dict(zip(names, (int(x) for x in lengths)))
        info = {set_key: {u: {"sp": self.lengths[u]} for u in names}}
        self._labels_cache = {u: rng.integers(4, vocab_size, size=max(int(n) - 2, 0)).tolist() for u, n in zip(names, target_lengths)}
        self._finish_init(data_cfg, model_dir, gpuid, info_dict=info)

    @classmethod
    def fisher_shaped(cls, data_cfg, model_dir, gpuid, feat_dim, vocab_size, n_utts=17306, seed=0, set_key="fisher_train"):
        rng = np.random.default_rng(seed)
        counts = np.asarray(cls.FISHER_20H_BUCKETS, dtype=np.float64)
        b = rng.choice(len(counts), size=n_utts, p=counts / counts.sum())
        lens = b * 80 + rng.integers(0, 80, size=n_utts)
        lens = np.maximum(lens, 27)
        lens[b == len(counts) - 1] += rng.integers(0, 400, size=int((b == len(counts) - 1).sum()))
        # target length ~ 1.3 BPE per word, words ~ frames/42 (Appendix C medians: 302 frames <-> 7 words)
        words = np.maximum(1, np.round(lens / 42.0 * rng.uniform(0.6, 1.4, size=n_utts))).astype(int)
        tl = np.minimum(np.round(1.3 * words).astype(int) + 2, data_cfg["max_pred"])
        return cls(data_cfg, model_dir, gpuid, feat_dim, vocab_size, lens, tl, seed=seed, set_key=set_key)

    def _load_utt(self, utt, set_key):
        n = self.lengths[utt]
        rng = np.random.default_rng(zlib.crc32(utt.encode()) ^ self._seed)
        return rng.standard_normal((n, self._feat_dim), dtype=np.float32)

    def _labels(self, utt, set_key):
        return self._labels_cache[utt]

    def true_frames(self, utts, set_key="fisher_train"):
        max_sp = (self.buckets[set_key]["num_b"] + 1) * self.buckets[set_key]["width_b"]
        return sum(min(self.lengths[u], max_sp) for u in utts)
