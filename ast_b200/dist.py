"""Data-parallel training over the 8 B200s of one box: one process per GPU, full replica per
rank, disjoint slice of every global batch, one NCCL sum all-reduce of the flat gradient buffer
(14.4 M fp32 = 57.7 MB, ~0.2 ms over NVLink 5 / NVSwitch) folded into the optimizer as
``grad_scale = 1/world_size``; WeightDecay -> global-norm clip -> AMSGrad then run identically on every
rank, the clip norm being that of the REDUCED gradient.  BatchNorm statistics stay per replica (the
reference has no DP; parity is defined per replica).

The reference has no collective anywhere; this is new work required by BASELINE config 4.
The host-side logic (sharding, bucket order, reduction arithmetic) is backend-agnostic and is covered by
world_size-2 gloo tests on CPU.
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    """Join the job described by RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_utterances(utts, rank, world):
    """Rank r takes utterances r::G of a global batch: every rank sees the same bucket, so the
    padded lengths - and step times - stay balanced."""
    return list(utts[rank::world])


def shard_batch_plan(batches, rank, world):
    """Global batch plan [(utts, width)] built with batch_size*world -> this rank's plan [(utts, width, n_global)].
    Every rank must have drawn the same plan: use a dedicated `random.Random` (dataloader.plan_batches(rng=)), the global
    `random` streams of the ranks diverge after one step.  A tail batch smaller than the world leaves ranks with an EMPTY
    shard; they still join the all-reduce with weight 0 (`shard_weight`) - no utterance is ever counted twice."""
    return [(shard_utterances(utts, rank, world), width, len(utts)) for utts, width in batches]


def shard_weight(n_local, n_global, world):
    """Factor a rank applies to its local gradient BEFORE the sum all-reduce so that the reduced gradient, multiplied by
    the optimizer's grad_scale = 1/world, is the gradient of the global-batch mean loss: each replica's loss is a mean
    over its OWN n_local utterances (seq2seq.py:468, divisor = local batch), so the global mean is
    sum_r (n_r / n_global) g_r = (1/world) sum_r w_r g_r with w_r = n_r * world / n_global (1 for equal shards)."""
    return float(n_local) * world / float(n_global) if n_global else 0.0


def allreduce_sum_(flat, group=None):
    """In-place sum all-reduce of a flat fp32 tensor (gradient bucket)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class GradAllReduce:
    """optimizer.pre_update hook: sum all-reduce of the flat gradient buffer in the three contiguous buckets the library
    completes in backward order (``Engine.grad_buckets()``: decoder 56 % of the bytes, encoder, CNN).  With
    ``overlap=True`` the collectives run on their own stream, each one gated on the event ``ast_backward`` recorded when
    that bucket became final (``Engine.grad_bucket_wait``), so the decoder bucket crosses NVLink while the encoder and
    CNN backward are still running; the optimizer's stream then waits for the communication stream.  ``overlap=False``
    (and any engine without buckets, e.g. the CPU gloo tests) reduces the same ranges on the caller's stream."""

    def __init__(self, engine, optimizer, world, overlap=True):
        self.e, self.opt, self.world = engine, optimizer, world
        self.overlap = bool(overlap) and world > 1 and engine.grads.is_cuda
        self.comm = torch.cuda.Stream(device=engine.grads.device) if self.overlap else None
        self.buckets = list(engine.grad_buckets()) if hasattr(engine, "grad_buckets") else [(0, engine.grads.numel())]
        covered = sorted(self.buckets)
        assert covered[0][0] == 0 and covered[-1][0] + covered[-1][1] == engine.grads.numel() and \
            all(a[0] + a[1] == b[0] for a, b in zip(covered, covered[1:])), "gradient buckets must tile the flat buffer"
        optimizer.grad_scale = 1.0 / world
        optimizer.pre_update = self

    def __call__(self):
        if self.world <= 1:
            return
        g = self.e.grads
        if not self.overlap:
            for off, cnt in self.buckets:
                allreduce_sum_(g[off:off + cnt])
            return
        main = torch.cuda.current_stream(g.device)
        with torch.cuda.stream(self.comm):
            for i, (off, cnt) in enumerate(self.buckets):
                self.e.grad_bucket_wait(i, self.comm)
                allreduce_sum_(g[off:off + cnt])
        main.wait_stream(self.comm)


def broadcast_params_(engine, src=0):
    """Make every replica start from rank `src`'s parameters and BN running statistics."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(engine.params, src)
        dist.broadcast(engine.bn_state, src)
        engine.weights_changed()


def max_over_ranks(value, device):
    """Timing rule: every multi-GPU number is the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
