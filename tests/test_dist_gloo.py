"""world_size-2 gloo test (CPU) of the data-parallel host logic: identical batch plans on every rank,
disjoint utterance shards, sum all-reduce + 1/world scaling equals the big-batch mean gradient."""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ast_b200 import dist as adist
    from ast_b200.dataloader import create_buckets, plan_batches
    r, lr, w = adist.init_process_group("gloo")
    assert (r, w) == (rank, world)
    info = {f"u{i:03d}": {"sp": 30 + 7 * i} for i in range(40)}
    random.seed("seed-ast-20h")                     # every rank seeds identically (nn.py:54) ...
    for _ in range(rank * 3):
        random.random()                             # ... but the global streams diverge (rank-local scheduled-sampling draws)
    plan_rng = random.Random("seed-ast-20h/plan")   # the plan therefore comes from its own generator (NN._plan_rng)
    plan = plan_batches(create_buckets(info, 20, 80, "sp", 1, "haha"), 4 * world, plan_rng)
    plan2 = plan_batches(create_buckets(info, 20, 80, "sp", 1, "haha"), 4 * world, plan_rng)     # next epoch
    mine = adist.shard_batch_plan(plan, rank, world)
    assert len(mine) == len(plan)
    # weighted reduction of a ragged global batch: replica-mean gradients g_r over n_r utterances
    n_glob = 3
    n_loc = len(list(range(n_glob))[rank::world])
    per_utt = torch.arange(1.0, n_glob + 1.0)
    gl = per_utt[rank::world].mean().reshape(1) if n_loc else torch.zeros(1)
    gl = gl * adist.shard_weight(n_loc, n_glob, world)
    adist.allreduce_sum_(gl)
    ragged = float(gl) / world
    # gradients: rank-local mean gradients -> sum all-reduce -> x 1/world
    rng = np.random.default_rng(rank)
    g = torch.tensor(rng.standard_normal(1000), dtype=torch.float32)
    local = g.clone()
    adist.allreduce_sum_(g)
    # the bucketed hook (same ranges the library reports on the GPU: CNN | encoder | decoder, reduced decoder-first)
    class _Opt:
        grad_scale = 1.0
        pre_update = None

    class _Eng:
        grads = torch.tensor(np.random.default_rng(10 + rank).standard_normal(1000), dtype=torch.float32)

        def grad_buckets(self):
            return [(600, 400), (100, 500), (0, 100)]

    eng, opt = _Eng(), _Opt()
    blocal = eng.grads.clone()
    hook = adist.GradAllReduce(eng, opt, world)
    assert opt.pre_update is hook and opt.grad_scale == 1.0 / world and not hook.overlap
    hook()
    tmax = adist.max_over_ranks(float(rank + 1), torch.device("cpu"))
    tsum = adist.sum_over_ranks(float(rank + 1), torch.device("cpu"))
    torch.save({"plan": plan, "plan2": plan2, "ragged": ragged, "mine": mine, "local": local, "reduced": g, "blocal": blocal, "breduced": eng.grads, "tmax": tmax, "tsum": tsum}, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_dp_host_logic_world2(tmp_path):
    world, port = 2, 29000 + random.randrange(2000)
    out = str(tmp_path / "r")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = [torch.load(out + f".{r}", weights_only=False) for r in range(world)]
    assert res[0]["plan"] == res[1]["plan"] and res[0]["plan2"] == res[1]["plan2"] and res[0]["plan"] != res[0]["plan2"]
    tails = 0
    for b in range(len(res[0]["plan"])):
        utts = res[0]["plan"][b][0]
        s0, s1 = res[0]["mine"][b][0], res[1]["mine"][b][0]
        assert sorted(s0 + s1) == sorted(utts) and not set(s0) & set(s1)        # a partition: nothing dropped, nothing twice
        assert res[0]["mine"][b][2] == res[1]["mine"][b][2] == len(utts)
        tails += len(s0) != len(s1)
    assert tails > 0                                                            # ragged tail batches were exercised
    assert abs(res[0]["ragged"] - 2.0) < 1e-6 and abs(res[1]["ragged"] - 2.0) < 1e-6   # mean of (1, 2, 3) over the GLOBAL batch
    want = res[0]["local"] + res[1]["local"]
    assert torch.allclose(res[0]["reduced"], want) and torch.allclose(res[1]["reduced"], want)
    bwant = res[0]["blocal"] + res[1]["blocal"]
    assert torch.allclose(res[0]["breduced"], bwant) and torch.allclose(res[1]["breduced"], bwant)
    assert res[0]["tmax"] == 2.0 and res[1]["tsum"] == 3.0
