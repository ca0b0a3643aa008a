"""world_size-2 gloo test of the data-parallel host logic: the single gradient/loss all-reduce and
the per-rank RNG fold-in (trainer/general_diffusion_trainer.py:251-253,325,334)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flaxdiff_b200 import utils
    from flaxdiff_b200.trainer import dp_allreduce_sum_, rank_key
    g = torch.full((1024,), float(rank + 1))
    loss = torch.tensor([float(rank)])
    scale = dp_allreduce_sum_(g, loss, world)
    key = rank_key(utils.PRNGKey(4), rank)
    q.put((rank, float((g * scale)[0]), float(loss[0] * scale), key))
    dist.destroy_process_group()


def test_grad_allreduce_mean_and_rank_keys():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == 1.5            # mean of (1, 2)
    assert res[0][2] == res[1][2] == 0.5            # mean loss
    assert res[0][3] != res[1][3]                   # distinct per-rank RNG streams
