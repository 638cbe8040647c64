"""N > 1 host logic on CPU: world_size-2 gloo process group (the GPU box uses NCCL with the same code)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dpmn_b200.dist import FlatGradBucket, max_over_ranks, shard_range
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)),
              torch.nn.Parameter(torch.zeros(2), requires_grad=False)]
    bucket = FlatGradBucket(params)
    assert bucket.flat.numel() == 22
    for i, p in enumerate(bucket.params):
        p.grad.fill_(float((rank + 1) * (i + 1)))        # views: writes land in the flat buffer
    bucket.allreduce_mean()
    want0, want1 = (1 + 2) / 2 * 1, (1 + 2) / 2 * 2
    ok = bool(torch.all(params[0].grad == want0)) and bool(torch.all(params[1].grad == want1)) and params[2].grad is None
    lo, hi = shard_range(97, rank, world)
    slow = max_over_ranks(10.0 + rank, "cpu")
    q.put((rank, ok, lo, hi, slow))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 49, 49, 97)     # contiguous, exhaustive shards
    assert res[0][4] == res[1][4] == 11.0


def test_shard_range_covers_batch_exactly():
    from dpmn_b200.dist import shard_range
    for n in (48, 97, 384):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
