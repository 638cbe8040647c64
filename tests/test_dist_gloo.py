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


def _worker_state(rank, world, port, q):
    """FlatTrainState + broadcast + weighted all-reduce on CPU modules (host logic of HotPathTrainer at world size 2)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dpmn_b200.dist import FlatGradBucket, FlatTrainState, broadcast_module_state, shard_range
    torch.manual_seed(100 + rank)                      # ranks build DIFFERENT replicas ...
    mods = [torch.nn.Linear(5, 3), torch.nn.Sequential(torch.nn.Linear(3, 7), torch.nn.BatchNorm1d(7))]
    broadcast_module_state(mods, src=0)                # ... and are made identical here (parameters AND buffers)
    flat0 = torch.cat([t.detach().reshape(-1).float() for m in mods for t in list(m.parameters()) + list(m.buffers())])
    gathered = [torch.zeros_like(flat0) for _ in range(world)]
    dist.all_gather(gathered, flat0)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    st = FlatTrainState(mods)
    # segments: 16-byte aligned starts, parameters and gradients are views of the flat buffers
    ok_layout = st.offsets[0] == 0 and all(o % 64 == 0 for o in st.offsets) and st.n_params == 18 + 28 + 14
    ok_layout = ok_layout and all((p.data_ptr() - st.flat_params.data_ptr()) % 256 == 0 and
                                   (p.grad.data_ptr() - st.flat_grads.data_ptr()) % 256 == 0 for p in st.params)
    p0 = mods[0].weight
    p0.data.fill_(3.0)
    ok_views = bool(torch.all(st.flat_params[:15] == 3.0)) and float(st.flat_params[15:64].abs().max()) == 0.0 \
        and p0.grad.data_ptr() == st.flat_grads.data_ptr()
    ok_sink = set(mods[0]._grad_sink) == {"weight", "bias"} and mods[1]._grad_sink["0.weight"].shape == (7, 3)
    # uneven shards: global batch 5 -> 3 + 2 images; each rank holds the gradient of its LOCAL mean
    lo, hi = shard_range(5, rank, world)
    per_image = torch.arange(5, dtype=torch.float32)            # "gradient" contributed by image i is i
    local_mean = per_image[lo:hi].mean()
    b = FlatGradBucket([torch.nn.Parameter(torch.zeros(4))])
    b.params[0].grad.fill_(float(local_mean))
    b.allreduce_mean(local_items=hi - lo, global_items=5)
    ok_weighted = bool(torch.allclose(b.params[0].grad, torch.full((4,), float(per_image.mean()))))
    q.put((rank, same, ok_layout, ok_views, ok_sink, ok_weighted))
    dist.destroy_process_group()


def test_flat_train_state_broadcast_and_weighted_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_state, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r


def test_pgrm_mask_seed_is_salted_by_rank():
    """ranks that seed torch identically must still draw different Dropout / DropPath masks."""
    from dpmn_b200.pgrm import PGRM
    seeds = []
    for salt in (0, 1, 2):
        torch.manual_seed(7)
        m = PGRM.__new__(PGRM)
        torch.nn.Module.__init__(m)
        m._seed_salt = salt
        seeds.append(m._new_seed())
    assert len(set(seeds)) == 3 and all(0 <= s < 2 ** 62 for s in seeds)
