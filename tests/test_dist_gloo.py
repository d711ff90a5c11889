"""N>1 host path on CPU: world_size-2 gloo processes shard frames with no data-path collective and
reduce timings with max / counters with sum (bench.py's multi-GPU contract)."""
import os
import socket

import torch.multiprocessing as mp

from com_b200 import dist as cdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = cdist.init(backend="gloo")
    mine = cdist.frame_shard(7, r, w)
    cdist.barrier()
    t = cdist.max_over_ranks(10.0 + r, device="cpu")
    n = cdist.sum_over_ranks(len(mine), device="cpu")
    # FlatGradSync: one flat all-reduce == the mean of the ranks' gradients (what DistributedDataParallel produces);
    # parameters without a gradient on this rank count as zeros; a single-element and a bf16 parameter ride along
    import torch
    import torch.distributed as dist
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s_)) for s_ in ((3, 4), (5,), (1,))] + [torch.nn.Parameter(torch.zeros((2, 2), dtype=torch.bfloat16))]
    gen = torch.Generator().manual_seed(100 + r)
    for i, p in enumerate(params):
        if not (r == 1 and i == 1):
            p.grad = torch.randn(p.shape, generator=gen).to(p.dtype)
    mine_g = [None if p.grad is None else p.grad.clone().float() for p in params]
    cdist.FlatGradSync(params)()
    gathered = [None, None]
    dist.all_gather_object(gathered, mine_g)
    ok = True
    for i, p in enumerate(params):
        want = sum((g[i] if g[i] is not None else torch.zeros(p.shape)) for g in gathered) / w
        ok = ok and p.grad is not None and p.grad.dtype == p.dtype and torch.allclose(p.grad.float(), want, atol=1e-2 if i == 3 else 1e-6)
    # in-place form: the gradients are views of one flat buffer (what the fused train step's backward graph hands out)
    flat = torch.randn(20, generator=gen)
    mine_f = flat.clone()
    ps = [torch.nn.Parameter(torch.zeros(s_)) for s_ in ((3, 4), (8,))]
    ps[0].grad, ps[1].grad = flat[:12].view(3, 4), flat[12:]
    sync = cdist.FlatGradSync(ps, flat_provider=lambda: flat)
    sync()
    both = [None, None]
    dist.all_gather_object(both, mine_f)
    ok = ok and sync.in_place and torch.allclose(ps[0].grad.reshape(-1), ((both[0] + both[1]) / w)[:12]) \
        and torch.allclose(ps[1].grad, ((both[0] + both[1]) / w)[12:])
    q.put((r, mine, t, n, ok))
    dist.destroy_process_group()


def test_frame_shard_partition():
    for world in (1, 2, 3, 8):
        got = sorted(sum((cdist.frame_shard(13, r, world) for r in range(world)), []))
        assert got == list(range(13))
    assert cdist.frame_shard(0, 0, 2) == [] and cdist.frame_shard(1, 1, 2) == []


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    assert res[0][2] == res[1][2] == 11.0           # max over ranks
    assert res[0][3] == res[1][3] == 7.0            # every frame processed exactly once
    assert res[0][4] and res[1][4]                  # flat gradient all-reduce == mean over ranks


def test_single_process_is_a_noop():
    assert cdist.max_over_ranks(3.5) == 3.5 and cdist.sum_over_ranks(2) == 2
