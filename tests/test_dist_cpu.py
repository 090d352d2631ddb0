"""World-size-2 gloo tests of the multi-GPU host logic (vidtok_b200/dist.py): contiguous clip sharding, the PSNR
partial-sum all-reduce and the reconstruction all-gather give the single-process answer."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from vidtok_b200 import dist as vdist
    r, w, _ = vdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(n_clips, 3, 5, 8, 8, generator=g) * 2 - 1
    y = (x + 0.1 * torch.randn(x.shape, generator=g)).clamp(-1.2, 1.2)
    s, e = vdist.shard_range(n_clips, rank, world)
    part = vdist.psnr_partial(x[s:e], y[s:e])
    psnr = vdist.global_psnr(part)
    counts = [vdist.shard_range(n_clips, k, world)[1] - vdist.shard_range(n_clips, k, world)[0] for k in range(world)]
    gathered = vdist.gather_clips(y[s:e].contiguous(), counts)
    t = vdist.allreduce_max(torch.tensor([float(rank + 1)], dtype=torch.float64))
    q.put((rank, psnr, bool(torch.equal(gathered, y)), float(t[0]), (s, e)))
    vdist.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5])
def test_sharded_psnr_and_gather_world2(n_clips):
    from vidtok_b200 import dist as vdist
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(n_clips, 3, 5, 8, 8, generator=g) * 2 - 1
    y = (x + 0.1 * torch.randn(x.shape, generator=g)).clamp(-1.2, 1.2)
    single = vdist.psnr_partial(x, y)
    want = float(single[0] / single[1])
    for rank, psnr, gathered_ok, tmax, (s, e) in res:
        assert abs(psnr - want) < 1e-7 and gathered_ok and tmax == 2.0
    assert res[0][4][0] == 0 and res[0][4][1] == res[1][4][0] and res[1][4][1] == n_clips


def test_shard_range_properties():
    from vidtok_b200.dist import shard_range
    for n in (0, 1, 7, 8, 32, 33):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in rs]
            assert max(sizes) - min(sizes) <= 1


def test_psnr_matches_reference_definition():
    from oracle.vidtok_oracle import compute_psnr
    from vidtok_b200.dist import psnr_partial
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 4, 8, 8, generator=g) * 2 - 1
    y = torch.rand(2, 3, 4, 8, 8, generator=g) * 2.4 - 1.2
    p = psnr_partial(x, y)
    ref = compute_psnr((x.clamp(-1, 1) + 1) / 2, (y.clamp(-1, 1) + 1) / 2)
    assert abs(float(p[0] / p[1]) - float(ref)) < 1e-5
