"""world_size-2 gloo test of the clip-parallel plumbing (ccedit_b200/parallel.py): clip sharding covers every clip once,
rank 0's weights reach rank 1 through the bucketed broadcast, timing is reduced with MAX, results gather in clip order."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from ccedit_b200 import parallel
    from ccedit_b200.modules import ResBlock3D, SpatialTransformer3D
    r, _, w = parallel.init_from_env("gloo")
    torch.manual_seed(1234 + rank)                     # different init per rank: broadcast must make them equal
    net = torch.nn.ModuleList([ResBlock3D(64, 128, 96), SpatialTransformer3D(64, 8, 8, 768)])
    before = torch.cat([p.detach().flatten() for p in net.parameters()]).clone()
    nbytes = parallel.broadcast_weights(net, src=0, bucket_bytes=64 << 10)     # small buckets: several broadcasts
    after = torch.cat([p.detach().flatten() for p in net.parameters()])
    clips = parallel.shard_clips(7, r, w)
    tmax = parallel.max_over_ranks(10.0 + rank, "cpu")
    gathered = parallel.gather_results([f"clip{i}" for i in clips], w)
    parallel.barrier()
    q.put((rank, float(before.double().sum()), float(after.double().sum()), nbytes, clips, tmax, gathered))
    torch.distributed.destroy_process_group()


def test_two_rank_clip_parallel_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, a0, n0, c0, t0, g0), (r1, b1, a1, n1, c1, t1, g1) = res
    assert b0 != b1                                   # ranks started from different weights
    assert a0 == a1 == b0                             # ... and ended with rank 0's
    assert n0 == n1 > 0
    assert sorted(c0 + c1) == list(range(7)) and c0 == [0, 2, 4, 6] and c1 == [1, 3, 5]
    assert t0 == t1 == 11.0                           # max over ranks
    assert g0 == g1 == [f"clip{i}" for i in range(7)]


def test_single_process_is_a_no_op():
    from ccedit_b200 import parallel
    assert parallel.shard_clips(5, 0, 1) == [0, 1, 2, 3, 4]
    assert parallel.broadcast_weights(torch.nn.Linear(2, 2)) == 0
    assert parallel.max_over_ranks(3.5, "cpu") == 3.5
