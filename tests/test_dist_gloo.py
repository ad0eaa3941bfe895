"""world_size-2 gloo test of the N>1 host path: sharding + the single filter broadcast."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    import imagemagick_b200 as im
    from imagemagick_b200 import dist as mdist
    dist.init_process_group("gloo")
    job = None
    if rank == 0:
        taps = im.AcquireKernelInfo("blur:0x4").arrays()[0][0].ravel()
        job = mdist.FilterJob(0.0, 4.0, 2048, 2048, 22, taps)
    job = mdist.broadcast_job(job)
    mine = mdist.shard_indices(7, rank, world)
    slow = mdist.max_over_ranks(float(rank + 1))
    k = mdist.blur_kernel_from_taps(job.taps).arrays()
    mdist.barrier()
    q.put((rank, job.taps.tobytes(), job.out_columns, job.resize_filter, mine, slow,
           [a[0].shape for a in k], k[0][0].tobytes()))
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (r0, taps0, oc0, f0, idx0, slow0, shapes0, kb0), (r1, taps1, oc1, f1, idx1, slow1, shapes1, kb1) = res
    assert taps0 == taps1 and len(taps0) == 33 * 8 and kb0 == kb1 == taps0      # bit-identical weights everywhere
    assert (oc0, f0) == (oc1, f1) == (2048, 22)
    assert idx0 == [0, 2, 4, 6] and idx1 == [1, 3, 5]                            # ragged batch, no overlap
    assert slow0 == slow1 == 2.0
    assert shapes0 == [(1, 33), (33, 1)]


def test_shard_indices_cover_batch():
    from imagemagick_b200.dist import shard_indices
    for n in (0, 1, 5, 256):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in shard_indices(n, r, w))
            assert got == list(range(n))
