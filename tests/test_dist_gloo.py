"""The N > 1 path on CPU: world_size-2 gloo processes shard a batch of frames and all-gather
the per-frame parameter vectors (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from smplifyx_b200 import sharding
    r, w = sharding.init_from_env('gloo')
    assert (r, w) == (rank, world)
    mine = sharding.shard_range(n_frames, rank, world)
    # stand-in for the fitted parameters of this rank's frames: row f = f everywhere
    local = torch.tensor([[float(f)] * 122 for f in mine], dtype=torch.float32).reshape(len(mine), 122)
    full = sharding.gather_frames(local, n_frames)
    auto = sharding.gather_frames(local)
    q.put((rank, full.numpy(), auto.numpy()))
    sharding.finalize()


@pytest.mark.parametrize('n_frames', [8, 7])
def test_two_rank_gather(n_frames):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.repeat(np.arange(n_frames, dtype=np.float32)[:, None], 122, axis=1)
    for rank, full, auto in outs:
        assert np.array_equal(full, want)
        assert np.array_equal(auto, want)
