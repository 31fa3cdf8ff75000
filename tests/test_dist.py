"""CPU, world_size 2 over gloo: the sharding + all-gather logic of
miniframe_b200/dist.py, with the oracle standing in for the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from miniframe_b200 import dist as mfd


def test_shard_bounds_cover_batch():
    for total in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [mfd.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = mfd.shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import gp_oracle as go
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(12, seed=0)
    y = go.biggp_y(rv, rhk, bis)
    A = go.synthetic_biggp_hypers(total, seed=0)

    def compute(gp, a, bm, c):
        return torch.tensor([go.biggp_loglike_literal("QuasiPeriodic", row, t, y, e1, e2, e3)
                             for row in a], dtype=torch.float64)

    full = mfd.log_likelihood_batch_sharded(None, A, compute=compute)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 1])
def test_sharded_batch_two_ranks_gloo(tmp_path, total):
    from oracle import gp_oracle as go
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    t, (rv, rhk, bis), (e1, e2, e3) = go.synthetic_series(12, seed=0)
    y = go.biggp_y(rv, rhk, bis)
    want = np.array([go.biggp_loglike_literal("QuasiPeriodic", row, t, y, e1, e2, e3)
                     for row in go.synthetic_biggp_hypers(total, seed=0)])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == (total,)
        assert np.array_equal(got, want)
