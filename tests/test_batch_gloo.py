"""CPU, world_size 2 over gloo: the host logic of the multi-GPU batch path (sharding, gather
order, max-over-ranks), with the CPU oracle standing in for the per-pair compute."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mccnn_b200  # noqa: F401
    from mccnn_b200 import batch, pipeline, synth
    from oracle import oracle as orc

    opt = pipeline.make_params("kitti", "fast")
    op = orc.Params(**opt.as_dict())
    H, W, C, D = 12, 30, 4, 6
    done = []

    def compute(i):
        p = synth.make_pair(H, W, C, D, seed=100 + i)
        done.append(i)
        return torch.from_numpy(orc.stereo_predict(p["featL"], p["featR"], p["imgL"], p["imgR"], D, op))

    res = batch.run_sharded(n_items, compute)
    assert done == batch.shard_indices(n_items, rank, world)
    t = batch.max_over_ranks(1.0 + rank)
    assert t == float(world)
    if rank == 0:
        assert len(res) == n_items and all(r is not None for r in res)
        np.save(os.path.join(outdir, "gathered.npy"), torch.stack(res).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 4, 1])      # 1: a rank that owns nothing still takes part in the gather
def test_sharded_batch_world2(tmp_path, n_items):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    sys.path.insert(0, ROOT)
    import mccnn_b200  # noqa: F401
    from mccnn_b200 import pipeline, synth
    from oracle import oracle as orc

    opt = pipeline.make_params("kitti", "fast")
    op = orc.Params(**opt.as_dict())
    for i in range(n_items):
        p = synth.make_pair(12, 30, 4, 6, seed=100 + i)
        want = orc.stereo_predict(p["featL"], p["featR"], p["imgL"], p["imgR"], 6, op)
        assert np.array_equal(got[i], want, equal_nan=True), "pair %d came back wrong / out of order" % i


def test_shard_indices_cover_exactly_once():
    from mccnn_b200 import batch

    for n in (0, 1, 7, 194):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in batch.shard_indices(n, r, world))
            assert seen == list(range(n))
            for i in range(n):
                assert i in batch.shard_indices(n, batch.owner_of(i, world), world)
    with pytest.raises(ValueError):
        batch.shard_indices(4, 2, 2)
