"""world_size-2 (gloo, CPU) coverage of the N>1 path: pair-block shard plan + the one all-reduce of [J; K].

Each rank computes the partial J/K of the rows of ITS tile-range shard with the oracle's C port (standing in for the GPU kernel,
which needs a device), the product's ``reduce_partials`` sums them, and the result must equal the unsharded build.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n, seed, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cjk, synth
    from sqcc_b200.backend import ShardPlan, reduce_partials
    plan = ShardPlan(n, world, rank)
    d = synth.synth_density_uhf(n, 5, 3, 11)
    j, k = cjk.jk_packed_segments(n, seed, plan.segments(), d)
    buf = torch.from_numpy(np.concatenate([j[None], k]))
    reduce_partials(buf)
    if rank == 0:
        np.save(os.path.join(out_dir, "jk.npy"), buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_partials_allreduce(tmp_path, world):
    from oracle import cjk, synth
    n, seed = 23, 4711
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "jk.npy")
    d = synth.synth_density_uhf(n, 5, 3, 11)
    j, k = cjk.jk_packed(cjk.synth_canonical(n, seed), n, d)
    assert np.abs(got[0] - j).max() < 1e-12 and np.abs(got[1:] - k).max() < 1e-12
