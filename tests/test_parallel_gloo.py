"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic: batch sharding of the op's
inputs and the flat-bucket gradient all-reduce.  The op itself has no CPU path, so the oracle plays
the op here — what is under test is that sharded evaluation + all-reduce reproduces the
single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from simpb_b200 import parallel, synthetic


def test_shard_range_partitions_every_batch_exactly():
    for bs in range(0, 19):
        for world in (1, 2, 3, 4, 8):
            owned = []
            for r in range(world):
                lo, hi = parallel.shard_range(bs, world, r)
                assert 0 <= lo <= hi <= bs
                owned += list(range(lo, hi))
            assert owned == list(range(bs))
            sizes = [parallel.shard_range(bs, world, r) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, alias=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                # same replicated "module" on every rank
        lin = torch.nn.Linear(16, 24)
        d = synthetic.op_inputs_uniform(bs=5, A=6, P=3, K=2, levels=((4, 6), (2, 3)), C=16, G=4, seed=7)
        mine = parallel.shard_batch(d, world, rank)
        lo, hi = parallel.shard_range(5, world, rank)
        assert mine["mc_ms_feat"].shape[0] == hi - lo
        assert mine["spatial_shape"] is d["spatial_shape"]
        # forward of the op on the shard (oracle as the op), then a replicated layer on top
        out = torch.from_numpy(oracle.forward(mine["mc_ms_feat"], mine["spatial_shape"],
                                              mine["scale_start_index"], mine["sampling_location"],
                                              mine["weights"])).float()
        bucket = None
        if alias:   # DDP style: the gradients are views of the flat bucket before the backward runs
            bucket = parallel.GradBucket(lin.parameters(), alias_grads=True)
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
        y = lin(out)
        # loss = sum over the GLOBAL batch / global batch size: each rank contributes its share
        loss = y.square().sum() / 5.0
        loss.backward()
        if alias:
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
        else:
            bucket = parallel.GradBucket(lin.parameters())
        # DDP semantics: mean over ranks of per-rank gradients; scale so the sum is what we want
        for p in lin.parameters():
            p.grad.mul_(world)
        bucket.all_reduce_mean()
        bucket.wait()
        gathered = [torch.zeros(5, 6, 16) for _ in range(world)] if rank == 0 else None
        pad = torch.zeros(5, 6, 16)
        pad[lo:hi] = out
        dist.reduce(pad, 0)
        if rank == 0:
            np.savez(os.path.join(out_dir, "rank0.npz"), out=pad.numpy(),
                     gw=lin.weight.grad.numpy(), gb=lin.bias.grad.numpy())
        # every rank must hold the same averaged gradients
        chk = [torch.zeros_like(lin.weight.grad) for _ in range(world)]
        dist.all_gather(chk, lin.weight.grad)
        assert all(torch.equal(c, chk[0]) for c in chk)
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("alias", [False, True])
def test_sharded_op_plus_bucket_allreduce_equals_single_process(tmp_path, alias):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), alias), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    torch.manual_seed(0)
    lin = torch.nn.Linear(16, 24)
    d = synthetic.op_inputs_uniform(bs=5, A=6, P=3, K=2, levels=((4, 6), (2, 3)), C=16, G=4, seed=7)
    out = torch.from_numpy(oracle.forward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                                          d["sampling_location"], d["weights"])).float()
    (lin(out).square().sum() / 5.0).backward()
    np.testing.assert_allclose(got["out"], out.numpy(), rtol=0, atol=0)       # shards are independent
    np.testing.assert_allclose(got["gw"], lin.weight.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got["gb"], lin.bias.grad.numpy(), rtol=1e-5, atol=1e-6)
