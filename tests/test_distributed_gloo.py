"""world_size-2 gloo test of the multi-GPU host logic (sharding + scalar statistics reduction)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from codeps_b200.distributed import global_mean_from_shards, reduce_loss_dict, shard_bounds


def test_shards_partition_the_batch():
    for batch in (1, 7, 16, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, per_sample):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end = shard_bounds(per_sample.shape[0], rank, world)
        local = per_sample[begin:end]
        # each rank's loss is the mean over its shard (what the kernels return per rank)
        got = global_mean_from_shards(local[:, 0].mean(), end - begin)
        want = per_sample[:, 0].mean()
        assert torch.allclose(got, want, rtol=1e-6), (rank, got, want)
        both = reduce_loss_dict({"recon": local[:, 0].mean(), "smth": local[:, 1].mean()}, end - begin)
        assert torch.allclose(both["recon"], per_sample[:, 0].mean(), rtol=1e-6)
        assert torch.allclose(both["smth"], per_sample[:, 1].mean(), rtol=1e-6)
    finally:
        dist.destroy_process_group()


def test_global_statistics_over_two_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    per_sample = torch.rand(7, 2, generator=torch.Generator().manual_seed(3))  # uneven shards: 4 + 3
    mp.spawn(_worker, args=(2, port, per_sample), nprocs=2, join=True)


def test_single_process_reduction_is_identity():
    out = reduce_loss_dict({"a": torch.tensor(0.25), "b": torch.tensor(2.0)}, 4)
    assert float(out["a"]) == 0.25 and float(out["b"]) == 2.0
    assert float(global_mean_from_shards(torch.tensor(0.5), 3)) == 0.5
