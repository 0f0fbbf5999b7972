"""N > 1 host path on CPU: two gloo ranks shard a batch, 'sample' it with a pure function of the
noise, and gather — the result must equal the single-process result (no collective on the data path,
one optional gather at the end)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, gb, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audiodiffuser_b200.sharding import shard_noise, gather_shards
    noise = shard_noise(gb, rank, world, length=50, base_seed=11)
    local = torch.tanh(noise * 0.5) + 1.0                     # stands in for the per-sample trajectory
    full = gather_shards(local, gb, rank, world)
    if rank == 0:
        torch.save(full, out)
    dist.destroy_process_group()


def test_two_rank_sharded_sampling_matches_single_process(tmp_path):
    from audiodiffuser_b200.sharding import shard_noise
    gb = 7                                                     # ragged: shards of 4 and 3
    out = str(tmp_path / "full.pt")
    mp.spawn(_worker, args=(2, _free_port(), gb, out), nprocs=2, join=True)
    got = torch.load(out)
    want = torch.tanh(shard_noise(gb, 0, 1, length=50, base_seed=11) * 0.5) + 1.0
    assert torch.equal(got, want)
