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


def _dp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audiodiffuser_b200.training import allreduce_sum_
    from audiodiffuser_b200.sharding import shard_range
    g = torch.Generator().manual_seed(3)
    per_sample_grads = torch.randn(8, 1000, generator=g)            # gradient of each of the 8 global samples
    a, b = shard_range(8, rank, world)
    local_mean_grad = per_sample_grads[a:b].mean(dim=0)            # what one rank's backward of loss.mean() produces
    total = allreduce_sum_(local_mean_grad.clone(), world) * (1.0 / world)     # FusedTrainer: sum, then grad_scale = 1 / world
    if rank == 0:
        torch.save(total, out)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_global_mean(tmp_path):
    """The training step's only collective: sum of the flat gradients, scaled by 1 / world inside the optimizer
    kernel, equals the gradient of the mean loss over the global batch (equal shards)."""
    out = str(tmp_path / "grad.pt")
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g = torch.Generator().manual_seed(3)
    want = torch.randn(8, 1000, generator=g).mean(dim=0)
    assert torch.allclose(torch.load(out), want, atol=1e-6)
