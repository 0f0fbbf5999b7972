"""GPU, two PROCESSES: the batch-sharded sampling path on the real CUDA kernels.

Two ranks (gloo rendezvous on 127.0.0.1; both use cuda:0, the test box has one GPU) each sample their contiguous shard of a
7-sample global batch with the fused DiffWave trajectory — churn ON, so the in-kernel Philox noise must be keyed by the GLOBAL
sample index (`sample_offset`) — and gather; the result must be bit-identical to the same batch sampled by ONE process.
No collective touches the sampling path; the gather at the end is the only exchange.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
GB, L, STEPS, SEED = 7, 2048, 5, 31


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sample(noise, offset, precision):
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, KarrasSchedule, WaveNetNoise
    from oracle.weights import make_wavenet_state_dict
    dev = torch.device("cuda:0")
    net = WaveNetNoise(256, 3, 3, precision=precision)
    net.load_state_dict(make_wavenet_state_dict(256, 3, 5), strict=True)
    net = net.to(dev)
    diff = EluDiffusion(0.2)
    sig = KarrasSchedule(0.002, 80.0, 7.0, STEPS)()
    smp = EDMSampler(s_churn=40.0, s_noise=1.003, num_steps=STEPS)
    return smp(noise.to(dev), fn=diff.denoise_fn, net=net, sigmas=sig, churn_seed=SEED, sample_offset=offset).cpu()


def _worker(rank, world, port, precision, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audiodiffuser_b200.sharding import shard_noise, shard_range, gather_shards
    start, _ = shard_range(GB, rank, world)
    local = _sample(shard_noise(GB, rank, world, length=L, base_seed=SEED), start, precision)
    full = gather_shards(local, GB, rank, world)
    if rank == 0:
        torch.save(full, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_two_process_sharded_cuda_sampling_is_bit_identical(tmp_path, precision):
    from audiodiffuser_b200.sharding import shard_noise
    out = str(tmp_path / "full.pt")
    mp.spawn(_worker, args=(2, _free_port(), precision, out), nprocs=2, join=True)
    got = torch.load(out)
    want = _sample(shard_noise(GB, 0, 1, length=L, base_seed=SEED), 0, precision)
    assert got.shape == want.shape == (GB, 1, L)
    assert torch.equal(got, want)
