"""Waveform module + 16-bit WAV output: the caller and the data format on the far side of the path.

CPU tests: the PCM oracle against hand-computed known answers, the RIFF writer against the standard library's reader,
the per-rank test-generation plan (single process and two gloo ranks), config instantiation, deepcopy / pickle of the
backbone (EMA snapshots). GPU tests: the CUDA PCM kernel bit-exact against the oracle, the module's test-time generation
(files == PCM of the fused sampler's output, invariant to world size), training hooks.
"""
import os
import pickle
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- CPU ----------------------------------------------------------
def test_pcm16_oracle_known_answers():
    from oracle.wav import pcm16
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, -2.0, 1 / 32768, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768, -0.5 / 32768,
                  -1.5 / 32768, 0.99996948, 0.9999848], dtype=np.float32)
    want = [0, 16384, -16384, 32767, -32768, 32767, -32768, 1, 0, 2, 2, 0, -2, 32767, 32767]     # ties go to even
    assert pcm16(x).tolist() == want


def test_wav16_fixture_exact_rational_rule_and_stdlib_container(tmp_path):
    """tests/golden/wav16_fixture.*: ties on both sides of zero, +-1, beyond full scale, +-inf, NaN, denormals — integers from
    exact rational arithmetic, file assembled by the standard library's wave writer (oracle/make_wav_fixture.py). The oracle
    encoder must reproduce the integers, and this package's RIFF writer must reproduce the file byte for byte."""
    from audiodiffuser_b200.wav import write_wav16
    from oracle.wav import pcm16, read_wav16
    g = np.load(os.path.join(ROOT, "tests", "golden", "wav16_fixture.npz"))
    assert np.array_equal(pcm16(g["x"]), g["pcm"])
    frames = g["pcm"].size // 2
    stereo = g["pcm"][:2 * frames].reshape(2, frames)
    path = str(tmp_path / "mine.wav")
    write_wav16(path, torch.from_numpy(stereo.copy()), 16000)
    with open(path, "rb") as a, open(os.path.join(ROOT, "tests", "golden", "wav16_fixture.wav"), "rb") as b:
        assert a.read() == b.read()
    data, sr = read_wav16(os.path.join(ROOT, "tests", "golden", "wav16_fixture.wav"))
    assert sr == 16000 and np.array_equal(data, stereo)


def test_wav16_container_roundtrip(tmp_path):
    from audiodiffuser_b200.wav import wav16_header, write_wav16
    from oracle.wav import read_wav16
    rng = np.random.default_rng(0)
    for channels, frames in ((1, 0), (1, 1), (1, 16001), (2, 777)):
        pcm = torch.from_numpy(rng.integers(-32768, 32768, size=(channels, frames), dtype=np.int16))
        path = str(tmp_path / f"t_{channels}_{frames}.wav")
        write_wav16(path, pcm[0] if channels == 1 else pcm, 16000)
        data, sr = read_wav16(path)
        assert sr == 16000 and data.shape == (channels, frames) and np.array_equal(data, pcm.numpy())
        assert os.path.getsize(path) == 44 + 2 * channels * frames
    hdr = wav16_header(16000, 16000, 1)
    assert len(hdr) == 44 and hdr[:4] == b"RIFF" and hdr[8:16] == b"WAVEfmt " and hdr[36:40] == b"data"
    assert int.from_bytes(hdr[4:8], "little") == 36 + 32000 and int.from_bytes(hdr[40:44], "little") == 32000
    with pytest.raises(TypeError):
        write_wav16(str(tmp_path / "bad.wav"), torch.zeros(4), 16000)
    with pytest.raises(ValueError):
        wav16_header(-1, 16000)


def test_plan_test_shard_partitions_the_reference_index_space():
    from audiodiffuser_b200.waveform_module import plan_test_shard
    total, tb, classes = 37, 8, 10                                  # reference: 37 // 8 = 4 batches -> indices 0..31
    single = plan_test_shard(total, tb, 0, 1, classes)
    assert [i for idx, _, _ in single for i in idx] == list(range(32))
    assert single[0][2][:3] == ["test_0_0.wav", "test_1_1.wav", "test_2_2.wav"]
    assert single[1][1] == [j % classes for j in range(8)]          # position j of every batch gets class j % classes
    names_single = [n for _, _, names in single for n in names]
    for world in (2, 3, 5):
        names = [n for r in range(world) for _, _, ns in plan_test_shard(total, tb, r, world, classes) for n in ns]
        assert names == names_single                                # same files, each written exactly once
    assert all(l == 0 for _, labels, _ in plan_test_shard(16, 8, 0, 1, 1) for l in labels)
    assert plan_test_shard(5, 8, 0, 1, 10) == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _plan_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audiodiffuser_b200.waveform_module import _StandaloneModule, plan_test_shard
    m = _StandaloneModule()                                          # picks rank / world up from the process group
    plan = plan_test_shard(20, 4, m.trainer.global_rank, m.trainer.world_size, 3)
    with open(os.path.join(out_dir, f"plan{rank}.pkl"), "wb") as f:
        pickle.dump((m.trainer.is_global_zero, plan), f)
    dist.destroy_process_group()


def test_two_gloo_ranks_split_test_generation(tmp_path):
    from audiodiffuser_b200.waveform_module import plan_test_shard
    mp.spawn(_plan_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = [pickle.load(open(tmp_path / f"plan{r}.pkl", "rb")) for r in range(2)]
    assert [z for z, _ in got] == [True, False]
    merged = [b for _, plan in got for b in plan]
    assert [i for idx, _, _ in merged for i in idx] == list(range(20))
    single = plan_test_shard(20, 4, 0, 1, 3)
    assert [n for _, _, ns in merged for n in ns] == [n for _, _, ns in single for n in ns]


def test_module_config_instantiates_and_backbone_copies():
    import copy
    from audiodiffuser_b200.config import instantiate, load_yaml
    cfg = load_yaml(os.path.join(ROOT, "configs", "model", "diffwave.yaml"))
    cfg["net"].update(residual_channels=64, residual_layers=2, dilation_cycle=2)
    m = instantiate(cfg)
    assert type(m).__name__ == "DiffWaveformModule" and m.noise_scheduler.shape == (18,)
    assert abs(float(m.noise_scheduler[0]) - 80.0) < 1e-3 and m.total_test_samples == 2048
    opt = m.configure_optimizers()["optimizer"]
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["lr"] == 1e-4
    assert sum(p.numel() for g in opt.param_groups for p in g["params"]) == sum(p.numel() for p in m.net.parameters())
    twin = pickle.loads(pickle.dumps(copy.deepcopy(m.net).half()))   # what the EMA snapshot does with the backbone
    assert twin._handle is None and next(twin.parameters()).dtype == torch.float16
    assert all(torch.equal(a.half(), b) for a, b in zip(m.net.state_dict().values(), twin.state_dict().values()))
    exp = load_yaml(os.path.join(ROOT, "configs", "experiment", "sc09", "diffwave_sc09.yaml"))
    assert {"override /model": "diffwave.yaml"} in exp["defaults"]


# ---------------------------------------------------------------- GPU ----------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.gpu
def test_pcm16_kernel_matches_fixture():
    """The CUDA encoder on the committed fixture vector (ties, saturation, +-inf, NaN): bit-exact."""
    from audiodiffuser_b200.wav import pcm16_encode
    g = np.load(os.path.join(ROOT, "tests", "golden", "wav16_fixture.npz"))
    got = pcm16_encode(torch.from_numpy(g["x"]).cuda()).cpu().numpy()
    assert np.array_equal(got, g["pcm"])


@pytest.mark.gpu
def test_pcm16_kernel_bit_exact():
    from audiodiffuser_b200.wav import pcm16_encode
    from oracle.wav import pcm16
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for n in (1, 7, 8, 9, 4099, 16000 * 33 + 5):
        x = (torch.randn(n, generator=g) * 0.6)
        x[::5] = torch.round(x[::5] * 32768 * 2) / (32768 * 2)      # many exact .5 ties
        x[:1] = 1.0
        got = pcm16_encode(x.to(dev)).cpu().numpy()
        assert np.array_equal(got, pcm16(x.numpy())), n
    x = torch.randn(3, 1, 1001, generator=g).to(dev)
    assert np.array_equal(pcm16_encode(x[:, :, 1:]).cpu().numpy(), pcm16(x[:, :, 1:].cpu().numpy()))   # unaligned view
    assert pcm16_encode(torch.tensor([float("nan"), float("inf"), -float("inf")], device=dev)).tolist() == [0, 32767, -32768]
    with pytest.raises(Exception):
        pcm16_encode(torch.zeros(4))


def _small_module(tmp_path, total=6, use_ema=True, precision="fp32", L=2000):
    from audiodiffuser_b200.config import instantiate, load_yaml
    torch.manual_seed(0)
    cfg = load_yaml(os.path.join(ROOT, "configs", "model", "diffwave.yaml"))
    cfg["net"].update(residual_channels=64, residual_layers=4, dilation_cycle=2, precision=precision)
    cfg["sampler"]["num_steps"] = cfg["noise_scheduler"]["num_steps"] = 4
    cfg.update(generated_length=L, total_test_samples=total, use_ema=use_ema, num_ema_snapshot_item=4, audio_sample_rate=1600)
    m = instantiate(cfg)
    with torch.no_grad():
        m.net.output_projection.conv.weight.normal_(0, 1 / 8)       # the zero-initialised output conv would give F = 0
    m = m.cuda()
    m.logger.save_dir = str(tmp_path)
    return m


@pytest.mark.gpu
def test_module_test_generation_matches_sampler_and_is_world_size_invariant(tmp_path):
    from audiodiffuser_b200.sharding import noise_for_indices
    from oracle.wav import pcm16, read_wav16
    m = _small_module(tmp_path / "w1")
    files = m.test(batch_size=2)
    assert [os.path.basename(f) for f in files] == [f"test_{g % 2}_{g}.wav" for g in range(6)]
    noise = noise_for_indices(range(6), 2000, 0).cuda()
    want = m.sampler(noise, fn=m.diffusion.denoise_fn, net=m.net, sigmas=m.noise_scheduler.cuda())
    for g, f in enumerate(files):
        data, sr = read_wav16(f)
        assert sr == 1600 and data.shape == (1, 1600)                # audio_dur = 1 s of the 2000 generated frames
        ref = pcm16(want[g, :, :1600].cpu().numpy())
        assert np.abs(data.astype(np.int32) - ref).max() <= 1        # batch composition may move a value across a tie
        assert (data != ref).mean() < 1e-3
    # the same six files from a "world" of 2: rank 0 writes 0..2, rank 1 writes 3..5
    got = {}
    for rank in range(2):
        mr = _small_module(tmp_path / f"w2r{rank}")
        mr.trainer.global_rank, mr.trainer.world_size, mr.trainer.is_global_zero = rank, 2, rank == 0
        for f in mr.test(batch_size=2):
            got[os.path.basename(f)] = read_wav16(f)[0]
    assert sorted(got) == sorted(os.path.basename(f) for f in files)
    for f in files:
        a, b = read_wav16(f)[0].astype(np.int32), got[os.path.basename(f)].astype(np.int32)
        assert np.abs(a - b).max() <= 1 and (a != b).mean() < 1e-3


@pytest.mark.gpu
def test_module_training_hooks_ema_snapshot_and_reload(tmp_path):
    m = _small_module(tmp_path, precision="fp32", L=2048)
    g = torch.Generator().manual_seed(1)
    data = [{"audio": (torch.rand(4, 2048, generator=g) * 2 - 1).cuda(), "label": torch.zeros(4, dtype=torch.long).cuda()}
            for _ in range(4)]
    m.optimizer.keywords["lr"] = 5e-4
    losses = m.fit_steps(data * 5)
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])
    assert m.cur_nitem == 80 and m.global_step == 20 and "train/loss" in m.logged
    snaps = sorted(os.listdir(tmp_path / "ema_snapshots"))
    assert snaps and all(s.startswith("ema_prof_") for s in snaps)   # every 4 items, from global_step 1 on
    m.validation_step(data[0], 0)
    m.on_validation_epoch_end()
    assert len(os.listdir(tmp_path / "val_audio")) == 1 and np.isfinite(m.val_loss_best)
    # test-time generation from an fp16 EMA snapshot, like `ema_ckpt_path` in the reference
    m.ema_ckpt_path = str(tmp_path / "ema_snapshots" / snaps[-1])
    files = m.test(batch_size=2)
    assert len(files) == 6 and next(m.net.parameters()).dtype == torch.float32 and next(m.net.parameters()).is_cuda
