#!/usr/bin/env python
"""Benchmark of the EDM sampling hot path (BASELINE.json metric: DiffWave SC09 EDM-Heun samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: the full 18-step EDM-Heun trajectory (35
network evaluations of DiffWave C=256 / 36 layers / L=16000) for `--batch` samples per GPU
(default 256 = BASELINE.json configs[1]), synthetic N(0,1) noise, seeded random weights (the
zero-initialised output layer re-randomised), bf16 tensor-core path.

Prints ONE JSON line (rank 0). `value` = whole-job samples/s with inputs resident in HBM; `e2e` =
the same metric through the public sampler API with pinned HOST buffers (H2D of the noise and D2H
of the waveforms inside the timed region). Multi-GPU (torchrun, one process per GPU): the batch is
sharded, every rank samples its own `--batch` waveforms with no collective on the data path
(weak scaling); time = max over ranks.

`--impl reference` times the reference algorithm on the host CPU (the oracle port in oracle/, which
tests/test_oracle_golden.py pins bit-for-bit to the reference's own fp32 outputs) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "diffwave_sc09_edm_heun18_samples_per_sec"
UNIT = "samples/s"
C, LAYERS, CYCLE, L, STEPS_EDM, SIGMA_DATA = 256, 36, 12, 16000, 18, 0.2
BLOCK_KERNEL = int(os.environ.get("ADB_BLOCK_KERNEL", "3"))       # 3 = z-stash kernel + skip GEMM (default), see adb200.cu
NFE = 2 * STEPS_EDM - 1
# algorithmic conv FLOPs per sample per network evaluation; the last block's unused
# residual half (2.097 G) is not computed and not counted
FLOP_G1 = 2 * 512 * 768 * L                                        # dilated conv: 12.583 G per block
FLOP_G2_HALF = 2 * 256 * 256 * L                                   # residual OR skip half of the 1x1 conv: 2.097 G
FLOP_BLOCK = FLOP_G1 + 2 * FLOP_G2_HALF                            # 16.777 G per residual block
FLOP_EVAL_BLOCKS = LAYERS * FLOP_BLOCK - FLOP_G2_HALF              # everything the residual stack computes
# z-stash path: the block launches do G1 + the residual half, one skip GEMM per evaluation does all 36 skip halves
FLOP_EVAL_ZS_BLOCKS = LAYERS * FLOP_G1 + (LAYERS - 1) * FLOP_G2_HALF
FLOP_EVAL_SKIP_GEMM = LAYERS * FLOP_G2_HALF
# fused sampler step: fp32 state, fp32 net output; mid kernel r(x,F) w(d,x') = 16 B, post kernel
# r(x,d,F) w(x) = 16 B per state element
STEP_BYTES_PER_ELEM = 16.0
WORKLOAD = (f"DiffWave C={C} layers={LAYERS} cycle={CYCLE}, SC09 shape 1x{L}, EDM Heun {STEPS_EDM} steps ({NFE} network "
            f"evaluations), sigma_data={SIGMA_DATA}, Karras(0.002,80,rho=7), s_churn=0 (BASELINE.json configs[1])")


class quiet_stdout:
    """Route the process-level stdout (fd 1) to stderr while NCCL initialises: its version banner goes to stdout on some boxes and
    the contract is ONE JSON line there."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def init_nccl(dev):
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")            # keep stdout to the JSON line(s)
    with quiet_stdout():
        dist.init_process_group("nccl", device_id=dev)
        t = torch.zeros(1, device=dev)
        dist.all_reduce(t)                                             # communicator (and its banner) created here
        torch.cuda.synchronize(dev)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (bf16 sustained, HBM copy)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:                                          # noqa: BLE001
                pass
            self._halt.wait(0.25)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        busy = sorted(self.samples)[len(self.samples) // 4:] if self.samples else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def step_kernel_hbm(dev, elements=64 * 1024 * 1024, reps=6):
    """HBM roofline of the fused Heun step kernels: mid (r x,F / w d,x1) + post (r x,d,F / w x) = 32 B per element
    per pair, on a state far larger than L2, through the C ABI on torch's current stream."""
    from audiodiffuser_b200 import _native as N
    lib, st = N.lib(), N.stream_ptr(dev)
    bufs = [torch.randn(elements, device=dev) for _ in range(2)] + [torch.empty(elements, device=dev) for _ in range(3)]
    x, f, d, x1, out = bufs

    def pair():
        N.check(lib.adb_edm_heun_mid(N.ptr(x), N.ptr(f), 3.0, SIGMA_DATA, -1.3, N.ptr(d), N.ptr(x1), elements, st))
        N.check(lib.adb_edm_heun_post(N.ptr(x), N.ptr(d), N.ptr(f), 1.7, SIGMA_DATA, -1.3, N.ptr(out), elements, st))

    for _ in range(3):
        pair()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        pair()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    nbytes = 32.0 * elements
    return {"gbs": nbytes * reps / (ms * 1e-3) / 1e9, "bytes_per_launch": nbytes / 2, "launches": 2 * reps,
            "avg_ms": ms / (2 * reps), "elements": elements}


# BASELINE.json configs[0]: the reference's CPU-runnable case is batch 8 (ADB_BENCH_CPU_BATCH shrinks the sample for the
# contract test of this arm, tests/test_bench_contract.py)
CPU_BATCH = int(os.environ.get("ADB_BENCH_CPU_BATCH", "8"))


def cpu_reference_denoiser():
    """(kind, denoise(x, sigma)) of the CPU arm: the reference's OWN classes (WaveNetNoise + EluDiffusion.denoise_fn, imported
    unmodified from baseline/_ref — see baseline/install_ref.py — or /root/reference) when they are importable, else the
    oracle port (oracle/, pinned bit-exact to the reference's fp32 outputs by tests/test_oracle_golden.py)."""
    from oracle.weights import make_wavenet_state_dict
    sd = make_wavenet_state_dict(C, LAYERS, seed=0)
    try:
        from oracle import ref_loader
        if ref_loader.reference_available():
            ref = ref_loader.import_reference()
            net = ref.wavenet.WaveNetNoise(residual_channels=C, residual_layers=LAYERS, dilation_cycle=CYCLE)
            net.load_state_dict(sd, strict=True)
            net.eval()
            diff = ref.diffusion.EluDiffusion(sigma_data=SIGMA_DATA)
            adapter = ref_loader.WaveNetAdapter(net)
            return "reference", lambda x, s: diff.denoise_fn(x, net=adapter, sigma=s, inference=True)
    except Exception as e:                                             # noqa: BLE001
        print(f"[bench] reference classes not usable ({e!r}); timing the oracle port", file=sys.stderr)
    from oracle import edm as oedm, wavenet as owav
    net_fn = owav.make_net_fn(sd, CYCLE)
    return "port", lambda x, s: oedm.denoise(x, net_fn, SIGMA_DATA, sigma=s)


def cpu_reference_eval_time(n_evals, threads):
    """Time `n_evals` full-size denoiser calls (B = CPU_BATCH) of the reference on the host cores. Returns (kind, times)."""
    torch.set_num_threads(threads)
    kind, den = cpu_reference_denoiser()
    x = torch.randn(CPU_BATCH, 1, L, generator=torch.Generator().manual_seed(1))
    times = []
    with torch.no_grad():
        for s in ([80.0, 1.0, 0.05] * ((n_evals + 2) // 3))[:n_evals]:
            t0 = time.perf_counter()
            den(x * s, s)
            times.append(time.perf_counter() - t0)
    return kind, times


CPU_KIND_TEXT = {"reference": "the reference's own WaveNetNoise + EluDiffusion.denoise_fn (unmodified, imported from baseline/_ref)",
                 "port": "oracle port of the reference algorithm (pinned bit-exact to the reference's fp32 outputs)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    evals_per_step = 2
    for _ in range(args.warmup):
        cpu_reference_eval_time(1, threads)
    kind, t = cpu_reference_eval_time(evals_per_step * max(args.steps, 1), threads)
    t_eval = sum(t) / len(t)
    value = CPU_BATCH / (NFE * t_eval)                 # samples/s: one sample needs NFE evaluations
    sample = (f"B={CPU_BATCH} x {len(t)} full-size denoiser calls (of the {NFE} one trajectory needs), fp32, "
              f"extrapolated x{NFE}; {CPU_KIND_TEXT[kind]}; torch {torch.__version__} CPU, {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * evals_per_step * t_eval, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus,
                       "parallelism": "host CPU threads", "arm": f"reference on the host CPU: {CPU_KIND_TEXT[kind]}; bounded sample, "
                       "see cpu_baseline.sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


UNET_CFG4 = dict(   # SURVEY.md §8(d) config 4 (BASELINE.json configs[3]): audio-diffusion-pytorch defaults
    channels=128, cond_drop_prob=0.0, class_cond=False, text_cond=False, num_filters=128, window_length=32, stride=16,
    in_channels=2, resnet_groups=8, kernel_multiplier_downsample=2, multipliers=[1, 2, 4, 4, 4, 4, 4],
    factors=[4, 4, 4, 2, 2, 2], num_blocks=[2, 2, 2, 2, 2, 2], attentions=[False, False, False, True, True, True],
    attention_heads=8, attention_multiplier=2, use_nearest_upsample=False, use_skip_scale=True,
    use_attention_bottleneck=True)
UNET_L, UNET_STEPS, UNET_FLOP_EVAL = 262144, 50, 61.69e9


def run_unet(args, rank, world, local_rank):
    """Secondary workload (BASELINE.json configs[3]): UNet1d with attention, 2 x 262144 stereo 48 kHz, EDM Heun 50 steps
    (99 network evaluations per waveform). Same JSON contract; no roofline block (time is spread over ~400 launches per
    evaluation, see profiles/)."""
    import torch.distributed as dist
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, KarrasSchedule, UNet1dBase, _native
    from audiodiffuser_b200.sharding import shard_noise
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)
    B = args.batch or 128
    torch.manual_seed(0)
    net = UNet1dBase(precision=args.precision, **UNET_CFG4)
    net.unet.to_out.to_out.weight.data.uniform_(-0.06, 0.06)          # zero-initialised in the reference (unet1d.py:619)
    net = net.to(dev)
    diff = EluDiffusion(sigma_data=SIGMA_DATA)
    sampler = EDMSampler(s_tmin=0, s_tmax=float("inf"), s_churn=0.0, s_noise=1.0, num_steps=UNET_STEPS, cond_scale=1.0, use_heun=True)
    sigmas = KarrasSchedule(0.002, 80.0, 7.0, UNET_STEPS)().to(dev)
    noise_host = shard_noise(B * world, rank, world, UNET_L, base_seed=4321, channels=2).pin_memory()
    noise_dev = noise_host.to(dev)
    out_host = torch.empty_like(noise_host).pin_memory()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        return sampler(noise_dev, fn=diff.denoise_fn, net=net, sigmas=sigmas)

    def step_e2e():
        y = sampler(noise_host.to(dev, non_blocking=True), fn=diff.denoise_fn, net=net, sigmas=sigmas)
        out_host.copy_(y, non_blocking=True)
        return y

    for _ in range(max(args.warmup, 1)):
        y = step_resident()
    _native.check_async()
    assert torch.isfinite(y).all() and float(y.abs().max()) > 0

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    l0 = _native.lib().adb_launch_count(0) + net.graph_launches
    ms_res = timed(step_resident, args.steps)
    launches = _native.lib().adb_launch_count(0) + net.graph_launches - l0
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if clocks else None
    _native.check_async()
    if rank == 0:
        nfe = 2 * UNET_STEPS - 1
        total = B * world * args.steps
        line = {"metric": "unet1d_stereo48k_edm_heun50_samples_per_sec", "value": total / (ms_res * 1e-3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms_res / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic",
                "config": {"workload": f"UNet1d 102M (channels 128, multipliers 1-2-4-4-4-4-4, attention at the 3 deepest levels), "
                                       f"2x{UNET_L} stereo 48 kHz, EDM Heun {UNET_STEPS} steps ({nfe} network evaluations), "
                                       f"BASELINE.json configs[3]", "batch_per_gpu": B, "global_batch": B * world,
                           "parallelism": f"batch-sharded x{world}, no collective",
                           "l2": "CUDA-graph replay of ~400 launches per evaluation; activations of the upper levels exceed L2"},
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": noise_host.numel() * 4,
                        "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / args.steps if ms_e2e else None},
                "gpu_launches": int(launches), "effective_tflops": UNET_FLOP_EVAL * nfe * total / (ms_res * 1e-3) / 1e12,
                "ms_per_network_evaluation": ms_res / args.steps / nfe, "clocks": clk}
        peaks = load_peaks()
        per_gpu = line["effective_tflops"] / world
        line["roofline"] = {"bound": "tensor", "kernel": "whole network evaluation (at B=128: cl_conv3_gn_tc_kernel, the convolutions with GroupNorm apply + SiLU in their operand path, 43 %; "
                                                         "cl_conv_tc_kernel 32 %; GroupNorm statistics 12 %; profiles/r2_launches_unet1d_b128.csv)",
                            "achieved": per_gpu, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s (per GPU)",
                            "frac": per_gpu / peaks["bf16_tflops"], "traffic": None, "peak_source": peaks["source"]}
        if not args.no_cpu_baseline and world == 1:       # host baseline: rank 0 at N = 1 only
            from oracle import unet1d as ounet
            from oracle.weights import make_unet1d_state_dict
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            sd_cpu = make_unet1d_state_dict(UNET_CFG4, 0)
            xc, tc = torch.randn(1, 2, UNET_L), torch.zeros(1)
            ts = []
            with torch.no_grad():
                for _ in range(3):
                    t0 = time.perf_counter()
                    ounet.unet1d_forward(sd_cpu, UNET_CFG4, xc, tc)
                    ts.append(time.perf_counter() - t0)
            t_eval = sum(ts[1:]) / 2
            line["cpu_baseline"] = {"value": 1.0 / (nfe * t_eval), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"B=1 x 2 full-size network evaluations after 1 warm-up (of {nfe} per waveform), fp32 torch-CPU "
                                              f"oracle port, extrapolated x{nfe}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args, rank, world, local_rank):
    """BASELINE.json configs[2]: DiffWave SC09-shape EDM training step (DSM loss, LogNormal(-1.2, 1.2) sigmas, AdamW
    lr 1e-4) with one NCCL all-reduce of the flat gradient per step. --precision bf16: tcgen05 forward / dgrad / wgrad
    GEMMs (bf16 operands, fp32 accumulate and fp32 master weights); --precision fp32: CUDA-core kernels."""
    import torch.distributed as dist
    from audiodiffuser_b200 import EluDiffusion, WaveNetNoise, _native
    from audiodiffuser_b200.training import FusedTrainer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)
    B = args.batch or (32 if args.precision == "bf16" else 4)   # 32 = data.batch_size of the reference experiment (diffunet_complex_sc09.yaml:67)
    torch.manual_seed(0)                                   # identical initial weights on every rank, like DDP's broadcast
    net = WaveNetNoise(C, LAYERS, CYCLE, precision=args.precision)
    net.output_projection.conv.weight.data.normal_(0.0, 1.0 / 16.0)
    net = net.to(dev)
    trainer = FusedTrainer(net, EluDiffusion(sigma_data=SIGMA_DATA), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.01)
    g = torch.Generator().manual_seed(100 + rank)
    x_host = (torch.randn(B, 1, L, generator=g) * SIGMA_DATA).clamp(-1, 1).pin_memory()
    sig_host = (torch.randn(B, generator=g) * 1.2 - 1.2).exp().pin_memory()
    x_dev, sig_dev = x_host.to(dev), sig_host.to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        return trainer.step(x_dev, sig_dev)

    def step_e2e():
        loss = trainer.step(x_host.to(dev, non_blocking=True), sig_host.to(dev, non_blocking=True))
        return float(loss.mean())                          # device -> host read of the step's result

    for _ in range(max(args.warmup, 1)):
        loss = step_resident()
    _native.check_async()
    assert torch.isfinite(loss).all()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    net.set_timing(False)                                  # zero the handle's launch counters
    l0 = _native.lib().adb_launch_count(0)
    ms_res = timed(step_resident, args.steps)
    launches = _native.lib().adb_launch_count(0) - l0 + sum(v[1] for v in net.timers().values())
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if clocks else None
    _native.check_async()
    if rank == 0:
        total = B * world * args.steps
        flop_step = 3 * 606.093e9 * B * world
        line = {"metric": "diffwave_sc09_edm_train_samples_per_sec", "value": total / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms_res / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": f"DiffWave C={C} layers={LAYERS} SC09 shape 1x{L}: DSM loss forward + backward + flat-gradient "
                                       f"all-reduce + AdamW(lr 1e-4, wd 0.01), sigma ~ LogNormal(-1.2, 1.2) (BASELINE.json configs[2])",
                           "batch_per_gpu": B, "global_batch": B * world,
                           "parallelism": f"data-parallel x{world}, one NCCL all-reduce of {trainer.flat.numel() * 4 / 1e6:.1f} MB per step",
                           "l2": "saved activations (0.3 GB bf16 / 1.7 GB fp32 per sample) >> 126 MB L2"},
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 + sig_host.numel() * 4,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "effective_tflops": flop_step * args.steps / (ms_res * 1e-3) / 1e12, "clocks": clk}
        peaks = load_peaks()
        per_gpu = line["effective_tflops"] / world             # the roofline is one GPU's: whole-job rate / number of GPUs
        line["roofline"] = {"bound": "tensor", "kernel": "whole training step (z-stash forward with a bf16 stash + bf16 skip GEMM, CTA-pair cl_conv_tc dgrad, "
                                                          "wgrad_tc_pair, batched weight-norm / step-embedding tails, AdamW + refold; 3 x 606 GFLOP per sample; "
                                                          "launch list profiles/r2_launches_train_bf16_b32.csv)",
                            "achieved": per_gpu, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s (per GPU)",
                            "frac": per_gpu / peaks["bf16_tflops"], "traffic": None, "peak_source": peaks["source"]}
        if not args.no_cpu_baseline and world == 1:       # host baseline: rank 0 at N = 1 only
            from oracle import edm as oedm, wavenet as owav
            from oracle.weights import make_wavenet_state_dict
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            sd_cpu = {k: v.requires_grad_(True) for k, v in make_wavenet_state_dict(C, LAYERS, seed=0).items()}
            xc = (torch.randn(1, 1, L) * SIGMA_DATA).clamp(-1, 1)
            t0 = time.perf_counter()
            oedm.dsm_loss(xc, torch.randn(1, 1, L), torch.tensor([0.5]), owav.make_net_fn(sd_cpu, CYCLE), SIGMA_DATA).mean().backward()
            t_step = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / t_step, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "one full-size B=1 forward + backward (autograd) of the fp32 torch-CPU oracle port, no optimizer"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (default 256 DiffWave / 128 UNet1d / 32 train)")
    ap.add_argument("--workload", default="diffwave", choices=["diffwave", "unet1d", "train", "sweep"],
                    help="diffwave = the headline metric (BASELINE.json configs[1]); train = configs[2]; unet1d = configs[3]; "
                         "sweep = configs[4] (one line per global batch)")
    ap.add_argument("--sweep-batches", default="64,128,256,512,1024,2048,4096", help="global batches of --workload sweep")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in ("unet1d", "train"):
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference arm is defined for the headline DiffWave workload"}))
            return
        (run_unet if args.workload == "unet1d" else run_train)(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        args.batch = args.batch or 256
        run_reference(args, rank, world)
        return
    if args.workload == "sweep":
        # BASELINE.json configs[4]: GLOBAL batch 64 ... 4096 sharded over the N GPUs of this launch; one JSON line per batch
        batches = [int(b) for b in args.sweep_batches.split(",")]
        run_diffwave(args, rank, world, local_rank, [max(b // world, 1) for b in batches])
        return
    run_diffwave(args, rank, world, local_rank, [args.batch or 256])


def numa_pin(local_rank):
    """Pin this process to the CPU cores that are local to its GPU (nvidia-smi topo's CPU affinity), so that 8 ranks do not
    all launch from the same socket. Best effort: returns the core list used, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cores:
            per = max(len(cores) // 8, 1)                  # spread the ranks that share a socket over disjoint core groups
            mine = cores[(local_rank % 8) * per % len(cores):][:per] or cores
            os.sched_setaffinity(0, set(mine))
            return f"{mine[0]}-{mine[-1]}"
    except Exception:                                       # noqa: BLE001
        return None
    return None


def run_diffwave(args, rank, world, local_rank, batches):
    import torch.distributed as dist
    from audiodiffuser_b200 import EluDiffusion, EDMSampler, KarrasSchedule, WaveNetNoise, _native
    from audiodiffuser_b200.sharding import shard_noise

    cores = numa_pin(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)
    torch.manual_seed(0)
    net = WaveNetNoise(C, LAYERS, CYCLE, precision=args.precision)      # the reference's own init scheme (wavenet.py:75, :30)
    # the reference zero-initialises the output conv (wavenet.py:57-66): re-randomise it so the trajectory is not trivial
    net.output_projection.conv.weight.data.normal_(0.0, 1.0 / 16.0)
    net.output_projection.conv.bias.data.normal_(0.0, 0.05)
    net = net.to(dev)
    diff = EluDiffusion(sigma_data=SIGMA_DATA)
    sampler = EDMSampler(s_tmin=0, s_tmax=float("inf"), s_churn=0.0, s_noise=1.0, num_steps=STEPS_EDM, cond_scale=1.0,
                         use_heun=True)
    sigmas = KarrasSchedule(0.002, 80.0, 7.0, STEPS_EDM)().to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sweep = len(batches) > 1
    warm = max(args.warmup, 3)
    for bi, B in enumerate(batches):
        # this rank's shard of the global batch: sample g uses seed base + g, so results do not depend on world size
        noise_host = shard_noise(global_batch=B * world, rank=rank, world=world, length=L, base_seed=1234).pin_memory()
        noise_dev = noise_host.to(dev)
        out_host = torch.empty_like(noise_host).pin_memory()

        def step_resident():
            return sampler(noise_dev, fn=diff.denoise_fn, net=net, sigmas=sigmas)

        def step_e2e():
            x = noise_host.to(dev, non_blocking=True)
            y = sampler(x, fn=diff.denoise_fn, net=net, sigmas=sigmas)
            out_host.copy_(y, non_blocking=True)
            return y

        for _ in range(warm if bi == 0 else 0):       # sweep: the kernels are warm after the first batch size
            y = step_resident()
        if sweep and bi > 0:
            y = None
            diff.denoise_fn(noise_dev, net=net, sigma=1.0, inference=True)      # one evaluation: workspace for this batch size exists
        _native.check_async()
        # parity at the benchmarked configuration (outside the timed region): a waveform must not depend on the batch it was
        # sampled in — rows from the two ends and the middle of this batch, re-sampled as a batch of 3 in reversed order, must
        # come back bit-identical (different tiles, CTA-pair halves and passes compute them). The reference-golden check of a
        # full-size batch lives in tests/test_gpu_wavenet.py::test_full_size_batch_*.
        rows = sorted({0, B // 2, B - 1}, reverse=True)
        batch_invariant = None
        if y is not None:
            assert torch.isfinite(y).all() and float(y.abs().max()) > 0
            y_sub = sampler(noise_dev[rows].contiguous(), fn=diff.denoise_fn, net=net, sigmas=sigmas)
            batch_invariant = bool(torch.equal(y_sub, y[rows]))
            assert batch_invariant, "a waveform changed with its position in the batch"

        def timed(fn, steps, timing):
            net.set_timing(timing)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ms_own = e0.elapsed_time(e1)
            barrier()
            ms = ms_own
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
            return ms, ms_own

        clocks = ClockSampler(local_rank)
        clocks.start()
        ms_res, ms_res_own = timed(step_resident, args.steps, True)
        tm = net.timers()                                    # per-kernel-class CUDA-event time, this rank, timed region
        ms_e2e = None if sweep else timed(step_e2e, args.steps, False)[0]      # the sweep reports the resident number only
        clk = clocks.stop()
        _native.check_async()
        per_rank = None
        if world > 1:
            # every rank's own step time, dominant-kernel time and clocks: names the slowest rank of a weak-scaling run
            mine = torch.tensor([ms_res_own / args.steps, tm["conv"][0] / max(tm["conv"][1], 1), tm["skip"][0] / max(tm["skip"][1], 1),
                                 clk["sm_mhz"] or 0.0, 1.0 if "sw_power_cap" in clk["reasons"] else 0.0,
                                 1.0 if any(r != "sw_power_cap" for r in clk["reasons"]) else 0.0], device=dev, dtype=torch.float64)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = [{"rank": r, "ms_per_step": float(v[0]), "block_kernel_avg_ms": float(v[1]), "skip_gemm_avg_ms": float(v[2]),
                         "sm_mhz_median": float(v[3]), "sw_power_cap": bool(v[4]), "other_throttle": bool(v[5])} for r, v in enumerate(allr)]
        if rank != 0:
            continue
        peaks = load_peaks()
        total = B * world * args.steps
        value = total / (ms_res * 1e-3)
        e2e = total / (ms_e2e * 1e-3) if ms_e2e else None
        conv_ms, conv_n = tm["conv"]
        skip_ms, skip_n = tm["skip"]
        step_ms, step_n = tm["step"]
        launches = sum(v[1] for v in tm.values())
        zs = BLOCK_KERNEL == 3
        flop_blocks = FLOP_EVAL_ZS_BLOCKS if zs else FLOP_EVAL_BLOCKS
        conv_tflops = (B * flop_blocks * NFE * args.steps) / (conv_ms * 1e-3) / 1e12 if conv_ms else None
        stack_ms = conv_ms + skip_ms
        stack_tflops = (B * FLOP_EVAL_BLOCKS * NFE * args.steps) / (stack_ms * 1e-3) / 1e12 if stack_ms else None
        step_bytes = STEP_BYTES_PER_ELEM * B * L
        n_step_main = (2 * STEPS_EDM - 1) * args.steps   # mid + post (+ final Euler) kernels
        step_gbs_l2 = (step_bytes * n_step_main) / (step_ms * 1e-3) / 1e9 if step_ms else None
        traffic = load_traffic()
        kname = {3: "wavenet_block_zs_kernel", 2: "wavenet_block_pair_kernel", 1: "wavenet_block_pair_kernel", 0: "wavenet_block_tc_kernel"}[BLOCK_KERNEL]
        roofline = {"bound": "tensor", "kernel": kname,
                    "achieved": conv_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": conv_tflops / peaks["bf16_tflops"] if conv_tflops else None,
                    "traffic": traffic.get("dram_bytes_per_launch_B256") if traffic.get("kernel") == kname else None,
                    "traffic_note": traffic.get("note") if traffic.get("kernel") == kname else None,
                    "peak_source": peaks["source"],
                    "launches": conv_n, "avg_launch_ms": conv_ms / conv_n if conv_n else None,
                    "flop_per_launch": B * flop_blocks / LAYERS,
                    "residual_stack": {"what": "block launches + skip GEMM together on ALL the stack's algorithmic FLOPs (36 x 16.777 G - 2.097 G per sample)",
                                       "achieved": stack_tflops, "frac": stack_tflops / peaks["bf16_tflops"] if stack_tflops else None,
                                       "ms": stack_ms}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm if bi == 0 else 0, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no collective",
                           "l2": "working set per block launch (h 2 x 2.1 GB + stash 2.1 GB per 256-sample pass) >> 126 MB L2: no flush needed",
                           "cpu_affinity": cores},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": noise_host.numel() * 4,
                        "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / args.steps if ms_e2e else None},
                "gpu_launches": int(launches),
                "roofline": roofline,
                "kernel_ms": {k: v[0] for k, v in tm.items()},
                "parity": {"batch_position_invariant_rows": rows, "bit_identical": batch_invariant},
                "clocks": clk}
        if zs and skip_n:
            skip_tflops = (B * FLOP_EVAL_SKIP_GEMM * NFE * args.steps) / (skip_ms * 1e-3) / 1e12
            line["roofline_skip_gemm"] = {"bound": "tensor", "kernel": "wavenet_skip_gemm_kernel", "achieved": skip_tflops,
                                          "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": skip_tflops / peaks["bf16_tflops"],
                                          "launches": skip_n, "avg_launch_ms": skip_ms / skip_n,
                                          "hbm_gbs": (B * L * 512.0 * LAYERS * NFE * args.steps) / (skip_ms * 1e-3) / 1e9,
                                          "note": "reads the 36-layer fp16 stash once (HBM) while contracting it: both roofs are close"}
        if per_rank:
            line["per_rank"] = per_rank
        if len(batches) == 1:
            big = step_kernel_hbm(dev)
            line["roofline_step_kernel"] = {
                "bound": "hbm", "kernel": "edm_kernel<OP_MID> + edm_kernel<OP_POST> (adb_edm_heun_mid/post)",
                "achieved": big["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": big["gbs"] / peaks["hbm_gbs"], "bytes_per_launch": big["bytes_per_launch"],
                "launches": big["launches"], "avg_launch_ms": big["avg_ms"],
                "note": f"timed on a {big['elements']}-element state (268 MB per array > 126 MB L2, rotating buffers); "
                        f"inside the trajectory the B*L state (16 MB at B=256) is L2-resident and the same kernels "
                        f"run at {step_gbs_l2:.0f} GB/s effective, launch-latency bound ({step_n} launches, "
                        f"{step_ms:.2f} ms total)"}
        if not args.no_cpu_baseline and world == 1 and len(batches) == 1:       # host baseline: rank 0 at N = 1 only
            threads = os.cpu_count() or 1
            kind, t = cpu_reference_eval_time(4, threads)
            t_eval = sum(t[1:]) / len(t[1:])
            line["cpu_baseline"] = {"value": CPU_BATCH / (NFE * t_eval), "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"B={CPU_BATCH} x 3 full-size denoiser calls after 1 warm-up (of {NFE} per trajectory, "
                                              f"BASELINE.json configs[0] shape), fp32, {CPU_KIND_TEXT[kind]}, extrapolated x{NFE}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
