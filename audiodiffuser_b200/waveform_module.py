"""Waveform-domain caller of the hot path: the 1-D twin of the reference's LightningModule.

The reference ships `DiffUnetComplexModule` (src/models/diffunet_complex_module.py:13-290) for complex-STFT U-Nets and
points `configs/train.yaml:8` at a `model: diffwave.yaml` that is missing from the public tree. This module is that
missing caller for raw waveforms: the same constructor groups (net / noise_scheduler / noise_distribution / sampler /
diffusion / optimizer / scheduler), the same hooks with the same bodies minus the STFT front end:

  forward(batch)                diffunet_complex_module.py:105-125   sigmas ~ noise_distribution; diffusion(x, net, classes, sigmas).mean()
  training_step                 :144-171    loss, fp16 EMA snapshots every `num_ema_snapshot_item` items (rank 0), EMA update
  validation_step / epoch end   :185-219    loss; one generated sample written as WAV by global rank 0
  on_test_epoch_end             :231-266    `total_test_samples` waveforms written as 16-bit WAV `test_<class>_<index>.wav`
  configure_optimizers          :268-290    optimizer(params=...), optional scheduler dict monitoring "val/loss"

What differs, deliberately:
  * test sampling is SHARDED: the reference generates the same `total_test_samples` files on every DDP rank (every rank
    is seeded identically, src/train.py:50-51, and writes the same names). Here rank r generates the contiguous index
    range `shard_range(total, r, world)`; sample g always starts from `Generator(base_seed + g)` noise, so the files are
    identical for any world size, and each file is written once.
  * float -> PCM16 runs on the device (`adb_pcm16_encode`); the host only writes the RIFF container.
  * every sampler/denoiser/loss call goes through the fused CUDA path of this package (no torch fallback).

Lightning is not part of this image: with `pytorch_lightning` installed the class derives from `LightningModule` and is
driven by `Trainer` unchanged; without it, it derives from `_StandaloneModule`, which supplies the handful of attributes
the hooks read (`device`, `log`, `global_step`, `trainer.global_rank / is_global_zero / world_size`, `logger.save_dir`) and
`fit_steps()` / `test()` drive the same hooks in-process.
"""
import os
import pickle
from types import SimpleNamespace
from typing import Any, Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch import Tensor

from . import sharding, wav
from .ema import PowerFunctionEMA, TraditionalEMA

try:                                                    # pragma: no cover - not installed in this image
    from pytorch_lightning import LightningModule as _LightningBase
    HAVE_LIGHTNING = True
except ImportError:
    _LightningBase = None
    HAVE_LIGHTNING = False


class _StandaloneModule(nn.Module):
    """The slice of LightningModule the hooks below use, for environments without Lightning."""

    def __init__(self):
        super().__init__()
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.trainer = SimpleNamespace(global_rank=rank, world_size=world, is_global_zero=(rank == 0),
                                       datamodule=SimpleNamespace(batch_size=8))
        self.logger = SimpleNamespace(save_dir=os.getcwd())
        self.global_step = 0
        self.logged = {}

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def log(self, name, value, **kwargs):
        self.logged[name] = float(value.detach()) if isinstance(value, Tensor) else float(value)


_Base = _LightningBase if HAVE_LIGHTNING else _StandaloneModule


def test_sample_filename(target_class: int, global_index: int) -> str:
    """'test_<class>_<index>.wav' (diffunet_complex_module.py:263)."""
    return f"test_{int(target_class)}_{int(global_index)}.wav"


def test_sample_classes(start: int, stop: int, test_batch: int, generated_sample_class: int):
    """Class label of each global sample index: the reference labels position j of every batch `j % classes`
    (diffunet_complex_module.py:252-255), i.e. `(g % test_batch) % classes`."""
    if generated_sample_class > 1:
        return [(g % test_batch) % generated_sample_class for g in range(start, stop)]
    return [0] * (stop - start)


def plan_test_shard(total_test_samples: int, test_batch: int, rank: int, world: int, generated_sample_class: int):
    """The batches rank `rank` generates at test time: a list of (global indices, class labels, file names).
    The reference generates `total // test_batch` full batches (diffunet_complex_module.py:236) on EVERY rank; here the
    same index space is split into contiguous per-rank ranges, walked in chunks of at most `test_batch`."""
    total = (total_test_samples // test_batch) * test_batch
    start, stop = sharding.shard_range(total, rank, world)
    plan = []
    for lo in range(start, stop, test_batch):
        hi = min(lo + test_batch, stop)
        labels = test_sample_classes(lo, hi, test_batch, generated_sample_class)
        names = [test_sample_filename(c, g) for c, g in zip(labels, range(lo, hi))]
        plan.append((list(range(lo, hi)), labels, names))
    return plan


class DiffWaveformModule(_Base):
    def __init__(
        self,
        net: nn.Module,
        noise_scheduler,
        noise_distribution: nn.Module,
        sampler: nn.Module,
        diffusion: nn.Module,
        optimizer,
        scheduler,
        generated_length: int,
        generated_sample_class: int,
        audio_sample_rate: int,
        audio_channels: int = 1,
        norm_wav: bool = False,
        use_ema: bool = True,
        use_phema: bool = False,
        num_ema_snapshot_item: Optional[int] = 96000,
        total_test_samples: Optional[int] = None,
        ema_ckpt_path: Optional[str] = None,
        base_seed: int = 0,
    ):
        super().__init__()
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.net = net
        self.use_ema = use_ema
        self.use_phema = use_phema
        self.cur_nitem = 0
        self.num_ema_snapshot_item = num_ema_snapshot_item
        self.ema_ckpt_path = ema_ckpt_path
        self.sampler = sampler
        self.diffusion = diffusion
        self.noise_distribution = noise_distribution            # for training
        while callable(noise_scheduler):                        # module (or partial of one) -> sigma tensor, evaluated once
            noise_scheduler = noise_scheduler()                 # like `self.noise_scheduler = noise_scheduler()` (:64)
        self.noise_scheduler = noise_scheduler                  # for sampling
        self.generated_length = generated_length
        self.generated_sample_class = generated_sample_class
        self.audio_channels = audio_channels
        self.total_test_samples = total_test_samples
        self.audio_sample_rate = audio_sample_rate
        self.norm_wav = norm_wav
        self.base_seed = base_seed
        self.ema_prof = None
        self._val_sum, self._val_n, self.val_loss_best = 0.0, 0, float("inf")

    # ---- sampling ------------------------------------------------------------------------------------------------
    def _sigmas(self) -> Tensor:
        return self.noise_scheduler.to(self.device)

    @torch.no_grad()
    def _synthesize_device(self, initial_noise: Tensor, target_class) -> Tensor:
        """[B, C, L] N(0,1) noise -> [B, C, L] waveforms on the device (diffunet_complex_module.py:82-89)."""
        return self.sampler(initial_noise, classes=target_class, fn=self.diffusion.denoise_fn, net=self.net,
                            sigmas=self._sigmas())

    @torch.no_grad()
    def synthesize_from_noise(self, initial_noise: Tensor, target_class, ema_model=None) -> Tensor:
        """CPU waveforms [B, L] (mono) or [B, C, L], like the reference's hook (:82-103)."""
        x = self._synthesize_device(initial_noise, target_class).cpu()
        return x[:, 0] if x.shape[1] == 1 else x

    # ---- training ------------------------------------------------------------------------------------------------
    def forward(self, x: Any) -> Tensor:
        audio_classes = x.get("label") if isinstance(x, dict) else None
        audio = x["audio"] if isinstance(x, dict) else x
        audio = audio.to(torch.float32)
        if audio.ndim == 2:
            audio = audio[:, None, :]                                       # [B, L] -> [B, 1, L]
        sigmas = self.noise_distribution(num_samples=audio.shape[0], device=audio.device)
        loss = self.diffusion(audio, self.net, classes=audio_classes, sigmas=sigmas)
        return loss.mean()

    def on_fit_start(self):
        if self.use_ema and self.use_phema:
            self.ema_prof = PowerFunctionEMA(self.net.to(self.device), stds=[0.050, 0.100])
        elif self.use_ema:
            self.ema_prof = TraditionalEMA(self.net.to(self.device), halflife_Mimg=0.3, rampup_ratio=0.09)

    def model_step(self, batch: Any) -> Tensor:
        return self.forward(batch)

    def write_ema_snapshots(self) -> list:
        """fp16 pickles `ema_snapshots/ema_prof<suffix>_<global_step>` (diffunet_complex_module.py:158-167)."""
        ema_list = self.ema_prof.get()
        ema_list = ema_list if isinstance(ema_list, list) else [(ema_list, "")]
        folder = os.path.join(self.logger.save_dir, "ema_snapshots")
        os.makedirs(folder, exist_ok=True)
        paths = []
        for ema_net, suffix in ema_list:
            snap = ema_net.cpu().eval().requires_grad_(False).to(torch.float16)   # get() already returned a private copy
            path = os.path.join(folder, f"ema_prof{suffix}_{self.global_step}")
            with open(path, "wb") as f:
                pickle.dump(snap, f)
            paths.append(path)
        return paths

    def training_step(self, batch: Any, batch_idx: int):
        loss = self.model_step(batch)
        self.log("train/loss", loss, on_step=True, on_epoch=True, prog_bar=True, sync_dist=True)
        self.log("seen items", self.cur_nitem * 1.0, on_step=True, prog_bar=True, sync_dist=True)
        if self.use_ema and self.ema_prof is not None:
            first = batch[list(batch.keys())[0]] if isinstance(batch, dict) else batch
            batch_size = first.shape[0]
            if (self.num_ema_snapshot_item and int(self.cur_nitem) % self.num_ema_snapshot_item == 0
                    and self.trainer.global_rank == 0 and self.global_step > 0):
                self.write_ema_snapshots()
            self.cur_nitem += batch_size
            self.ema_prof.update(self.cur_nitem, batch_size)
        return {"loss": loss}

    @torch.no_grad()
    def validation_step(self, batch: Any, batch_idx: int):
        loss = self.model_step(batch)
        self._val_sum += float(loss)
        self._val_n += 1
        self.log("val/loss", self._val_sum / self._val_n, on_step=False, on_epoch=True, prog_bar=True, sync_dist=True)
        return {"loss": loss}

    @torch.no_grad()
    def on_validation_epoch_end(self):
        if self._val_n:
            self.val_loss_best = min(self.val_loss_best, self._val_sum / self._val_n)
        self._val_sum, self._val_n = 0.0, 0
        self.log("val/loss_best", self.val_loss_best, prog_bar=True, sync_dist=True)
        classes = self.generated_sample_class if self.generated_sample_class > 1 else 1
        target = int(torch.randint(classes, (1,)).item())
        noise = torch.randn((1, self.audio_channels, self.generated_length), device=self.device)
        x = self._synthesize_device(noise, torch.tensor([target], device=self.device))
        if self.trainer.is_global_zero:
            folder = os.path.join(self.logger.save_dir, "val_audio")
            os.makedirs(folder, exist_ok=True)
            self._write(os.path.join(folder, f"val_{target}_{self.global_step}.wav"), x[0], self.generated_length)

    # ---- test-time generation --------------------------------------------------------------------------------------
    def _write(self, path: str, waveform: Tensor, frames: int):
        """waveform: CUDA fp32 [C, L] -> 16-bit WAV of its first `frames` frames."""
        w = waveform[:, :frames]
        if self.norm_wav:
            w = w / w.abs().amax().clamp_min(1e-8)
        wav.write_wav16(path, wav.pcm16_encode(w).cpu(), self.audio_sample_rate)

    def test_step(self, batch: Any, batch_idx: int):
        pass                                                               # like the reference (:221-229)

    @torch.no_grad()
    def on_test_epoch_end(self):
        """Generate this rank's shard of the `total_test_samples` test waveforms (diffunet_complex_module.py:231-266)."""
        test_batch = self.trainer.datamodule.batch_size
        audio_dur = 1
        folder = os.path.join(self.logger.save_dir, "test_samples")
        if self.ema_ckpt_path is not None:                                 # override the weights with an EMA snapshot
            with open(self.ema_ckpt_path, "rb") as f:
                self.net = pickle.load(f).to(torch.float32).to(self.device)
        os.makedirs(folder, exist_ok=True)
        frames = min(int(audio_dur * self.audio_sample_rate), self.generated_length)
        written = []
        for indices, labels, names in plan_test_shard(self.total_test_samples, test_batch, self.trainer.global_rank,
                                                      getattr(self.trainer, "world_size", 1), self.generated_sample_class):
            noise = sharding.noise_for_indices(indices, self.generated_length, self.base_seed,
                                               self.audio_channels).to(self.device, non_blocking=True)
            x = self._synthesize_device(noise, torch.tensor(labels, device=self.device))
            if self.norm_wav:
                x = x / x.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-8)
            pcm = wav.pcm16_encode(x[:, :, :frames]).cpu()                 # one D2H of int16 per batch
            for j, name in enumerate(names):
                wav.write_wav16(os.path.join(folder, name), pcm[j], self.audio_sample_rate)
                written.append(os.path.join(folder, name))
        return written

    def configure_optimizers(self):
        optimizer = self.optimizer(params=self.parameters())
        if self.scheduler is not None:
            scheduler = self.scheduler(optimizer=optimizer)
            return {"optimizer": optimizer,
                    "lr_scheduler": {"scheduler": scheduler, "monitor": "val/loss", "interval": "epoch", "frequency": 1}}
        return {"optimizer": optimizer}

    # ---- in-process driver (only meaningful without Lightning) -------------------------------------------------------
    def fit_steps(self, batches: Iterable[Any], save_dir: Optional[str] = None):
        """Run the training hooks over `batches` with the configured optimizer; returns the per-step losses.
        With world size > 1 the gradients are averaged with one all-reduce of the flat gradient per step."""
        if save_dir is not None:
            self.logger.save_dir = save_dir
        opt = self.configure_optimizers()["optimizer"]
        self.on_fit_start()
        world = getattr(self.trainer, "world_size", 1)
        losses = []
        for i, batch in enumerate(batches):
            opt.zero_grad(set_to_none=True)
            loss = self.training_step(batch, i)["loss"]
            loss.backward()
            if world > 1:                                               # one all-reduce of the flattened gradients per step
                grads = [p.grad for p in self.parameters() if p.grad is not None]
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.div_(world)
                torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])
            opt.step()
            self.global_step += 1
            losses.append(float(loss.detach()))
        return losses

    def test(self, save_dir: Optional[str] = None, batch_size: Optional[int] = None):
        if save_dir is not None:
            self.logger.save_dir = save_dir
        if batch_size is not None:
            self.trainer.datamodule.batch_size = batch_size
        return self.on_test_epoch_end()
