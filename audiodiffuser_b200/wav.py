"""16-bit WAV output of generated waveforms — the data format on the far side of the sampling path.

The reference writes each test sample with `torchaudio.save(path, wav[None, :], sr, bits_per_sample=16)`
(src/models/diffunet_complex_module.py:263-266). torchaudio's float -> s16 rule (ffmpeg / sox backends) is
round-half-even(x * 2^15) saturated to [-32768, 32767]; here that conversion runs on the device
(`adb_pcm16_encode`, 4 B read + 2 B written per sample) so only the int16 PCM crosses PCIe, and the RIFF container
(44-byte canonical header + little-endian PCM) is written by the host.
"""
import struct

import torch
from torch import Tensor

from . import _native as N


def pcm16_encode(x: Tensor) -> Tensor:
    """fp32 CUDA tensor (any shape) -> int16 CUDA tensor of the same shape. No CPU path."""
    x = N.require_cuda_f32(x, "x").contiguous()
    out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    if x.numel():
        N.check(N.lib().adb_pcm16_encode(N.ptr(x), N.ptr(out), x.numel(), N.stream_ptr(x.device)))
    return out


def wav16_header(num_frames: int, sample_rate: int, channels: int = 1) -> bytes:
    """Canonical 44-byte RIFF/WAVE header for 16-bit PCM."""
    if num_frames < 0 or sample_rate <= 0 or channels <= 0:
        raise ValueError("num_frames >= 0, sample_rate > 0 and channels > 0 required")
    data_bytes = num_frames * channels * 2
    if data_bytes > 0xFFFFFFFF - 36:
        raise ValueError("waveform too long for a RIFF container")
    return (b"RIFF" + struct.pack("<I", 36 + data_bytes) + b"WAVE"
            + b"fmt " + struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * channels * 2, channels * 2, 16)
            + b"data" + struct.pack("<I", data_bytes))


def write_wav16(path: str, pcm: Tensor, sample_rate: int) -> None:
    """pcm: int16 CPU tensor [frames] (mono) or [channels, frames] (torchaudio's layout) -> file at `path`."""
    if pcm.dtype != torch.int16 or pcm.is_cuda:
        raise TypeError("write_wav16 expects an int16 CPU tensor (use pcm16_encode(...).cpu())")
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    if pcm.ndim != 2:
        raise ValueError(f"expected [frames] or [channels, frames]; got {tuple(pcm.shape)}")
    channels, frames = pcm.shape
    body = pcm.t().contiguous().numpy().astype("<i2", copy=False).tobytes()      # interleave channels, little endian
    with open(path, "wb") as f:
        f.write(wav16_header(frames, sample_rate, channels))
        f.write(body)
