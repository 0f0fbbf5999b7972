"""CPU restatement of the 16-bit WAV output (test infrastructure only, see oracle/__init__).

The reference calls `torchaudio.save(path, wav[None, :], sample_rate, bits_per_sample=16)`
(src/models/diffunet_complex_module.py:263-266). torchaudio is a third-party dependency that is absent from this image
(requirements.txt pins no version; README installs the one matching torch>=2.0), so its published conversion rule is
restated: the FFmpeg writer (default backend since torchaudio 2.1) converts with libswresample/audioconvert.c's
`av_clip_int16(lrintf(x * (1 << 15)))`, i.e. round-half-to-even of x * 2^15, saturated — the rule implemented here. The older
SoX writer (sox.h SOX_FLOAT_32BIT_TO_SAMPLE, then SOX_SAMPLE_TO_SIGNED_16BIT = add 2^15, shift) rounds exact ties up instead
and differs by one LSB on odd multiples of 2^-16 only; the soundfile backend scales by 0x7FFF and can differ by 1 LSB anywhere.
PARITY UNPINNED against torchaudio itself (not runnable here); pinned against an independently assembled fixture (exact
rational arithmetic + the stdlib `wave` writer, oracle/make_wav_fixture.py -> tests/golden/wav16_fixture.*) and
hand-computed known answers in tests/test_wav_module.py.
"""
import numpy as np


def pcm16(x: np.ndarray) -> np.ndarray:
    q = np.rint(np.asarray(x, dtype=np.float32) * np.float32(32768.0))      # rint = round half to even, like lrintf
    return np.clip(q, -32768, 32767).astype(np.int16)


def read_wav16(path: str):
    """(int16 array [channels, frames], sample_rate) through the standard library's RIFF parser."""
    import wave
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getcomptype() == "NONE"
        ch, sr, n = w.getnchannels(), w.getframerate(), w.getnframes()
        data = np.frombuffer(w.readframes(n), dtype="<i2").reshape(n, ch).T
    return data, sr
