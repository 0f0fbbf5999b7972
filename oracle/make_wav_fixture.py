"""Write tests/golden/wav16_fixture.{wav,npz}: a float32 vector (ties on both sides of zero, +-1.0, values past full scale, +-inf,
NaN, denormals) and the 16-bit file it must produce under the rule this repo implements, assembled with the standard
library's `wave` writer (an independent container implementation) from integers computed with exact rational arithmetic
(`fractions.Fraction`, no floating-point rounding function involved).

The rule — q = round-half-to-even(x * 2^15), saturated to [-32768, 32767], NaN -> 0 — is what the reference's call
`torchaudio.save(path, wav, sr, bits_per_sample=16)` (src/models/diffunet_complex_module.py:263-266) does in torchaudio's
FFmpeg writer, the default backend since torchaudio 2.1: FFmpeg libswresample/audioconvert.c converts FLT -> S16 with
`av_clip_int16(lrintf(*(const float*)pi * (1 << 15)))` (lrintf rounds to nearest-even in the default rounding mode).
torchaudio's older SoX writer goes through a 32-bit sample (sox.h SOX_FLOAT_32BIT_TO_SAMPLE, then SOX_SAMPLE_TO_SIGNED_16BIT
= add 2^15 and shift) and therefore rounds exact ties UP instead of to even: it differs from the rule above by one LSB on
inputs that are odd multiples of 2^-16 exactly, and nowhere else. torchaudio is not in this image, so neither library was
run: the fixture pins this repo's encoder and container to an independently assembled file, NOT to torchaudio's bytes
(PARITY UNPINNED against torchaudio itself).

    python oracle/make_wav_fixture.py
"""
import math
import os
import struct
import wave
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def exact_pcm16(v: float) -> int:
    if math.isnan(v):
        return 0
    if math.isinf(v):
        return 32767 if v > 0 else -32768
    q = Fraction(v) * 32768                       # exact: a float32 is a dyadic rational
    fl = q.numerator // q.denominator             # floor
    rem = q - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return max(-32768, min(32767, fl))


def main():
    ties = [k + 0.5 for k in (-32769, -32768, -3, -2, -1, 0, 1, 2, 3, 16383, 32766, 32767)]
    vals = [t / 32768.0 for t in ties] + [0.0, -0.0, 1.0, -1.0, 0.999969482421875, 0.99998474, -0.99998474, 1.0000001, 2.0, -2.0,
                                            1e-45, -1e-45, 1e-30, 3.0517578125e-05, -3.0517578125e-05, 0.25, -0.75, 0.1, -0.1,
                                            float("inf"), float("-inf"), float("nan")]
    rng = np.random.default_rng(20261018)
    vals += list(rng.uniform(-1.1, 1.1, size=2000).astype(np.float32))
    x = np.asarray(vals, dtype=np.float32)
    pcm = np.asarray([exact_pcm16(float(v)) for v in x], dtype=np.int16)
    frames = x.size // 2
    stereo = pcm[:2 * frames].reshape(2, frames)                      # [channels, frames], torchaudio's layout
    with wave.open(os.path.join(OUT, "wav16_fixture.wav"), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(b"".join(struct.pack("<hh", int(stereo[0, i]), int(stereo[1, i])) for i in range(frames)))
    np.savez(os.path.join(OUT, "wav16_fixture.npz"), x=x, pcm=pcm)
    print(x.size, "values,", frames, "stereo frames")


if __name__ == "__main__":
    main()
