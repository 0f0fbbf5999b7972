"""CPU restatement (numpy) of the in-kernel churn noise of `adb_edm_churn_rng` (csrc/edm_kernels.cuh) — TEST INFRASTRUCTURE.

The reference draws `epsilon = torch.randn_like(x)` per sampler step (src/models/components/sampler_edm.py:346); the CUDA path
draws the same distribution inside the update kernel with the counter-based generator Philox4x32-10 (Salmon et al., "Parallel
random numbers: as easy as 1, 2, 3", SC'11 — the generator behind curand's Philox and torch's CUDA RNG), pinned here by the
known-answer vectors of the Random123 distribution (tests/test_host_logic.py). Counter = (group of 4 elements, sample index low
word, step, sample index high word), key = 64-bit seed; four N(0,1) values per block by Box-Muller on 24-bit uniforms.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over equally shaped uint32 arrays (or scalars). Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & MASK, p0 & MASK, n0, n2
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def churn_normals(seed, sample, step, n_per):
    """The n_per N(0,1) values the kernel uses for global sample `sample` at sampler step `step` (fp32)."""
    groups = np.arange((n_per + 3) // 4, dtype=np.uint64)
    r = philox4x32_10(groups, sample & 0xFFFFFFFF, step, (sample >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((groups.size, 4), dtype=np.float32)
    scale = np.float32(2.0 ** -24)
    for h in range(2):
        u1 = ((r[2 * h] >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * scale
        u2 = (r[2 * h + 1] >> np.uint32(8)).astype(np.float32) * scale
        rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
        ang = (np.float64(2.0) * np.pi * u2.astype(np.float64))
        out[:, 2 * h] = rad * np.cos(ang).astype(np.float32)
        out[:, 2 * h + 1] = rad * np.sin(ang).astype(np.float32)
    return out.reshape(-1)[:n_per]
