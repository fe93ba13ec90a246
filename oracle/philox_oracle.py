"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the in-kernel noise generator
(polyffusion_b200/csrc/kernels.cu: philox4x32_10, philox_normal, fused_step_apply).

Philox4x32-10 is the counter-based generator of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as
easy as 1, 2, 3" (SC'11); the reference (aik2mlj/polyffusion) draws its noise with torch.randn, whose CUDA
backend is the same Philox4x32-10 with a different counter layout -- this generator is NOT bit-compatible
with torch's stream, it is the rank-count-invariant alternative of SURVEY.md section 8(e).  Pinned against
the Random123 known-answer vectors in tests/test_philox.py.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
GOLDEN64 = 0x9E3779B97F4A7C15
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (broadcast); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK32 for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def effective_seed(seed: int, nonce: int) -> int:
    return (int(seed) + int(nonce) * GOLDEN64) & ((1 << 64) - 1)


def normal(seed: int, sample0: int, n_samples: int, per_sample: int, index: int, which: int) -> np.ndarray:
    """[n_samples, per_sample] fp32 normals: counter (element, sample0 + b, index, which), key = seed;
    Box-Muller on the first two words (fp32 arithmetic, as the kernel)."""
    elem = np.arange(per_sample, dtype=np.uint64)[None, :]
    samp = (np.arange(n_samples, dtype=np.uint64) + np.uint64(sample0))[:, None]
    r0, r1, _, _ = philox4x32_10(elem, samp, np.uint64(index), np.uint64(which), seed & 0xFFFFFFFF, seed >> 32)
    scale = np.float32(5.9604644775390625e-08)
    u1 = ((r0 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * scale
    u2 = ((r1 >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * scale
    rad = np.sqrt(np.float32(-2.0) * np.log(u1), dtype=np.float32)
    return (rad * np.cos(np.float32(6.283185307179586) * u2, dtype=np.float32)).astype(np.float32)
