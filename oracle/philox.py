"""Philox4x32-10 in NumPy (Salmon et al., SC'11) — independent restatement of the device RNG.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Counter layout and uniform construction must match
gymrl_b200/csrc/common.cuh: counter = (entity_lo, draw, stream, entity_hi), key = (seed_lo, seed_hi).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_ENV_RESET, STREAM_ENV_STEP, STREAM_ACTION, STREAM_PERMUTE, STREAM_REPLAY, STREAM_NOISYNET, STREAM_UPDATE = range(7)


def philox(seed, entity, draw, stream):
    """Vectorised over `entity` / `draw` (broadcast).  Returns uint32 array [..., 4]."""
    entity = np.asarray(entity, dtype=np.uint64)
    draw = np.asarray(draw, dtype=np.uint64)
    entity, draw = np.broadcast_arrays(entity, draw)
    seed = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
    c0 = (entity & MASK32).astype(np.uint64)
    c1 = (draw & MASK32).astype(np.uint64)
    c2 = np.full(entity.shape, np.uint64(stream), dtype=np.uint64)
    c3 = (entity >> np.uint64(32)).astype(np.uint64)
    k0 = np.uint64(seed & MASK32)
    k1 = np.uint64(seed >> np.uint64(32))
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & MASK32
            n1 = p1 & MASK32
            n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & MASK32
            n3 = p0 & MASK32
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = (k0 + np.uint64(W0)) & MASK32
            k1 = (k1 + np.uint64(W1)) & MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u01_f64(a, b):
    """53-bit uniform in [0,1) from two uint32 words (numpy Generator.random construction)."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def u01_f32(a):
    return ((np.asarray(a, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def u01_open0_f32(a):
    return ((np.asarray(a, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
