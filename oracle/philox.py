"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11 "Parallel random numbers: as easy as 1,2,3").

TEST INFRASTRUCTURE.  The reference draws from Python ``random`` / ``np.random`` one scalar at a time
(srl/algorithms/dqn/dqn.py:200-202, srl/envs/grid.py:174,203,
srl/rl/memories/priority_memories/proportional_memory.py:147); a parallel device cannot replay those
streams, so the device uses Philox counters and this module reproduces the identical words on the CPU
so that every function can be compared *given identical pre-drawn uniforms* (SURVEY.md section 7 "RNG").
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)

# stream ids (word 3 of the counter) -- must match csrc/philox.cuh
STREAM_ENV_RESET = 1
STREAM_ENV_STEP = 2
STREAM_POLICY = 3
STREAM_NOISE = 4
STREAM_SAMPLE = 5
STREAM_PAD_ACTION = 6
STREAM_UNIFORM_SAMPLE = 7


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = c0.astype(np.uint64) * M0
            p1 = c2.astype(np.uint64) * M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def seed_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def words(seed, stream, a, b=0, c=0):
    k0, k1 = seed_key(seed)
    return philox4x32(a, b, c, np.uint32(stream), k0, k1)


def u01_f32(w):
    """24-bit uniform in [0,1) as float32 (exact)."""
    return ((np.asarray(w, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def u01_f64(w_hi, w_lo):
    """53-bit uniform in [0,1) as float64 (exact): (hi>>5)*2^26 + (lo>>6), scaled by 2^-53."""
    a = (np.asarray(w_hi, dtype=np.uint32) >> np.uint32(5)).astype(np.float64)
    b = (np.asarray(w_lo, dtype=np.uint32) >> np.uint32(6)).astype(np.float64)
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)
