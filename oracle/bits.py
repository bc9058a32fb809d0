"""Bit <-> secret-tensor mapping of IDEAS in numpy (TEST INFRASTRUCTURE).

Restates utils.py:74-97 and the BER of train.py:285-286 of the reference with the
exact fp32 operation order, because decode is *not* ``z >= 0``: ``(-1e-9) + 1`` rounds
to 1.0f and yields bit 1 whereas ``-6e-8`` yields bit 0.

Randomness is an explicit argument (``u`` = the uniform draws ``torch.rand_like`` would
have produced) so the functions are pure.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def message_to_tensor(message: np.ndarray, sigma: int, delta: float, u: np.ndarray | None = None) -> np.ndarray:
    """utils.py:74-83.  message (B, sigma*L) of 0/1 -> secret tensor (B, L) in [-1, 1].

    Groups of ``sigma`` bits, MSB first (``message[:, i::sigma] * 2**(sigma-i-1)``), become
    an integer n; z = step*(n+0.5) - 1 with step = 2/2**sigma; then a jitter
    ``u*r*2 - r`` with r = step*delta is added, evaluated left to right in fp32."""
    m = np.asarray(message, dtype=f32)
    B, total = m.shape
    L = total // sigma
    step = 2 / 2 ** sigma
    r = step * delta
    nums = np.zeros((B, L), dtype=f32)
    for i in range(sigma):
        nums = nums + m[:, i::sigma][:, :L] * f32(2 ** (sigma - i - 1))
    z = f32(step) * (nums + f32(0.5)) - f32(1)
    if u is None:
        u = np.full((B, L), 0.5, dtype=f32)                    # centre of the interval: zero jitter
    jitter = (np.asarray(u, dtype=f32) * f32(r)) * f32(2) - f32(r)
    return (z + jitter).astype(f32)


def tensor_to_message(z: np.ndarray, sigma: int) -> np.ndarray:
    """utils.py:86-97.  (B, L) fp32 -> (B, sigma*L) float 0/1 by threshold peeling:
    t = (clamp(z,-1,1)+1)/step ; for i: bit = t >= 2**(sigma-i-1) ; t -= bit*2**(sigma-i-1)."""
    z = np.asarray(z, dtype=f32)
    B, L = z.shape
    step = 2 / 2 ** sigma
    t = (np.clip(z, f32(-1), f32(1)) + f32(1)) / f32(step)
    msg = np.zeros((B, L * sigma), dtype=f32)
    for i in range(sigma):
        w = f32(2 ** (sigma - i - 1))
        bit = (t >= w).astype(f32)
        msg[:, i::sigma] = bit
        t = t - bit * w
    return msg


def ber(message: np.ndarray, decoded: np.ndarray) -> float:
    """train.py:285: mean |M - M_hat| over all bits."""
    return float(np.mean(np.abs(np.asarray(message, dtype=f32) - np.asarray(decoded, dtype=f32))))


# -- packed representation used by the integer kernels --------------------------------
def pack_bits(message: np.ndarray) -> np.ndarray:
    """(B, nbits) 0/1 -> (B, ceil(nbits/32)) uint32, bit j of word w = message[:, 32*w + j]."""
    m = (np.asarray(message) != 0)
    B, n = m.shape
    words = (n + 31) // 32
    padded = np.zeros((B, words * 32), dtype=np.uint64)
    padded[:, :n] = m
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (padded.reshape(B, words, 32) * weights).sum(-1).astype(np.uint32)


def unpack_bits(words: np.ndarray, nbits: int) -> np.ndarray:
    w = np.asarray(words, dtype=np.uint32)
    bits = (w[:, :, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    return bits.reshape(w.shape[0], -1)[:, :nbits].astype(f32)


def bit_errors_packed(a: np.ndarray, b: np.ndarray) -> int:
    """XOR + popcount over packed words (padding bits are zero on both sides)."""
    x = np.bitwise_xor(np.asarray(a, dtype=np.uint32), np.asarray(b, dtype=np.uint32))
    return int(np.unpackbits(x.view(np.uint8)).sum())
