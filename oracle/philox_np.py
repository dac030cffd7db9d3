"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy restatement of the counter-based generator the device compositor uses for ``reset_mode="random"`` in
throughput mode (``transflow_b200/csrc/compositor.cu``: ``philox4x32_7`` / ``philox_uniform53``).  The reference
draws ``numpy.random.random((H, W))`` per frame (``transflow/compositor/layers/reference.py:59``); which numbers
are drawn is not part of its contract, the compare ``r < factor * reset_mask`` in float64 is (``:60-67``).  Feeding
the oracle layer the numbers THIS function returns for (seed, frame) therefore pins the device fast path
(``k_moveref_fast<RESET_RANDOM>``, which only runs with device draws) bit-exactly against the reference semantics.

Algorithm: Philox4x32 with 7 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
multipliers 0xD2511F53 / 0xCD9E8D57 and Weyl constants 0x9E3779B9 / 0xBB67AE85 of Random123).  Counter =
(pixel >> 1, frame low, frame high, 0x7f4a7c15), key = the 64-bit seed; one evaluation yields 128 bits = two
53-bit uniforms built like NumPy's ``random_sample``: ``((a >> 5) * 2**26 + (b >> 6)) / 2**53``; the even pixel of a
pair takes words (0, 1), the odd pixel words (2, 3).
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(counter, key, rounds: int = 7):
    """Philox4x32-``rounds``: ``counter`` = four uint32 arrays (or ints), ``key`` = two 32-bit ints -> four uint32
    arrays.  Pinned by Random123's known-answer vectors (tests/test_oracle.py)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in counter)
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _M0 * c0            # 32 x 32 -> 64 bit products (no overflow in uint64)
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def philox4x32_7(seed: int, frame: int, ctr: np.ndarray):
    """The device's block for every counter value in ``ctr``: counter (ctr, frame lo, frame hi, 0x7f4a7c15), key =
    the 64-bit seed."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    return philox4x32((ctr, np.full_like(ctr, frame & 0xFFFFFFFF), np.full_like(ctr, (frame >> 32) & 0xFFFFFFFF),
                       np.full_like(ctr, 0x7F4A7C15)), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF), 7)


def reset_draws(seed: int, frame: int, height: int, width: int) -> np.ndarray:
    """The float64 (H, W) field of uniforms in [0, 1) the device layer compares against for its ``frame``-th update
    (0-based) with Philox key ``seed`` (``Layer.rng_seed``)."""
    n = height * width
    pixel = np.arange(n, dtype=np.uint32)
    r0, r1, r2, r3 = philox4x32_7(int(seed), int(frame), pixel >> np.uint32(1))
    odd = (pixel & np.uint32(1)).astype(bool)
    a = np.where(odd, r2, r0)
    b = np.where(odd, r3, r1)
    u = ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) \
        * (1.0 / 9007199254740992.0)
    return u.reshape(height, width)
