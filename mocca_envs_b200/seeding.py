"""gym<=0.21 seeding shim (``gym.utils.seeding.np_random``; used by the reference at env_base.py:164-166).

``np_random(seed)`` hashes the integer seed with sha512, keeps 8 bytes, and feeds the resulting uint32 words to
``numpy.random.RandomState.seed`` (MT19937 ``init_by_array``).  The device kernels continue exactly that
stream (csrc/mb_env.cuh: mt_fill / mt_double), so reset states and targets are bit-exact with the reference's
draws (after rounding to the float32 state representation).
"""
from __future__ import annotations

import hashlib
import os
import struct

import numpy as np


def create_seed(a=None, max_bytes: int = 8) -> int:
    if a is None:
        return int.from_bytes(os.urandom(max_bytes), "little")
    if not (isinstance(a, (int, np.integer)) and a >= 0):
        raise ValueError("Seed must be a non-negative integer or omitted, not {}".format(a))
    return int(a) % 2 ** (8 * max_bytes)


def seed_words(seed: int):
    """The uint32 key list RandomState is seeded with for integer ``seed``."""
    h = hashlib.sha512(str(create_seed(seed)).encode("utf8")).digest()[:8]
    h += b"\0" * 4
    big = sum(2 ** (32 * i) * v for i, v in enumerate(struct.unpack("3I", h)))
    if big == 0:
        return [0]
    words = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        words.append(mod)
    return words


def mt_state_rows(seeds) -> np.ndarray:
    """[len(seeds), 625] uint32: MT19937 key + position of RandomState(seed_words(s)) for every s."""
    out = np.empty((len(seeds), 625), dtype=np.uint32)
    for i, s in enumerate(seeds):
        st = np.random.RandomState(seed_words(int(s))).get_state()
        out[i, :624] = st[1]
        out[i, 624] = st[2]
    return out
