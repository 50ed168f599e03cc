"""Pure-Python restatement of numpy's ``Generator(PCG64(SeedSequence(seed)))``.

Oracle only (see oracle/__init__.py).  Follows the algorithm the reference reaches
through ``gymnasium/utils/seeding.py:39-41`` (``np.random.SeedSequence`` ->
``np.random.PCG64`` -> ``np.random.Generator``); numpy itself is a third-party
dependency of the reference (numpy 2.3.5 in this image), restated from its
published algorithm and pinned bit-exactly against numpy
in tests/test_oracle_rng.py.
"""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
M128 = (1 << 128) - 1

INIT_A, MULT_A = 0x43B0D7E5, 0x931E8875
INIT_B, MULT_B = 0x8B51F9DD, 0x58F38DED
MIX_L, MIX_R = 0xCA01F9DD, 0x4973F715
XSHIFT = 16
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


def seed_sequence_state(seed: int, n_words: int = 8) -> list[int]:
    """``SeedSequence(seed).generate_state(n_words, uint32)`` (pool_size=4)."""
    if seed < 0:
        raise ValueError("seed must be non-negative")
    words = []
    s = seed
    while s > 0:
        words.append(s & M32)
        s >>= 32
    if not words:
        words = [0]

    hc = [INIT_A]

    def hashmix(v: int) -> int:
        v = (v ^ hc[0]) & M32
        hc[0] = (hc[0] * MULT_A) & M32
        v = (v * hc[0]) & M32
        v ^= v >> XSHIFT
        return v

    def mix(x: int, y: int) -> int:
        r = ((MIX_L * x) & M32) - ((MIX_R * y) & M32)
        r &= M32
        r ^= r >> XSHIFT
        return r

    pool = [0] * 4
    for i in range(4):
        pool[i] = hashmix(words[i] if i < len(words) else 0)
    for s_ in range(4):
        for d in range(4):
            if s_ != d:
                pool[d] = mix(pool[d], hashmix(pool[s_]))
    for i in range(4, len(words)):
        for d in range(4):
            pool[d] = mix(pool[d], hashmix(words[i]))

    out = []
    hb = INIT_B
    for i in range(n_words):
        v = pool[i % 4]
        v ^= hb
        hb = (hb * MULT_B) & M32
        v = (v * hb) & M32
        v ^= v >> XSHIFT
        out.append(v)
    return out


ZIGGURAT_NOR_R = 3.6541528853610087963519472518
ZIGGURAT_NOR_INV_R = 0.27366123732975827203338247596
_ZIG = None


def _ziggurat_tables():
    """(wi, ki, fi) as Python lists: wi_double / ki_double / fi_double of numpy's ziggurat_constants.h."""
    global _ZIG
    if _ZIG is None:
        import os

        import numpy as np

        t = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ziggurat_normal.npz"))
        _ZIG = ([float(v) for v in t["wi"]], [int(v) for v in t["ki"]], [float(v) for v in t["fi"]])
    return _ZIG


class PCG64:
    """numpy's PCG64 (XSL-RR 128/64) bit generator seeded from a SeedSequence."""

    def __init__(self, seed: int):
        w = seed_sequence_state(seed, 8)
        u = [w[2 * k] | (w[2 * k + 1] << 32) for k in range(4)]
        initstate = (u[0] << 64) | u[1]
        initseq = (u[2] << 64) | u[3]
        self.inc = ((initseq << 1) | 1) & M128
        self.state = 0
        self.has_uint32, self.uinteger = False, 0  # pcg64_next32's one-word buffer (numpy/random/src/pcg64/pcg64.h)
        self._step()
        self.state = (self.state + initstate) & M128
        self._step()

    def _step(self) -> None:
        self.state = (self.state * PCG_MULT + self.inc) & M128

    def next_uint64(self) -> int:
        self._step()
        hi = self.state >> 64
        lo = self.state & M64
        x = hi ^ lo
        rot = hi >> 58
        return ((x >> rot) | (x << ((-rot) & 63))) & M64

    def next_double(self) -> float:
        """``Generator.random()``: 53 random bits scaled by 2**-53."""
        return (self.next_uint64() >> 11) * (1.0 / 9007199254740992.0)

    def uniform(self, low: float, high: float) -> float:
        """``Generator.uniform(low, high)`` = ``low + (high-low)*next_double`` (two roundings)."""
        return low + (high - low) * self.next_double()

    def next_uint32(self) -> int:
        """``pcg64_next32``: the low half of a fresh 64-bit draw now, the high half on the next call."""
        if self.has_uint32:
            self.has_uint32 = False
            return self.uinteger
        x = self.next_uint64()
        self.has_uint32, self.uinteger = True, x >> 32
        return x & M32

    def standard_normal(self) -> float:
        """``Generator.standard_normal()`` / ``normal()``: the 256-strip ziggurat of numpy/random/src/distributions/
        distributions.c (random_standard_normal) over this bit generator's 64-bit words -- 8 bits pick the strip, 1 bit the
        sign, 52 bits the abscissa; 99.3 % of the draws return on the first comparison, the wedges take one more uniform, the
        tail two per trial.  Tables: oracle/data/ziggurat_normal.npz (numpy's own doubles, see extract_ziggurat.py)."""
        import math

        wi, ki, fi = _ziggurat_tables()
        while True:
            r = self.next_uint64()
            idx = r & 0xFF
            r >>= 8
            sign = r & 1
            rabs = (r >> 1) & 0x000FFFFFFFFFFFFF
            x = rabs * wi[idx]
            if sign:
                x = -x
            if rabs < ki[idx]:
                return x
            if idx == 0:
                while True:
                    xx = -ZIGGURAT_NOR_INV_R * math.log1p(-self.next_double())
                    yy = -math.log1p(-self.next_double())
                    if yy + yy > xx * xx:
                        return -(ZIGGURAT_NOR_R + xx) if (rabs >> 8) & 1 else ZIGGURAT_NOR_R + xx
            elif (fi[idx - 1] - fi[idx]) * self.next_double() + fi[idx] < math.exp(-0.5 * x * x):
                return x

    def bounded_uint32(self, n_excl: int) -> int:
        """Uniform integer in ``[0, n_excl)`` for ``1 <= n_excl <= 2**32``: Lemire's multiply-and-reject on buffered
        32-bit words (``buffered_bounded_lemire_uint32``, numpy/random/src/distributions/distributions.c).  This is
        what ``Generator.choice(seq)`` (no ``p``, scalar) and ``Generator.integers(0, n)`` reduce to for small ranges."""
        rng = n_excl - 1
        if rng == 0:
            return 0
        m = self.next_uint32() * n_excl
        leftover = m & M32
        if leftover < n_excl:
            threshold = (M32 - rng) % n_excl
            while leftover < threshold:
                m = self.next_uint32() * n_excl
                leftover = m & M32
        return m >> 32
