"""Philox4x32-10 and the per-attempt random-slot layout.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws from ``numpy.random.Generator(PCG64)`` (``src/quansino/mc/driver.py:77-78``) through
``context.rng`` -- an attribute that is only ever *stored and called* (``src/quansino/mc/contexts.py:49-52``),
so any object with ``random / uniform / choice / standard_normal`` can be injected.  The B200 build does
not emulate PCG64; it defines a counter-based stream that the kernels, this oracle and (through
``InjectedStream``) the genuine quansino sources all consume, so that "same inputs, same random stream"
is true by construction.

Layout (DESIGN.md section "Random stream"):

    key      = (seed_lo, seed_hi)
    counter  = (attempt_lo, attempt_hi, replica_id, block)
    one Philox call -> 4 x u32 (w0..w3) -> two doubles
        d0 = ((w0 >> 5) * 2**26 + (w1 >> 6)) * 2**-53        d1 = same from (w2, w3)

    block 0: d0 = u_select   (move selection, ``mc/core.py:344-349``)
             d1 = u_label    (label index = floor(u * n_unique), ``moves/displacement.py:164``)
    block 1: d0, d1 = operation draws 0, 1   (Ball: r, phi -- ``operations/displacement.py:146-147``)
    block 2: d0 = operation draw 2           (Ball: cos(theta) -- ``:148``)
             d1 = u_accept                   (``mc/criteria.py:108``)
    block 3: d0 = u_coin  (insert / delete, ``moves/exchange.py:148``)   d1 = u_swap (parallel tempering)
    block 4 + j / 2 (of the FIRST attempt of a step): forced-move placement draw j (``mc/core.py:335-337``)
    block 32 + 2 (c - 1), 33 + 2 (c - 1): sub-move c >= 1 of a CompositeDisplacementMove (``moves/displacement.py:392-431``;
             c counts the sub-moves that draw, the first one uses the standard slots above; from c = 17 on the pair of blocks is
             0x8000 + 2 (c - 17), + 1 -- ``submove_block``):
             (u_label, operation draw 0) and (operation draw 1, operation draw 2); operation draws beyond the third of a
             (sub-)move -- the further operations of a ``CompositeOperation`` (``Ball(0.1) + Box(0.05)``), the CellMove
             closing a hybrid ``CompositeMove`` after displacements drew -- are numbered e = 0, 1, ... through the attempt
             and live in block 64 + e / 2, half e & 1
    block 16 + p: Box-Muller pair p of the 3N momentum normals (``utils/dynamics.py:31``), row-major:
             normal m = 3*atom + axis lives in pair p = m // 2; z_even = r cos(t), z_odd = r sin(t),
             r = sqrt(-2 ln(1 - d0)), t = 2 pi d1.
"""

from __future__ import annotations

import numpy as np

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

BLOCK_SELECT_LABEL = 0
BLOCK_OP01 = 1
BLOCK_OP2_ACCEPT = 2
BLOCK_COIN_SWAP = 3
BLOCK_FORCED = 4
BLOCK_NORMALS = 16
BLOCK_COMPOSITE = 32
BLOCK_EXTRA = 64
BLOCK_COMPOSITE_FAR = 0x8000  # sub-moves 17 and up (blocks 32 .. 63 are full, the extra draws start at 64)


def submove_block(c: int) -> int:
    """Block of (u_label, op draw 0) of drawing sub-move c >= 1 of a composite move; (op draw 1, op draw 2): the next block."""
    return BLOCK_COMPOSITE + 2 * (c - 1) if c <= 16 else BLOCK_COMPOSITE_FAR + 2 * (c - 17)


_MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Scalar Philox4x32-10 (Salmon et al., SC'11).  ``counter`` 4 x u32, ``key`` 2 x u32 -> 4 x u32."""
    c0, c1, c2, c3 = (int(c) & _MASK for c in counter)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + PHILOX_W0) & _MASK
        k1 = (k1 + PHILOX_W1) & _MASK
    return c0, c1, c2, c3


def words_to_doubles(w):
    """4 x u32 -> two doubles in [0, 1) with 53 random bits each."""
    d0 = ((w[0] >> 5) * 67108864 + (w[1] >> 6)) / 9007199254740992.0
    d1 = ((w[2] >> 5) * 67108864 + (w[3] >> 6)) / 9007199254740992.0
    return d0, d1


def uniform_pair(seed: int, attempt: int, replica: int, block: int):
    """The two doubles of (attempt, replica, block) under ``seed`` (64-bit)."""
    key = (seed & _MASK, (seed >> 32) & _MASK)
    ctr = (attempt & _MASK, (attempt >> 32) & _MASK, replica & _MASK, block & _MASK)
    return words_to_doubles(philox4x32_10(ctr, key))


def normal(seed: int, attempt: int, replica: int, m: int) -> float:
    """Standard normal number ``m`` (row-major index into the (N, 3) momentum draw) of an attempt."""
    d0, d1 = uniform_pair(seed, attempt, replica, BLOCK_NORMALS + m // 2)
    r = np.sqrt(-2.0 * np.log(1.0 - d0))
    t = 2.0 * np.pi * d1
    return float(r * np.cos(t)) if m % 2 == 0 else float(r * np.sin(t))


def fbmc_draw(seed: int, step: int, replica: int, call: int, p: int) -> float:
    """Element ``p`` of draw call ``call`` of ForceBias step ``step`` (include/quansino_b200.h, qmc_fbmc_step)."""
    return uniform_pair(seed, step, replica, ((call + 1) << 16) | (p >> 1))[p & 1]


class InjectedStream:
    """A ``numpy.random.Generator`` stand-in that hands the slot layout above to the GENUINE quansino code.

    The driver advances ``attempt`` once per cycle by consuming ``choice(names, p=...)`` -- the first draw of
    every cycle in ``MonteCarlo.yield_moves`` (``mc/core.py:349``).  Every other call maps onto a fixed slot
    of the current attempt, in the order the reference makes them:

    * ``choice(a, p=p)``            -> new attempt; u_select; ``searchsorted(cumsum(p), u, 'right')``
    * ``choice(a)`` (no ``p``)      -> u_label; ``a[floor(u * len(a))]``.  The c-th such call of an attempt (c >= 1, only
                                       ``CompositeDisplacementMove`` makes more than one) starts sub-move c: its label and
                                       operation draws come from blocks 32 + 2 (c - 1) and 33 + 2 (c - 1)
    * ``choice(arange(M), size=K, replace=False)`` -> the forced-move placement of a step (``mc/core.py:335-337``): nothing for
                                       K = 0, else a partial Fisher-Yates shuffle fed from blocks 4 + j / 2 of the step's first
                                       attempt (include/quansino_b200.h, "Forced moves")
    * ``mark_cycle()``              -> called by the driving loop once per yielded cycle: a cycle that made no selection draw
                                       (a forced one) still advances the attempt index
    * ``uniform(lo, hi, size)``     -> operation draws, in call order: ``lo + (hi - lo) * u``
    * ``random()``                  -> u_coin when it is the FIRST draw after the selection (only
                                       ``ExchangeMove.__call__`` does that, ``moves/exchange.py:148``);
                                       otherwise u_accept (every criteria draws last)
    * ``standard_normal(shape)``    -> Box-Muller normals of the attempt, row-major
    """

    def __init__(self, seed: int, replica: int = 0, attempt_base: int = 0, mode: str = "mc"):
        self.mode = mode  # "mc": the per-attempt slots above; "fbmc": ForceBias (see fbmc_draw)
        self._fb_call = 0
        self.seed = int(seed)
        self.replica = int(replica)
        self.attempt = int(attempt_base) - 1
        self._op_draws = 0
        self._calls = 0
        self._labels = 0  # label draws of this attempt
        self._extra = 0  # operation draws of this attempt beyond the third of a (sub-)move
        self._select_pending = False  # a selection draw opened the attempt that the next mark_cycle() belongs to
        self.log: list[tuple[str, float]] = []

    # -- slots ---------------------------------------------------------------------------------
    def _pair(self, block):
        return uniform_pair(self.seed, self.attempt, self.replica, block)

    def _next_attempt(self):
        self.attempt += 1
        self._op_draws = 0
        self._calls = 0
        self._labels = 0
        self._extra = 0

    def _op_draw(self) -> float:
        k = self._op_draws
        self._op_draws += 1
        self._calls += 1
        if k >= 3:  # beyond the three operation draws of a (sub-)move: the further operations of a CompositeOperation
            e = self._extra  # (operations/composite.py:58-73) and the CellMove that closes a hybrid CompositeMove
            self._extra += 1
            return self._pair(BLOCK_EXTRA + e // 2)[e & 1]
        if self._labels > 1:  # sub-move c = _labels - 1 >= 1 of a composite displacement move
            c = self._labels - 1
            if k == 0:
                return self._pair(submove_block(c))[1]
            if k in (1, 2):
                return self._pair(submove_block(c) + 1)[k - 1]
            raise RuntimeError("more than three operation draws in one sub-move: not a supported operation")
        if k == 0:
            return self._pair(BLOCK_OP01)[0]
        if k == 1:
            return self._pair(BLOCK_OP01)[1]
        if k == 2:
            return self._pair(BLOCK_OP2_ACCEPT)[0]
        raise RuntimeError("more than three operation draws in one attempt: not a supported operation")

    # -- numpy.random.Generator surface used by quansino -------------------------------------------
    def choice(self, a, size=None, replace=True, p=None):
        a = np.asarray(a)
        if size is not None:
            if size == 0 or (hasattr(size, "__len__") and len(size) == 0):
                return a[:0]
            if replace or p is not None:
                raise NotImplementedError("only choice(arange(M), size=K, replace=False) is part of the injected stream")
            perm = list(range(len(a)))  # partial Fisher-Yates with the step's first attempt (= the next one)
            for j in range(int(size)):
                u = uniform_pair(self.seed, self.attempt + 1, self.replica, BLOCK_FORCED + j // 2)[j & 1]
                t = j + int(u * (len(a) - j))
                perm[j], perm[t] = perm[t], perm[j]
            return a[perm[: int(size)]]
        if p is not None:
            self._next_attempt()
            self._select_pending = True
            u = self._pair(BLOCK_SELECT_LABEL)[0]
            self.log.append(("select", u))
            cdf = np.cumsum(np.asarray(p, dtype=float))
            cdf /= cdf[-1]
            return a[int(np.searchsorted(cdf, u, side="right"))]
        self._calls += 1
        self._labels += 1
        self._op_draws = 0
        c = self._labels - 1
        u = self._pair(BLOCK_SELECT_LABEL)[1] if c == 0 else self._pair(submove_block(c))[0]
        self.log.append(("label", u))
        return a[int(u * len(a))]

    def mark_cycle(self):
        """Once per cycle the driver yields, before the move runs."""
        if not self._select_pending:
            self._next_attempt()  # a forced cycle: no selection draw opened this attempt
        self._select_pending = False

    # -- ForceBias (mc/fbmc.py:224-241): step s draws zeta = uniform(-1, 1, (N, 3)) and random((N, 3)), then again for the
    # coordinates whose trial was refused, until none is left.  Call c of step s (0, 1, 2, ...; zeta calls even, random calls
    # odd), element p of that call -> Philox(seed; attempt = s, replica, block = ((c + 1) << 16) | (p >> 1)), half p & 1.
    def _fbmc(self, size, new_step_if_2d):
        shape = (size,) if isinstance(size, (int, np.integer)) else tuple(size)
        if new_step_if_2d and len(shape) == 2:
            self._next_attempt()
            self._fb_call = 0
        c = self._fb_call
        self._fb_call += 1
        n = int(np.prod(shape))
        out = np.array([fbmc_draw(self.seed, self.attempt, self.replica, c, p) for p in range(n)])
        return out.reshape(shape)

    def uniform(self, low=0.0, high=1.0, size=None):
        if self.mode == "fbmc":
            return low + (high - low) * self._fbmc(size, True)
        if size is None:
            return low + (high - low) * self._op_draw()
        n = int(np.prod(size))
        out = np.array([low + (high - low) * self._op_draw() for _ in range(n)])
        return out.reshape(size)

    def random(self, size=None):
        if self.mode == "fbmc":
            return self._fbmc(size, False)
        first = self._calls == 0
        self._calls += 1
        if first:
            u = self._pair(BLOCK_COIN_SWAP)[0]
            self.log.append(("coin", u))
            return u
        u = self._pair(BLOCK_OP2_ACCEPT)[1]
        self.log.append(("accept", u))
        return u

    def standard_normal(self, size=None):
        self._calls += 1
        if size is None:
            return normal(self.seed, self.attempt, self.replica, 0)
        n = int(np.prod(size))
        out = np.array([normal(self.seed, self.attempt, self.replica, m) for m in range(n)])
        return out.reshape(size)

    @property
    def bit_generator(self):  # Driver.to_dict touches rng.bit_generator.state; not used in traces
        raise AttributeError("InjectedStream has no bit generator")
