"""NumPy/Python twin of the device-side plan sampler (csrc/dks_sampler.cuh) -- test infrastructure.

``PhiloxPlanStream`` looks like the legacy NumPy stream to the oracle's ``build_plan`` (``choice`` once, then one
``permutation(M)`` per draw) but produces the numbers the CUDA sampler produces for (seed, global row): Philox4x32-10
with key = seed and counter = (draw t, row lo, row hi, block).  Feeding it to the oracle therefore yields, through the
oracle's restatement of upstream's sequential loop, the plan the device must have drawn.
"""
import numpy as np

M32 = 0xFFFFFFFF


def philox4x32_10(key, ctr):
    k0, k1 = key
    c0, c1, c2, c3 = ctr
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        hi0, lo0 = p0 >> 32, p0 & M32
        hi1, lo1 = p1 >> 32, p1 & M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class _LazySubset:
    """What ``permutation(M)`` returns: only ``[:size]`` is ever taken from it, and the size-``size`` subset is drawn
    the way the device draws it (Floyd's algorithm: for j = M-size .. M-1 pick t uniform in [0, j]; take j if t is
    already in the set, else t)."""

    def __init__(self, stream, t, M):
        self.stream, self.t, self.M = stream, t, M

    def __getitem__(self, sl):
        assert isinstance(sl, slice) and sl.start is None and sl.step is None
        size, M, t = sl.stop, self.M, self.t
        rnd = list(self.stream._block(t, 0))
        have, block = 1, 0
        chosen, members = 0, []
        for j in range(M - size, M):
            if have == 4:
                block += 1
                rnd = list(self.stream._block(t, block))
                have = 0
            r32 = rnd[have]
            have += 1
            pick = (r32 * (j + 1)) >> 32
            bit = j if (chosen >> pick) & 1 else pick
            chosen |= 1 << bit
            members.append(bit)
        return np.array(members, dtype=np.int64)


class PhiloxPlanStream:
    def __init__(self, seed, row):
        self.key = (seed & M32, (seed >> 32) & M32)
        self.row = (row & M32, (row >> 32) & M32)
        self.t = 0

    def _block(self, t, block):
        return philox4x32_10(self.key, (t, self.row[0], self.row[1], block))

    def choice(self, n, size, p=None):
        cdf = np.cumsum(np.asarray(p, dtype=np.float64))
        cdf[-1] = 1.0
        u = np.array([(self._block(t, 0)[0] + 0.5) * 2.0 ** -32 for t in range(size)])
        return np.minimum(np.searchsorted(cdf, u, side="right"), n - 1)

    def permutation(self, M):
        t = self.t
        self.t += 1
        return _LazySubset(self, t, M)
