"""The kernels convert unorm texels with q0 = u * fl(1/max); q = fma(fma(-q0, max, u), fl(1/max), q0)
instead of an IEEE division (libplacebo_b200/csrc/kernels_common.cuh). Exhaustive proof, in exact
rational arithmetic, that this equals the correctly rounded u / max (what the oracle and texture
units compute) for every 8-bit and 16-bit code value."""
from fractions import Fraction

import numpy as np
import pytest


def rnd32(fr):
    """Fraction -> nearest float32, ties to even (exact)."""
    if fr == 0:
        return np.float32(0)
    f = np.float32(float(fr))
    best = None
    for c in (f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))):
        d = abs(Fraction(float(c)) - fr)
        even = (int(np.float32(c).view(np.uint32)) & 1) == 0
        if best is None or d < best[0] or (d == best[0] and even):
            best = (d, np.float32(c))
    return best[1]


@pytest.mark.parametrize("maxv", [255, 65535])
def test_fma_sequence_equals_ieee_division(maxv):
    c = Fraction(float(rnd32(Fraction(1, maxv))))
    for u in range(maxv + 1):
        q0 = Fraction(float(rnd32(u * c)))
        r = Fraction(float(rnd32(u - q0 * maxv)))
        q = rnd32(r * c + q0)
        assert q == rnd32(Fraction(u, maxv)), u
        assert q == np.float32(u) / np.float32(maxv)
