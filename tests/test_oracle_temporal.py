"""CPU-only properties of the oracle's temporal functions (the checkers of test_gpu_deinterlace.py
and test_gpu_mix.py), so that the checkers themselves are pinned to what the algorithms
guarantee (shaders/deinterlacing.c, renderer.c:3661-3757)."""
import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T

ALGOS = {"weave": 0, "bob": 1, "yadif": 2, "bwdif": 3}


@pytest.mark.parametrize("algo", list(ALGOS))
@pytest.mark.parametrize("field", [1, 2])
def test_deinterlace_keeps_the_shown_field_and_flat_content(algo, field):
    orc = T.oracle()
    rng = np.random.default_rng(field)
    cur, prev, nxt = (rng.random((24, 31, 2), dtype=np.float32) for _ in range(3))
    out = orc.deinterlace(cur, prev, nxt, field=field, first_field=1, algo=ALGOS[algo])
    keep = slice(0, None, 2) if field == 1 else slice(1, None, 2)
    assert np.array_equal(out[keep, :, :2], cur[keep, :, :2])      # the field's own lines: untouched
    assert np.all(out[..., 2] == 0) and np.all(out[..., 3] == 1)   # vec4(0, 0, 0, 1) elsewhere
    # content that does not depend on y or t is reproduced exactly by every algorithm
    flat = np.repeat(rng.random((1, 31, 2), dtype=np.float32), 24, axis=0)
    out = orc.deinterlace(flat, flat, flat, field=field, first_field=1, algo=ALGOS[algo])
    assert np.abs(out[..., :2] - flat).max() < 1e-6


def test_bob_repeats_the_neighbouring_line_and_weave_is_identity():
    orc = T.oracle()
    cur = np.random.default_rng(0).random((16, 9, 1), dtype=np.float32)
    assert np.array_equal(orc.deinterlace(cur, algo=0, field=1)[..., :1], cur)
    bob = orc.deinterlace(cur, algo=1, field=1)[..., 0]
    assert np.array_equal(bob[1::2], cur[0::2, :, 0])     # top field shown: odd lines = line above
    bob = orc.deinterlace(cur, algo=1, field=2)[..., 0]
    assert np.array_equal(bob[0:-1:2], cur[1::2, :, 0])   # bottom field: even lines = line below


def test_temporal_algorithms_on_a_static_vertical_ramp():
    """A frame that is linear in y and constant in time: the missing lines are the mean of their
    neighbours (yadif: spatial prediction clamped around the temporal one; bwdif: 4-tap spline
    whose coefficients sum to one)."""
    orc = T.oracle()
    y = np.arange(32, dtype=np.float32)[:, None, None] / 64 + np.zeros((32, 12, 1), np.float32)
    for algo in (2, 3):
        out = orc.deinterlace(y, y, y, field=1, first_field=1, algo=algo)[..., 0]
        assert np.abs(out[4:-4] - y[4:-4, :, 0]).max() < 1e-6


def test_mix_weights():
    t0 = T.t0()
    # oversampling: visible fractions of the vsync interval, summing to one when it is covered
    ws = P.mix_weights([-0.4, 0.6, 1.6], 1.0, "oversample", t0)
    assert [i for i, _ in ws] == [0, 1] and abs(sum(w for _, w in ws) - 1.0) < 1e-6
    assert abs(ws[0][1] - 0.6) < 1e-6
    # no mixer: the frame nearest to the vsync, alone
    assert P.mix_weights([-0.4, 0.6, 1.6], 1.0, "", t0) == [(0, 1.0)]
    assert P.mix_weights([-0.7, 0.3, 1.3], 1.0, "", t0) == [(1, 1.0)]
    # a filter: sampled at the timestamps, frames outside its radius dropped
    ws = P.mix_weights([-0.75, 0.25, 1.25], 1.0, "hermite", t0)
    assert [i for i, _ in ws] == [0, 1] and ws[1][1] > ws[0][1] > 0


def test_mix_accumulate_is_a_fused_multiply_add():
    orc = T.oracle()
    rng = np.random.default_rng(2)
    acc, c = (rng.random((8, 8, 4), dtype=np.float32) for _ in range(2))
    w = np.float32(0.3)
    got = orc.mix_accumulate(acc, c, float(w))
    want = (w.astype(np.float64) * c.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
    assert np.array_equal(got, want)   # one rounding (fma), as `mix_color += w * color` fuses
