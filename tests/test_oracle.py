"""CPU tests of the oracle (oracle/filters_oracle.c) against independent numpy restatements and
against the golden hashes minted from the reference's own kernels on a B200
(tests/golden/reference_hashes.json, written by tools/mint_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_hashes.json")


def np_weights(radius, sigma):
    """image_filters.cu:25-39 in numpy float32 (expf taken from the oracle's libm via float32 exp)."""
    s = np.float32(sigma)
    x = np.arange(-radius, radius + 1, dtype=np.float32)
    return x, s


def np_blur_1d(img, w, axis):
    """float64 separable pass with clamp-to-edge; returns the unrounded sums."""
    r = (len(w) - 1) // 2
    pad = [(0, 0)] * 3
    pad[axis] = (r, r)
    p = np.pad(img.astype(np.float64), pad, mode="edge")
    out = np.zeros(img.shape, dtype=np.float64)
    n = img.shape[axis]
    for i in range(2 * r + 1):
        sl = [slice(None)] * 3
        sl[axis] = slice(i, i + n)
        out += p[tuple(sl)] * float(w[i])
    return out


def np_box(img, r):
    """integer box blur, (2S+k)//(2k), u8 intermediate."""
    k = 2 * r + 1
    cur = img
    for axis in (1, 0):
        pad = [(0, 0)] * 3
        pad[axis] = (r, r)
        p = np.pad(cur.astype(np.int64), pad, mode="edge")
        cs = np.cumsum(p, axis=axis)
        zero = np.zeros_like(np.take(cs, [0], axis=axis))
        cs = np.concatenate([zero, cs], axis=axis)
        n = cur.shape[axis]
        hi = np.take(cs, np.arange(k, k + n), axis=axis)
        lo = np.take(cs, np.arange(0, n), axis=axis)
        cur = ((2 * (hi - lo) + k) // (2 * k)).astype(np.uint8)
    return cur


def np_sobel_gray_int(gray):
    """Exact integer Sobel on a 2-D integer gray image; borders 0 (image_filters.cu:1164-1232)."""
    g = gray.astype(np.int64)
    h, w = g.shape
    out = np.zeros((h, w), dtype=np.uint8)
    if h < 3 or w < 3:
        return out
    tl, tc, tr = g[:-2, :-2], g[:-2, 1:-1], g[:-2, 2:]
    ml, mr = g[1:-1, :-2], g[1:-1, 2:]
    bl, bc, br = g[2:, :-2], g[2:, 1:-1], g[2:, 2:]
    gx = -tl + tr - 2 * ml + 2 * mr - bl + br
    gy = -tl - 2 * tc - tr + bl + 2 * bc + br
    m2 = (gx * gx + gy * gy).astype(np.float32)      # < 2^24: exact in float32
    m = np.minimum(np.sqrt(m2), np.float32(255.0))   # numpy sqrt is correctly rounded
    out[1:-1, 1:-1] = (m + np.float32(0.5)).astype(np.uint8)
    return out


SHAPES = [(1, 1), (1, 9), (9, 1), (2, 2), (3, 3), (17, 33), (40, 5), (64, 48)]


@pytest.mark.parametrize("radius,sigma", [(0, 1.0), (1, 0.5), (3, 2.0), (15, 5.0), (31, 10.0)])
def test_weights_normalised_and_symmetric(radius, sigma):
    w = O.gaussian_weights(radius, sigma)
    assert w.dtype == np.float32 and len(w) == 2 * radius + 1
    assert abs(float(w.sum()) - 1.0) < 1e-5
    assert np.array_equal(w, w[::-1])
    ref = np.exp(-(np.arange(-radius, radius + 1) ** 2) / (2.0 * sigma * sigma))
    assert np.allclose(w, ref / ref.sum(), rtol=2e-6, atol=1e-9)


def test_box_float_rule_equals_integer_rule_exhaustively():
    """(uchar)(S*(1.0f/k)+0.5f), fused or not, == (2S+k)//(2k) for every S, every odd k <= 63."""
    for k in range(1, 64, 2):
        s = np.arange(0, 255 * k + 1, dtype=np.float32)
        inv = np.float32(1.0) / np.float32(k)
        unfused = (s * inv + np.float32(0.5)).astype(np.uint8)
        want = ((2 * np.arange(0, 255 * k + 1) + k) // (2 * k)).astype(np.uint8)
        assert np.array_equal(unfused, want), k
    for k in (1, 3, 7, 31, 63):                      # the fused form through the oracle's fmaf
        for s in list(range(0, 255 * k + 1, 97)) + [255 * k]:
            assert O.box_round_float(s, k, True) == (2 * s + k) // (2 * k)


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("hw", SHAPES)
@pytest.mark.parametrize("radius", [0, 1, 3, 8])
def test_box_matches_numpy_integer(hw, c, radius):
    img = synth.uniform(hw[0], hw[1], c, seed=radius + c)
    got = O.box_blur(img, radius)
    assert np.array_equal(got, O.box_blur(img, radius, integer=True))
    assert np.array_equal(got, np_box(img, radius))


@pytest.mark.parametrize("c", [1, 3, 4])
@pytest.mark.parametrize("hw", SHAPES)
@pytest.mark.parametrize("radius,sigma", [(1, 0.8), (3, 2.0), (7, 3.0)])
def test_gaussian_matches_float64_within_one_lsb(hw, c, radius, sigma):
    img = synth.uniform(hw[0], hw[1], c, seed=7 * radius + c)
    got = O.gaussian_blur(img, sigma, radius)
    w = O.gaussian_weights(radius, sigma)
    tmp = np.floor(np_blur_1d(img, w, 1) + 0.5).astype(np.uint8)
    ref = np.floor(np_blur_1d(tmp, w, 0) + 0.5)
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 1                                   # tolerance: 1 LSB (float32 vs float64 sums)
    assert (d > 0).mean() < 0.02


def test_gaussian_constant_and_identity():
    for v in (0, 1, 127, 255):
        img = synth.constant(13, 11, 3, v)
        assert np.array_equal(O.gaussian_blur(img, 2.0, 3), img)
        assert np.array_equal(O.box_blur(img, 4), img)
    img = synth.uniform(9, 9, 4)
    assert np.array_equal(O.gaussian_blur(img, 1.0, 0), img)
    assert np.array_equal(O.box_blur(img, 0), img)


@pytest.mark.parametrize("hw", SHAPES)
def test_sobel_gray_is_exact_integer_sobel(hw):
    img = synth.uniform(hw[0], hw[1], 1, seed=3)
    want = np_sobel_gray_int(img[:, :, 0])
    for level in (1, 2):
        assert np.array_equal(O.sobel(img, level)[:, :, 0], want)


@pytest.mark.parametrize("c", [3, 4])
@pytest.mark.parametrize("kind", ["uniform", "smooth"])
def test_sobel_colour_levels(c, kind):
    img = synth.KINDS[kind](37, 53, c, seed=5)
    l1, l2 = O.sobel(img, 1), O.sobel(img, 2)
    # borders are zero in every channel, the edge value is replicated into every channel
    for out in (l1, l2):
        assert not out[0].any() and not out[-1].any() and not out[:, 0].any() and not out[:, -1].any()
        for ch in range(1, c):
            assert np.array_equal(out[:, :, 0], out[:, :, ch])
    # level 2 == exact integer Sobel on the u8-rounded gray (image_filters.cu:1443-1444)
    f = img.astype(np.float32)
    gray = (np.float32(0.587) * f[:, :, 1] + np.float32(0.299) * f[:, :, 0]) + np.float32(0.114) * f[:, :, 2]
    g8 = (gray.astype(np.float64) + 0.5).astype(np.uint8)
    approx = np_sobel_gray_int(g8)
    d = np.abs(l2[:, :, 0].astype(int) - approx.astype(int))
    assert d.max() <= 8                                   # unfused numpy gray may round a gray byte differently
    d12 = np.abs(l1.astype(int) - l2.astype(int))
    assert d12.max() <= 6                                 # the two levels differ by a few LSB


def test_white_square_fixture_centre_unchanged():
    """tests/test_gaussian_blur.cu:215-221 checks the centre pixel; with r=3 it cannot change."""
    img = synth.white_square(108, 192, 1)
    out = O.gaussian_blur(img, 2.0, 3)
    assert out[54, 96, 0] == 255 and out[0, 0, 0] == 0
    assert 0 < out[54 - 24, 96, 0] < 255                  # the edge of the square is blurred


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden hashes not minted yet")
def test_oracle_reproduces_reference_gpu_hashes():
    """The pin: outputs of the reference's own level-1/level-2 kernels on a B200, hashed."""
    from tools.mint_golden import CASES, make_input
    gold = json.load(open(GOLDEN))["cases"]
    assert len(gold) >= len(CASES)
    for case in CASES:
        img = make_input(case)
        if case["filter"] == "gaussian":
            out = O.gaussian_blur(img, case["sigma"], case["radius"])
        elif case["filter"] == "box":
            out = O.box_blur(img, case["radius"])
        else:
            out = O.sobel(img, case["level"])
        assert _sha(out) == gold[case["name"]]["sha256"], case["name"]
